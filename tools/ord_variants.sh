#!/bin/bash
# development aid: parity subset + throughput of the ordered-group kernel shapes built into ppo-2dgrid_b200/lib/variants/
for lib in ppo-2dgrid_b200/lib/variants/lib_ord_*.so; do
  echo "== $lib"
  MERLIN_B200_LIB=$PWD/$lib python -m pytest tests/test_gpu_variants.py -x -q -m gpu -k "ordered_kernel" 2>&1 | tail -1
  MERLIN_B200_LIB=$PWD/$lib MERLIN_KERNEL_CHOICE=6 python tools/sweep.py --compact --modes rgb --steps 256 --sizes 262144,1048576 2>&1 | grep "N="
done
