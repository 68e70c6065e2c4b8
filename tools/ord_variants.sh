#!/bin/bash
# development aid: parity subset + throughput of the ordered-group kernel shapes built into ppo-2dgrid_b200/lib/variants/
for lib in ppo-2dgrid_b200/lib/variants/lib_ord_*.so; do
  echo "== $lib"
  MERLIN_B200_LIB=$PWD/$lib python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ordered_kernel and (trace_autoreset or random_rollout_all or seven_actions or outputs_stay or masked_reset)" 2>&1 | tail -1
  MERLIN_B200_LIB=$PWD/$lib MERLIN_KERNEL_CHOICE=6 python tools/sweep.py --compact --modes rgb --steps 256 --sizes 65536,262144,1048576 2>&1 | grep "N="
done
