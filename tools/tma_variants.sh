#!/bin/bash
# development aid: parity subset + throughput of the TMA tile kernel shapes built into ppo-2dgrid_b200/lib/variants/
echo "== shipped tile kernel (choice 3)"
MERLIN_KERNEL_CHOICE=3 python tools/sweep.py --compact --modes rgb --steps 256 --sizes 65536,262144,1048576 2>&1 | grep "N="
for lib in ppo-2dgrid_b200/lib/variants/lib_tma_*.so; do
  echo "== $lib"
  MERLIN_B200_LIB=$PWD/$lib python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tile_tma_kernel and (trace_autoreset or random_rollout_all or seven_actions or outputs_stay or masked_reset)" 2>&1 | tail -1
  MERLIN_B200_LIB=$PWD/$lib MERLIN_KERNEL_CHOICE=4 python tools/sweep.py --compact --modes rgb --steps 256 --sizes 65536,262144,1048576 2>&1 | grep "N="
done
