#!/bin/bash
# run the sweep for every kernel build variant under ppo-2dgrid_b200/lib/variants (development aid)
for so in ppo-2dgrid_b200/lib/libmerlin_b200.so ppo-2dgrid_b200/lib/variants/*.so; do
  echo "== $so"
  MERLIN_B200_LIB=$PWD/$so MERLIN_KERNEL_CHOICE=${CHOICE:-3} python tools/sweep.py --compact --modes ${MODES:-rgb} --steps ${STEPS:-256} --sizes ${SIZES:-4096,16384,65536,262144,1048576} 2>&1 | grep "N="
done
