#!/bin/bash
# sweep every kernel mapping (development aid): MERLIN_KERNEL_CHOICE 1 group / 2 warp / 3 tile
for k in ${KERNELS:-1 2 3}; do
  echo "== kernel choice $k"
  MERLIN_KERNEL_CHOICE=$k python tools/sweep.py --compact --modes ${MODES:-rgb} --steps ${STEPS:-256} --sizes ${SIZES:-4096,8192,16384,32768,65536,131072,262144,1048576} 2>&1 | grep "N="
done
