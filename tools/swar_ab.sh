#!/bin/bash
# development aid: A/B of the per-cell and the row-parallel observation path (tile + symbolic-only kernels)
python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -x -q -m gpu -k "tile_kernel or symbolic_only" 2>&1 | tail -3
for path in 1 0; do
  echo "== observation path $path"
  MERLIN_OBSERVATION_PATH=$path python tools/sweep.py --compact --modes symbolic,rgb --steps 256 --sizes 65536,262144,1048576 2>&1 | grep "N="
done
