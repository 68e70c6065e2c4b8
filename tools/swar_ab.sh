#!/bin/bash
# development aid: A/B of the observation paths (1 = per-cell everywhere, 0 = automatic, 2 = row-parallel wherever built)
python -m pytest tests/test_gpu_variants.py tests/test_gpu_dropin.py -x -q -m gpu -k "row_obs or symbolic_only" 2>&1 | tail -2
for path in 1 0 2; do
  echo "== observation path $path"
  MERLIN_OBSERVATION_PATH=$path python tools/sweep.py --compact --modes symbolic,rgb --steps 256 --sizes 262144,1048576 2>&1 | grep "N="
done
