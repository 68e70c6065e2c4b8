#!/bin/bash
# development aid: symbolic-only kernel occupancy variants (launch bounds) built into ppo-2dgrid_b200/lib/variants/
python -m pytest tests/test_gpu_dropin.py tests/test_gpu_variants.py -x -q -m gpu -k "symbolic_only" 2>&1 | tail -1
echo "== default"; python tools/sweep.py --compact --modes symbolic --steps 512 --sizes 65536,262144,1048576 2>&1 | grep "N="
for lib in ppo-2dgrid_b200/lib/variants/lib_sym_*.so; do
  echo "== $lib"
  MERLIN_B200_LIB=$PWD/$lib python -m pytest tests/test_gpu_dropin.py -x -q -m gpu -k "symbolic_only" 2>&1 | tail -1
  MERLIN_B200_LIB=$PWD/$lib python tools/sweep.py --compact --modes symbolic --steps 512 --sizes 65536,262144,1048576 2>&1 | grep "N="
done
