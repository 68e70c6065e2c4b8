#!/bin/bash
# development aid: GAE chunk-depth variants (tools/build_variant.sh gae_<small>_<large> -DMERLIN_GAE_CHUNK_SMALL=.. -DMERLIN_GAE_CHUNK_LARGE=..)
echo "== default"; python tools/bench_gae.py 2>&1 | grep -E '"T"|T=' | head -20
for lib in ppo-2dgrid_b200/lib/variants/lib_gae_*.so; do
  echo "== $lib"; MERLIN_B200_LIB=$PWD/$lib python tools/bench_gae.py 2>&1 | grep -E '"T"|T=' | head -20
done
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "gae and tile_kernel" 2>&1 | tail -1
