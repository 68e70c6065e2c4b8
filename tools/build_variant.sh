#!/bin/bash
# development aid: build libmerlin_b200 with tuning macros into ppo-2dgrid_b200/lib/variants/lib_<name>.so, e.g.
#   tools/build_variant.sh tma_32_128_3_1 -DMERLIN_TMA_T=32 -DMERLIN_TMA_THREADS=128 -DMERLIN_TMA_CTAS=3 -DMERLIN_TMA_NBUF=1
#   tools/build_variant.sh ord_8_128_3    -DMERLIN_ORD_G=8 -DMERLIN_ORD_THREADS=128 -DMERLIN_ORD_CTAS=3
#   tools/build_variant.sh sym_10         -DMERLIN_SYM_MINB=10
# and select it with MERLIN_B200_LIB=$PWD/ppo-2dgrid_b200/lib/variants/lib_<name>.so (tools/*_variants.sh loop over them).
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$root/ppo-2dgrid_b200/lib/variants"
cd "$root/ppo-2dgrid_b200"
${NVCC:-/usr/local/cuda/bin/nvcc} -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-extended-lambda \
  -shared -Xcompiler -fPIC -ccbin /usr/bin/g++ --threads 0 -I../include -Icsrc "$@" -o "lib/variants/lib_$name.so" csrc/*.cu
echo "$root/ppo-2dgrid_b200/lib/variants/lib_$name.so"
