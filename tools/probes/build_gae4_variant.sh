#!/bin/bash
# Builds libg2048 with extra -D flags for ONE source file (default g2048_gae4.cu; SRC=g2048_gae to pick another):
#   [SRC=g2048_gae] tools/probes/build_gae4_variant.sh NAME -DFLAG=...  -> gpurun_variants/libg2048_NAME.so
set -e
cd "$(dirname "$0")/../../2048-ppo-agent_b200/csrc"
name=$1; shift
tmp=$(mktemp -d)
src=${SRC:-g2048_gae4}
for f in g2048_env g2048_play g2048_play3 g2048_stream g2048_policy g2048_data g2048_gae g2048_gae3 g2048_gae4 g2048_embed; do [ $f = $src ] || cp build/$f.o $tmp/; done
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --fmad=false "$@" -c $src.cu -o $tmp/$src.o
mkdir -p ../../gpurun_variants
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../gpurun_variants/libg2048_$name.so $tmp/*.o -lcudart
rm -rf $tmp
