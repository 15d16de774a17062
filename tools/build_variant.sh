#!/bin/bash
# usage: tools/build_variant.sh <name> [nvcc flags...]  -> bear_b200/_variants/libbear_<name>.so (for A/B runs with BEAR_B200_LIB)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p bear_b200/_variants
S=bear_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --std=c++17 -Xcompiler -fPIC -shared -I include -I $S "$@" \
  -o bear_b200/_variants/libbear_$name.so $S/bear_pack.cpp $S/bear_dense.cu $S/bear_fused.cu $S/bear_train.cu $S/bear_heads.cu $S/bear_count.cu $S/bear_cnn.cu
echo built bear_b200/_variants/libbear_$name.so
