#!/bin/bash
# Builds tools/experiments/libbear_b200_exp.so = the product library + -DBEAR_TRAIN_EXPERIMENTS (scatter-off knob,
# per-role cycle counters of linear_train2_kernel).  Use it with BEAR_B200_LIB=$PWD/tools/experiments/libbear_b200_exp.so.
set -e
cd "$(dirname "$0")/../.."
S=bear_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --std=c++17 -Xcompiler -fPIC -shared -DBEAR_TRAIN_EXPERIMENTS \
    -I include -I $S -o tools/experiments/libbear_b200_exp.so \
    $S/bear_pack.cpp $S/bear_dense.cu $S/bear_fused.cu $S/bear_heads.cu $S/bear_count.cu $S/bear_cnn.cu
echo built tools/experiments/libbear_b200_exp.so
