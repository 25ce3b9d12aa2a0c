#!/bin/bash
# Experimental build of the library with extra -D flags (A/B runs on the GPU box: ABEA_LIB=<path> selects it).
#   tools/build_variant.sh NAME -DABEA_WIDE_SYNC=0 ...   ->  f5c_b200/lib/exp/libabea_NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p f5c_b200/lib/exp
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared "$@" \
     -o f5c_b200/lib/exp/libabea_$name.so f5c_b200/csrc/abea_host.cu
echo built f5c_b200/lib/exp/libabea_$name.so
