#!/bin/bash
# Builds libb200ivfpq.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
OUT=../b200ivfpq/libb200ivfpq.so
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
     -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v \
     --shared -cudart static -o "$OUT" api.cu "$@"
echo "built $OUT"
