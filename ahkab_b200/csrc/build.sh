#!/bin/sh
# Builds ahkab_b200/lib/libahkab_b200.so for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
mkdir -p ../lib
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC -shared ${AHK_PTXAS_V:+-Xptxas -v} \
  -o ../lib/libahkab_b200.so engine.cu -lcudart
echo "built ahkab_b200/lib/libahkab_b200.so"
