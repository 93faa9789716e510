#!/bin/sh
# Builds ahkab_b200/lib/libahkab_b200.so for sm_100a (nvcc cross-compiles without a GPU).
# The translation units are compiled in parallel, then linked.
set -e
cd "$(dirname "$0")"
mkdir -p ../lib ../lib/obj
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I../lib/obj ${AHK_PTXAS_V:+-Xptxas -v} ${AHK_DEFS}"
# sources of the circuit-specialised kernels, embedded for NVRTC (jit.hpp)
{
  for f in jit_kernels.cuh entry.cuh core.cuh devices.cuh fastmath.cuh ac.cuh program.h ../../include/ahkab_b200.h; do
    printf '{"%s", R"AHKSRC(' "$f"; cat "$f"; printf ')AHKSRC"},\n'
  done
} > ../lib/obj/jit_sources.inc
pids=""
for f in engine k_dc k_tran k_ac; do
  $NVCC $FLAGS -c -o ../lib/obj/$f.o $f.cu > ../lib/obj/$f.log 2>&1 &
  pids="$pids $!"
done
rc=0
for p in $pids; do wait $p || rc=1; done
for f in engine k_dc k_tran k_ac; do cat ../lib/obj/$f.log; done
[ $rc -eq 0 ] || { echo "nvcc failed"; exit 1; }
$NVCC -shared -o ../lib/libahkab_b200.so ../lib/obj/engine.o ../lib/obj/k_dc.o ../lib/obj/k_tran.o ../lib/obj/k_ac.o -lcudart -ldl
echo "built ahkab_b200/lib/libahkab_b200.so"
