#!/bin/sh
# static-schedule LU: replicated shapes on C3 over lanes / block shapes / batch sizes; guess with C1 D
export AHK_DEBUG=1
for B in 16384 8192 4096 2048; do
  timeout 600 python tools/tran_sweep.py c3 $B "9,2,1,64,4;9,2,1,64,0;9,2,1,32,0;9,4,1,64,0;9,4,1,128,0;9,8,1,64,0;9,1,1,32,0" >> gpurun_out/r3f_slu_c3.txt 2>> gpurun_out/r3f_slu_c3.err
done
for B in 16384 4096; do
  timeout 600 python tools/tran_sweep.py c4 $B "5,2,3,64,0;5,4,2,64,0;5,8,1,64,0" >> gpurun_out/r3f_slu_c4.txt 2>> gpurun_out/r3f_slu_c4.err
done
