#!/bin/sh
# static-schedule LU (core.cuh: slu_solve_thread): C3 / C4 shapes with and without it
export AHK_DEBUG=1
for slu in tran 0; do
  echo "AHK_SLU=$slu" >> gpurun_out/r3e_slu_c3.txt
  AHK_SLU=$slu timeout 600 python tools/tran_sweep.py c3 0 "9,1,1,64,4;9,1,1,32,8;9,1,1,64,8;9,2,1,64,4;5,2,1,32,7" >> gpurun_out/r3e_slu_c3.txt 2>> gpurun_out/r3e_slu_c3.err
  echo "AHK_SLU=$slu" >> gpurun_out/r3e_slu_c4.txt
  AHK_SLU=$slu timeout 600 python tools/tran_sweep.py c4 0 "5,2,3,64,0;5,1,3,64,0;5,4,2,64,0" >> gpurun_out/r3e_slu_c4.txt 2>> gpurun_out/r3e_slu_c4.err
done
