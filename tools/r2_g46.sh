#!/bin/sh
timeout 900 python tools/tran_sweep.py c4 0 "5,2,3,64,4;5,2,3,64,3;5,2,3,64,5;5,2,3,32,8;5,4,2,64,4;5,4,2,64,6;5,4,2,64,8;5,4,2,32,12;5,8,1,64,8;5,2,2,64,4;5,2,1,64,4" > gpurun_out/r3t_c4_shapes.txt 2> gpurun_out/r3t_c4_shapes.err
