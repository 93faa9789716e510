#!/bin/sh
timeout 600 python tools/tran_zero_copy.py c3 5 > gpurun_out/r4c_tran_zero_copy.txt 2>&1
timeout 900 python tools/tran_zero_copy.py c4 3 >> gpurun_out/r4c_tran_zero_copy.txt 2>&1
