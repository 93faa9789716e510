set -x
timeout 900 python tools/tran_sweep.py c4 0 "5,2,3,64,4;5,2,2,64,4;5,2,3,32,7;5,1,2,32,4;5,1,3,32,4;5,4,2,64,7;5,4,1,64,7;2,4,2,64,7" > gpurun_out/r2e_sweep_c4.txt 2>&1
timeout 300 python tools/tran_sweep.py c3 0 "3,4,1,128,4;5,2,1,32,7;9,2,1,32,7;9,4,1,64,7" > gpurun_out/r2e_sweep_c3.txt 2>&1
