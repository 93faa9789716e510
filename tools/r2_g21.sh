set -x
timeout 900 python tools/tran_sweep.py c3 16384 "5,2,1,32,0;5,2,1,64,0;5,2,1,32,7;5,2,1,32,8;5,2,1,64,4" > gpurun_out/r2v_c3_blocks.txt 2>&1
timeout 900 python tools/tran_sweep.py c4 16384 "5,2,3,32,0;5,2,3,64,0;5,2,3,32,7;5,2,3,32,8" > gpurun_out/r2v_c4_blocks.txt 2>&1
timeout 900 python tools/tran_sweep.py c3 8192 "5,2,1,32,0;5,2,1,64,0;3,4,1,32,0;3,4,1,64,0" >> gpurun_out/r2v_c3_blocks.txt 2>&1
timeout 900 python tools/tran_sweep.py c4 8192 "5,2,3,32,0;5,2,3,64,0;5,4,2,32,0;5,4,2,64,0;5,8,1,32,0" >> gpurun_out/r2v_c4_blocks.txt 2>&1
