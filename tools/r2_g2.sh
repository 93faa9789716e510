set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_gputests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2b_gputests.log
timeout 600 python tools/tran_sweep.py c4 0 "2,4,1,128,4;2,4,1,64,7;2,4,2,64,7;5,1,1,32,4;5,1,2,32,4;5,1,3,32,4;5,2,1,32,7;5,2,3,32,7;3,2,1,32,7;3,2,3,32,7;1,8,1,128,4" > gpurun_out/r2b_sweep_c4.txt 2>&1
timeout 300 python tools/tran_sweep.py c3 0 "3,4,1,128,4;3,4,1,64,7;5,2,1,32,7;9,2,1,32,7;9,1,1,32,4;2,8,1,128,4" > gpurun_out/r2b_sweep_c3.txt 2>&1
