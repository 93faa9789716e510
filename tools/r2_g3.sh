set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_gputests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2c_gputests.log
timeout 900 python tools/tran_sweep.py c4 0 "2,4,1,128,4;2,4,2,64,7;5,1,1,32,4;5,1,2,32,4;5,1,3,32,4;5,2,1,32,7;5,2,3,32,7;3,2,3,32,7;5,2,3,64,4;5,1,3,64,2" > gpurun_out/r2c_sweep_c4.txt 2>&1
timeout 300 python tools/tran_sweep.py c3 0 "3,4,1,128,4;5,2,1,32,7;3,4,1,64,7" > gpurun_out/r2c_sweep_c3.txt 2>&1
