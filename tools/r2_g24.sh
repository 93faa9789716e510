set -x
timeout 1500 python -m pytest tests -m gpu -x -q -k "ladder or mesh or bign or large" > gpurun_out/r2x_gputests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2x_gputests.log
AHK_DEBUG=1 timeout 900 python tools/bign_time.py > gpurun_out/r2x_bign_time.txt 2>&1
