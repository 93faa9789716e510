set -x
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_gputests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2w_gputests.log
timeout 900 python tools/bign_time.py > gpurun_out/r2w_bign_time.txt 2>&1
