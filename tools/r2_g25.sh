set -x
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2y_gputests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2y_gputests.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench rc $?" >> gpurun_out/r2y_bench.err
