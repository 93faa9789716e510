#!/bin/sh
# static-schedule LU as the default of the transient kernels: full GPU suite + bench
timeout 1700 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r3g_gputests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3g_gputests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3g_bench.json 2> gpurun_out/r3g_bench.err; echo "rc $?" >> gpurun_out/r3g_bench.err
