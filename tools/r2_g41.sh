#!/bin/sh
timeout 1700 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r3n_gputests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3n_gputests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3n_bench.json 2> gpurun_out/r3n_bench.err; echo "rc $?" >> gpurun_out/r3n_bench.err
