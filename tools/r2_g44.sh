#!/bin/sh
# final numbers of the round: bench (all workloads), reference arm, 2-GPU strong-scaling run
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r3r_bench_reference.json 2> gpurun_out/r3r_bench_reference.err; echo "rc $?" >> gpurun_out/r3r_bench_reference.err
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r3r_bench.json 2> gpurun_out/r3r_bench.err; echo "rc $?" >> gpurun_out/r3r_bench.err
