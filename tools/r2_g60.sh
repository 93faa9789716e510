#!/bin/sh
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r4h_bench_4gpu.json 2> gpurun_out/r4h_bench_4gpu.err; echo "rc $?" >> gpurun_out/r4h_bench_4gpu.err
