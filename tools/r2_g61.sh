#!/bin/sh
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r4i_bench_8gpu.json 2> gpurun_out/r4i_bench_8gpu.err; echo "rc $?" >> gpurun_out/r4i_bench_8gpu.err
