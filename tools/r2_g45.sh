#!/bin/sh
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r3s_bench_2gpu.json 2> gpurun_out/r3s_bench_2gpu.err; echo "rc $?" >> gpurun_out/r3s_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r3s_bench_2gpu_reference.json 2> gpurun_out/r3s_bench_2gpu_reference.err; echo "rc $?" >> gpurun_out/r3s_bench_2gpu_reference.err
