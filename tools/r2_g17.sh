set -x
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2r_bench_n$N.json 2> gpurun_out/r2r_bench_n$N.err; echo "rc $?" >> gpurun_out/r2r_bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 3 --warmup 1 --only > gpurun_out/r2r_bench_ref_n$N.json 2> gpurun_out/r2r_bench_ref_n$N.err; echo "rc $?" >> gpurun_out/r2r_bench_ref_n$N.err
