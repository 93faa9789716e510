set -x
export AHK_DEBUG=1
timeout 1500 python -m pytest tests -m gpu -x -q -k "c5 or static" > gpurun_out/r2i_gputests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2i_gputests.log
unset AHK_DEBUG
timeout 900 python bench.py --workload c5 --only --steps 20 --warmup 3 --no-cpu > gpurun_out/r2i_bench_c5.json 2> gpurun_out/r2i_bench_c5.err; echo "bench rc $?" >> gpurun_out/r2i_bench_c5.err
for p in 1 2 3 5 10; do AHK_BIGLIN3C_PTC=$p timeout 600 python bench.py --workload c5 --only --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ptc $p', d['ms_per_step'], d['e2e']['ms_per_step'])"; done > gpurun_out/r2i_ptc.txt 2>&1
