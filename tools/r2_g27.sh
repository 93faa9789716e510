set -x
for i in 1 2 3; do timeout 600 python bench.py --steps 50 --warmup 5 --only --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'])"; done > gpurun_out/r2z2_c2_repeat.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -k "jit or bench_parity or c2" > gpurun_out/r2z2_gputests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2z2_gputests.log
