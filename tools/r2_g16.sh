set -x
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_gputests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2z_gputests.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err; echo "bench rc $?" >> gpurun_out/r2z_bench_ref.err
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc $?" >> gpurun_out/r2z_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/r2z_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches_bench_c2.csv python bench.py --steps 2 --warmup 3 --only --no-cpu > gpurun_out/r2z_ncu_bench.log 2>&1
