#!/bin/sh
# final verification of the round
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/r4k_gputests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r4k_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4k_smoke.log 2>&1; echo "rc $?" >> gpurun_out/r4k_smoke.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r4k_bench_reference.json 2> gpurun_out/r4k_bench_reference.err; echo "rc $?" >> gpurun_out/r4k_bench_reference.err
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r4k_bench.json 2> gpurun_out/r4k_bench.err; echo "rc $?" >> gpurun_out/r4k_bench.err
timeout 600 python bench.py > gpurun_out/r4k_bench_default.json 2> gpurun_out/r4k_bench_default.err; echo "rc $?" >> gpurun_out/r4k_bench_default.err
timeout 600 python bench.py --steps 2 --warmup 3 --only --no-cpu > gpurun_out/r4k_b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r4k_launches_bench_c2.csv python bench.py --steps 2 --warmup 3 --only --no-cpu > gpurun_out/r4k_ncu.log 2>&1
