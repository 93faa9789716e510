#!/bin/sh
timeout 900 python -m pytest tests -m gpu -x -q -k "c1 or ac" > gpurun_out/r4g_ac_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r4g_ac_tests.log
timeout 600 python tools/ac_zero_copy.py 10 > gpurun_out/r4g_ac_zero_copy.txt 2>&1
timeout 600 python bench.py --workload c1 --only --no-cpu --steps 20 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'])" >> gpurun_out/r4g_ac_zero_copy.txt 2>&1
