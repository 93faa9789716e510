#!/bin/sh
rm -f gpurun_out/r3w_c5.txt
run() { echo "$*" >> gpurun_out/r3w_c5.txt; env "$@" timeout 600 python bench.py --workload c5 --only --no-cpu --steps 20 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'])" >> gpurun_out/r3w_c5.txt 2>&1; }
run AHK_B3_KBF=16
run AHK_B3_KBF=32
run AHK_B3_KBF=8
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ahk_big3c -c 6 --csv --log-file gpurun_out/r3w_c5_launches.csv python tools/ncu_target.py c5 -1 3 > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -k "c5 or bign or big" > gpurun_out/r3w_c5_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3w_c5_tests.log
