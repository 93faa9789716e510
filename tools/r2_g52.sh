#!/bin/sh
rm -f gpurun_out/r3z_c5.txt
run() { echo "$*" >> gpurun_out/r3z_c5.txt; env "$@" timeout 600 python bench.py --workload c5 --only --no-cpu --steps 20 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'])" >> gpurun_out/r3z_c5.txt 2>&1; }
run AHK_B3_KB=8
run AHK_B3_KB=12
run AHK_B3_KB=16
run AHK_B3_KB=24
run AHK_B3_KB=12 AHK_BIGLIN3C_THREADS=32
run AHK_B3_KB=12 AHK_BIGLIN3C_THREADS=128
run AHK_B3_KB=12 AHK_BIGLIN3C_PTC=5
