#!/bin/sh
export AHK_DEBUG=1
rm -f gpurun_out/r3q_c2.txt
run() { echo "$*" >> gpurun_out/r3q_c2.txt; env "$@" timeout 600 python bench.py --workload c2 --only --no-cpu --steps 50 --warmup 5 2>gpurun_out/r3q_c2.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['details'].get('kernel_shape_rank0'), d['details']['status_ok_rank0'])" >> gpurun_out/r3q_c2.txt 2>&1; grep -i "calibration" gpurun_out/r3q_c2.err | sort | uniq -c >> gpurun_out/r3q_c2.txt; }
run AHK_SLU=tran
run AHK_SLU=all
run AHK_SLU=all
timeout 900 python -m pytest tests -m gpu -x -q -k "static_schedule or c3_full or c4_full" > gpurun_out/r3q_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3q_tests.log
