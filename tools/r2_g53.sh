#!/bin/sh
timeout 900 python -m pytest tests -m gpu -x -q -k "c2 or op_host or specialised_op or diode or edge" > gpurun_out/r4a_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r4a_tests.log
rm -f gpurun_out/r4a_c2.txt
run() { echo "$*" >> gpurun_out/r4a_c2.txt; env "$@" timeout 600 python bench.py --workload c2 --only --no-cpu --steps 50 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e'].get('h2d_bytes_per_step'), d['e2e'].get('d2h_bytes_per_step'))" >> gpurun_out/r4a_c2.txt 2>&1; }
run AHK_X=1
run AHK_E2E_OP_HOST=0
run AHK_E2E_OP_HOST=0
run AHK_E2E_OP_HOST=1
run AHK_E2E_OP_HOST=2
