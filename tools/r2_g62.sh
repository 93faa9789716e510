#!/bin/sh
rm -f gpurun_out/r4j_c1.txt
run() { echo "$*" >> gpurun_out/r4j_c1.txt; env "$@" timeout 600 python bench.py --workload c1 --only --no-cpu --steps 20 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'])" >> gpurun_out/r4j_c1.txt 2>&1; }
run AHK_X=1
run AHK_AC_MINB=2 AHK_AC_THREADS=128
run AHK_AC_MINB=3 AHK_AC_THREADS=128
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,launch__registers_per_thread,l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum --clock-control none -k regex:ahk_jit_ac -s 2 -c 1 --csv --log-file gpurun_out/r4j_c1_launches.csv python tools/ncu_target.py c1 -1 3 > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -k "c1 or ac" > gpurun_out/r4j_ac_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r4j_ac_tests.log
