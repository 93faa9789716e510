#!/bin/sh
rm -f gpurun_out/r4e_c5.txt
run() { echo "$*" >> gpurun_out/r4e_c5.txt; env "$@" timeout 600 python bench.py --workload c5 --only --no-cpu --steps 20 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['details']['status_ok_rank0'])" >> gpurun_out/r4e_c5.txt 2>&1; }
run AHK_X=1
run AHK_B3_POINT_MINOR=1
run AHK_B3_POINT_MINOR=1 AHK_B3_KB=12
run AHK_B3_POINT_MINOR=1 AHK_BIGLIN3C_THREADS=32
AHK_B3_POINT_MINOR=1 timeout 600 python -m pytest tests -m gpu -x -q -k "c5" > gpurun_out/r4e_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r4e_tests.log
AHK_B3_POINT_MINOR=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:ahk_big3c_solve -s 1 -c 1 --csv --log-file gpurun_out/r4e_c5_launches.csv python tools/ncu_target.py c5 -1 3 > /dev/null 2>&1
