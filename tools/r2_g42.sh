#!/bin/sh
# final ncu --set full captures of the timed kernels (static-schedule builds), launch list of the bench command
cap() { tag=$1; pat=$2; skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c 1 -o /tmp/$tag -f python tools/ncu_target.py "$@" > gpurun_out/ncu_$tag.log 2>&1
  python tools/ncu_summary.py /tmp/$tag.ncu-rep > gpurun_out/sum_$tag.txt 2>&1
  python tools/ncu_lines.py /tmp/$tag.ncu-rep 50 > gpurun_out/lines_$tag.txt 2>&1; }
cap r3p_c3 ahk_jit_tran 3 c3 -1 3
cap r3p_c4 ahk_jit_tran 3 c4 -1 3
cap r3p_c5s ahk_big3c_solve 1 c5 -1 3
cap r3p_c5f ahk_big3c_factor 1 c5 -1 3
timeout 600 python bench.py --steps 2 --warmup 3 --only --no-cpu > gpurun_out/r3p_b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3p_launches_bench_c2.csv python bench.py --steps 2 --warmup 3 --only --no-cpu > gpurun_out/r3p_ncu.log 2>&1
