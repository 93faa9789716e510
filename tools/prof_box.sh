#!/bin/sh
# usage (on the GPU box): sh tools/prof_box.sh <tag> <kernel regex> <ncu_target args...>
# Captures one launch with ncu --set full and brings back only text summaries.
tag=$1; shift; pat=$1; shift
python tools/ncu_target.py "$@" > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$pat -s 1 -c 1 -o /tmp/$tag -f \
  python tools/ncu_target.py "$@" > gpurun_out/ncu_$tag.log 2>&1
python tools/ncu_summary.py /tmp/$tag.ncu-rep > gpurun_out/sum_$tag.txt 2>&1
python tools/ncu_lines.py /tmp/$tag.ncu-rep 70 > gpurun_out/lines_$tag.txt 2>&1
python tools/ncu_lines.py /tmp/$tag.ncu-rep 40 stall_no_inst > gpurun_out/noinst_$tag.txt 2>&1
ls -la /tmp/$tag.ncu-rep
