#!/bin/sh
tag=r3l_c1
ncu --set full --clock-control none --import-source on -k regex:ahk_jit_ac -s 3 -c 1 -o /tmp/$tag -f python tools/ncu_target.py c1 -1 3 > gpurun_out/ncu_$tag.log 2>&1
python tools/ncu_summary.py /tmp/$tag.ncu-rep > gpurun_out/sum_$tag.txt 2>&1
python tools/ncu_lines.py /tmp/$tag.ncu-rep 40 > gpurun_out/lines_$tag.txt 2>&1
