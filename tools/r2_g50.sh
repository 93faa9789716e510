#!/bin/sh
cap() { tag=$1; pat=$2; skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c 1 -o /tmp/$tag -f python tools/ncu_target.py "$@" > gpurun_out/ncu_$tag.log 2>&1
  python tools/ncu_summary.py /tmp/$tag.ncu-rep > gpurun_out/sum_$tag.txt 2>&1; }
cap r3x_c5s ahk_big3c_solve 1 c5 -1 3
cap r3x_c5f ahk_big3c_factor 1 c5 -1 3
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/r3x_gputests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3x_gputests.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r3x_bench_reference.json 2> gpurun_out/r3x_bench_reference.err; echo "rc $?" >> gpurun_out/r3x_bench_reference.err
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r3x_bench.json 2> gpurun_out/r3x_bench.err; echo "rc $?" >> gpurun_out/r3x_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3x_smoke.log 2>&1; echo "rc $?" >> gpurun_out/r3x_smoke.log
