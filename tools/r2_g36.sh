#!/bin/sh
# AC: lanes of a group = frequency chunks of one instance, zgesv along a static schedule (trie of learned sequences)
export AHK_DEBUG=1
timeout 900 python -m pytest tests -m gpu -x -q -k "c1 or ac" > gpurun_out/r3i_ac_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3i_ac_tests.log
rm -f gpurun_out/r3i_c1.txt
run() { echo "$*" >> gpurun_out/r3i_c1.txt; env "$@" timeout 600 python bench.py --workload c1 --only --no-cpu --steps 20 --warmup 3 2>gpurun_out/r3i_c1.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['details'].get('kernel_shape_rank0'))" >> gpurun_out/r3i_c1.txt 2>&1; grep -i "calibration" gpurun_out/r3i_c1.err | sort | uniq -c >> gpurun_out/r3i_c1.txt; }
run AHK_SLU=tran
run AHK_AC_CHUNKS=4
run AHK_AC_CHUNKS=8
run AHK_AC_CHUNKS=16
run AHK_AC_CHUNKS=16 AHK_AC_THREADS=32
run AHK_AC_CHUNKS=16 AHK_AC_THREADS=128
run AHK_AC_CHUNKS=32
run AHK_AC_CHUNKS=32 AHK_AC_THREADS=128
