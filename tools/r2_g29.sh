#!/bin/sh
# compiled C5 solve with explicit scratch + block-wise loads: parity tests, KB sweep
timeout 900 python -m pytest tests -m gpu -x -q -k "c5" > gpurun_out/r3b_c5_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3b_c5_tests.log
for kb in 8 4 16 32; do
  echo "KB=$kb" >> gpurun_out/r3b_c5_kb.txt
  AHK_B3_KB=$kb timeout 600 python bench.py --workload c5 --only --no-cpu --steps 20 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['details'].get('phases'))" >> gpurun_out/r3b_c5_kb.txt 2>&1
done
