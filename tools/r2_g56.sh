#!/bin/sh
timeout 1200 python tools/sanitize_target.py > gpurun_out/r4d_plain.log 2>&1; echo "rc $?" >> gpurun_out/r4d_plain.log
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_target.py > gpurun_out/r4d_memcheck.log 2>&1; echo "rc $?" >> gpurun_out/r4d_memcheck.log
