#!/bin/sh
timeout 1700 python -m pytest tests -m gpu -q --durations=8 -k "static_schedule" > gpurun_out/r3h_gputests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3h_gputests.log
