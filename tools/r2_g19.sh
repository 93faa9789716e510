set -x
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_gputests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2t_gputests.log
timeout 600 python tools/c2_shape_sweep.py > gpurun_out/r2t_c2_shapes.txt 2>&1
