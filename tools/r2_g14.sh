set -x
timeout 1500 python -m pytest tests -m gpu -x -q -k "ladder or mesh" > gpurun_out/r2o_gputests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2o_gputests.log
timeout 900 python tools/bign_time.py > gpurun_out/r2o_bign_time.txt 2>&1
ncu --query-metrics 2>/dev/null | grep -i "dmma\|pipe_fp64\|tensor_op" | head -40 > gpurun_out/r2o_metrics.txt
timeout 600 ncu --metrics smsp__inst_executed_pipe_tensor_op_dmma.sum,sm__inst_executed_pipe_tensor_op_dmma.sum,smsp__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fp64.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:k_dc -c 2 python tools/bign_time.py mesh 200 256 > gpurun_out/r2o_ncu_mesh.txt 2>&1
timeout 600 ncu --metrics smsp__inst_executed_pipe_tensor_op_dmma.sum,sm__inst_executed_pipe_tensor_op_dmma.sum,smsp__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fp64.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:k_dc -c 2 python tools/bign_time.py ladder 250 256 > gpurun_out/r2o_ncu_ladder.txt 2>&1
