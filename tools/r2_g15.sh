set -x
timeout 1500 python -m pytest tests -m gpu -x -q -k "ladder or mesh or bign or large" > gpurun_out/r2p_gputests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2p_gputests.log
timeout 900 python tools/bign_time.py > gpurun_out/r2p_bign_time.txt 2>&1
M=sm__inst_executed_pipe_tensor_subpipe_dmma.sum,smsp__pipe_tensor_subpipe_dmma_cycles_active.sum,sm__inst_executed_pipe_fp64.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__cycles_active.avg,lts__t_sector_hit_rate.pct
timeout 600 ncu --metrics $M --clock-control none -k regex:k_dc -c 2 python tools/bign_time.py mesh 200 256 > gpurun_out/r2p_ncu_mesh.txt 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:k_dc -c 2 python tools/bign_time.py ladder 250 256 > gpurun_out/r2p_ncu_ladder.txt 2>&1
