set -x
timeout 600 python tools/c2_shape_sweep.py > gpurun_out/r2s_c2_shapes.txt 2>&1
for B in 2048 4096; do
timeout 900 python tools/tran_sweep.py c4 $B "5,2,3,64,0;5,4,2,64,0;5,8,1,64,0;5,8,1,32,0;5,4,2,32,0;5,2,3,32,0" >> gpurun_out/r2s_c4_small.txt 2>&1
timeout 900 python tools/tran_sweep.py c3 $B "5,2,1,32,0;3,4,1,32,0;3,4,1,64,0;2,8,1,32,0;2,8,1,64,0;1,16,1,32,0" >> gpurun_out/r2s_c3_small.txt 2>&1
done
