set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_gputests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2a_gputests.log
AHK_JIT=on sh tools/prof_box.sh r2a_c4full ahk_jit_tran c4 -1 2 16384
AHK_JIT=off sh tools/prof_box.sh r2a_c5f k_big3_factor c5 -1 2
AHK_JIT=off sh tools/prof_box.sh r2a_c5s k_big3_solve c5 -1 2
