set -x
AHK_JIT=on AHK_SHAPE=5,2,3,64,4 sh tools/prof_box.sh r2f_c4 ahk_jit_tran c4 -1 2 16384
