#!/bin/sh
# r2 re-entry: ncu --set full of the timed C3 and C2 specialised kernels (source lines)
sh tools/prof_box.sh r3a_c3 ahk_jit_tran c3 -1 2
sh tools/prof_box.sh r3a_c2 ahk_jit_dc c2 -1 3
