#!/bin/sh
AHK_AC_CHUNKS=32 sh tools/prof_box.sh r3j_c1 ahk_jit_ac c1 -1 3
