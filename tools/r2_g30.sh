#!/bin/sh
export AHK_B3_KB=16
sh tools/prof_box.sh r3c_c5s ahk_big3c_solve c5 -1 2
sh tools/prof_box.sh r3c_c5f ahk_big3c_factor c5 -1 2
