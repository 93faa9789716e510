#!/bin/sh
timeout 900 python tools/c2_block_sweep.py > gpurun_out/r4f_c2_blocks.txt 2>&1
