set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_gputests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2j_gputests.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc $?" >> gpurun_out/r2j_bench.err
sh tools/prof_box.sh r2j_c5f ahk_big3c_factor c5 -1 2
sh tools/prof_box.sh r2j_c5s ahk_big3c_solve c5 -1 2
