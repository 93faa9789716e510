set -x
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_gputests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2w_gputests.log
timeout 900 python tools/bign_time.py > gpurun_out/r2w_bign_time.txt 2>&1
for tag in ladder mesh; do
  N=250; [ $tag = mesh ] && N=200
  ncu --set full --clock-control none --import-source on -k regex:k_dc -s 1 -c 1 -o /tmp/r2w_$tag -f python tools/bign_time.py $tag $N 256 > gpurun_out/ncu_r2w_$tag.log 2>&1
  python tools/ncu_summary.py /tmp/r2w_$tag.ncu-rep > gpurun_out/sum_r2w_$tag.txt 2>&1
  python tools/ncu_lines.py /tmp/r2w_$tag.ncu-rep 60 > gpurun_out/lines_r2w_$tag.txt 2>&1
done
