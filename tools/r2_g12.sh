set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "api or host or results or abi or dist" > gpurun_out/r2l_gputests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2l_gputests.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc $?" >> gpurun_out/r2l_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2l_bench_ref.json 2> gpurun_out/r2l_bench_ref.err; echo "bench rc $?" >> gpurun_out/r2l_bench_ref.err
python - > gpurun_out/r2l_api_profile.txt 2>&1 <<'PY'
import sys, time, cProfile, pstats
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
import ahkab_b200 as ahkab
from ahkab_b200 import workloads as W
from ahkab_b200.circuit import Circuit
w = W.c2_diode_op(65536)
circ = w.circuit(Circuit); pb = w.batch()
def once():
    r = ahkab.run(circ, ahkab.new_op(), batch=pb)['op']
    return r.x, r.iterations, r.status
for _ in range(5): once()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): once()
print('api ms/step', (time.perf_counter() - t0) / 50 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(50): once()
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(30)
PY
