set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_gputests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2k_gputests.log
for pw in 1 2 5 8; do AHK_BIGLIN3C_PW=$pw timeout 600 python bench.py --workload c5 --only --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pw $pw', d['ms_per_step'], d['e2e']['ms_per_step'], d.get('e2e_api'))"; done > gpurun_out/r2k_pw.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc $?" >> gpurun_out/r2k_bench.err
python - > gpurun_out/r2k_api_profile.txt 2>&1 <<'PY'
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
pstats.Stats(pr).sort_stats('cumulative').print_stats(35)
PY
