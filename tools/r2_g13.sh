set -x
export AHK_DEBUG=1
timeout 1500 python -m pytest tests -m gpu -x -q -k "ladder or mesh" > gpurun_out/r2n_gputests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2n_gputests.log
unset AHK_DEBUG
timeout 600 python - > gpurun_out/r2n_bign_time.txt 2>&1 <<'PY'
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
import zoo
from ahkab_b200 import time_functions as TF
from ahkab_b200.circuit import Circuit
from ahkab_b200.batch import ParamBatch
from ahkab_b200.engine import Engine
for name, build, N in (('ladder', zoo.diode_ladder, 100), ('ladder', zoo.diode_ladder, 250), ('mesh', zoo.diode_mesh, 72), ('mesh', zoo.diode_mesh, 200)):
    for B in (256, 4096):
        c = build(Circuit, TF, N)
        pb = ParamBatch(B)
        rng = np.random.default_rng(0)
        first = [e for e in c if e.part_id.lower().startswith('r')][:8]
        for e in first:
            pb.set(e.part_id, 'value', e.value * (1 + 0.05 * rng.standard_normal(B)))
        eng = Engine(c, pb)
        r = eng.op(); torch.cuda.synchronize()
        t0 = time.perf_counter(); r = eng.op(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        it = r['iters'].cpu().numpy(); st = r['status'].cpu().numpy()
        print(name, 'N', N, 'n', eng.n, 'B', B, 'op %.2f ms' % (dt * 1e3), 'iters mean %.1f' % it.mean(), 'ok', int((st & 1).sum()), flush=True)
PY
