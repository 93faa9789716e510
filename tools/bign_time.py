"""Timing of the general large-n engine (n > 64, non-linear): python tools/bign_time.py [ladder|mesh] N B"""
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
import zoo
from ahkab_b200 import time_functions as TF
from ahkab_b200.circuit import Circuit
from ahkab_b200.batch import ParamBatch
from ahkab_b200.engine import Engine


def run(name, N, B, reps=2):
    build = zoo.diode_ladder if name == 'ladder' else zoo.diode_mesh
    c = build(Circuit, TF, N)
    pb = ParamBatch(B)
    rng = np.random.default_rng(0)
    for e in [e for e in c if e.part_id.lower().startswith('r')][:8]:
        pb.set(e.part_id, 'value', e.value * (1 + 0.05 * rng.standard_normal(B)))
    eng = Engine(c, pb)
    r = eng.op(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = eng.op()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    it = r['iters'].cpu().numpy(); st = r['status'].cpu().numpy()
    print(name, 'N', N, 'n', eng.n, 'B', B, 'op %.2f ms' % (dt * 1e3), 'iters mean %.1f' % it.mean(),
          'ok', int((st & 1).sum()), '%.3g instance-NR-iterations/s' % (it.sum() / dt), flush=True)
    return r


if __name__ == '__main__':
    if len(sys.argv) > 1:
        run(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))
    else:
        for name, N in (('ladder', 100), ('ladder', 250), ('mesh', 72), ('mesh', 200), ('mesh', 300)):
            for B in (256, 4096):
                if name == 'mesh' and N == 300 and B == 4096:
                    continue
                run(name, N, B)
