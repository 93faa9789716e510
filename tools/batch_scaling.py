"""How the kernels scale with the batch size (CUDA events, inputs resident): at the
BASELINE sizes the machine is far from full; this shows where each kernel saturates.
python tools/batch_scaling.py [c2 c1 c3 c4 c5]"""
import sys

import torch

sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import common  # noqa
from ahkab_b200 import workloads as W
from ahkab_b200.circuit import Circuit
from ahkab_b200.engine import Engine

SIZES = dict(c2=[16384, 65536, 262144, 1048576, 4194304], c1=[1024, 4096, 16384, 65536],
             c3=[4096, 16384, 65536, 262144], c4=[2048, 16384, 65536], c5=[1024, 4096, 16384])


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


for name in (sys.argv[1:] or ['c2', 'c1', 'c3', 'c4', 'c5']):
    for B in SIZES[name]:
        w = W.ALL[name](B)
        circ = w.circuit(Circuit)
        e = Engine(circ, w.batch(), opts=w.opts)
        r = {}
        if name == 'c2':
            ms = timeit(lambda: e.op(want_resid=False)); units = B
        elif name == 'c1':
            om = common.c1_omegas(); ms = timeit(lambda: e.ac(om), 2); units = B * len(om)
        elif name == 'c3':
            x0 = torch.as_tensor(common.c3_x0(e.op()['x'].cpu().numpy())).cuda()
            def f():
                r['t'] = e.tran(x0, max_rows=256, **W.C3_TRAN)
            ms = timeit(f, 2); units = int(r['t']['rows'].sum())
        elif name == 'c4':
            x0 = torch.as_tensor(common.c4_x0(circ, B)).cuda()
            def f():
                r['t'] = e.tran(x0, max_rows=2560, **W.C4_TRAN)
            ms = timeit(f, 1); units = int(r['t']['rows'].sum())
        else:
            sw = common.c5_sweep(); ms = timeit(lambda: e.dc_sweep('V1', sw), 2); units = B * len(sw)
        print('%s B=%8d  %10.3f ms  %.4g units/s' % (name, B, ms, units / ms * 1e3), flush=True)
        del e, r
        torch.cuda.empty_cache()
