"""Times every kernel variant that fits a workload, interpreted and circuit-specialised
(CUDA events, inputs resident, best of N).  python tools/variant_sweep.py c4 [B]"""
import sys

import torch

sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import common  # noqa
from ahkab_b200 import _abi, workloads as W
from ahkab_b200.circuit import Circuit
from ahkab_b200.engine import Engine, EngineError


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


name = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else None
only = [int(v) for v in sys.argv[3].split(',')] if len(sys.argv) > 3 else None
w = W.ALL[name](B) if B else W.ALL[name]()
circ = w.circuit(Circuit)
e = Engine(circ, w.batch(), opts=w.opts)
table = {i: (NC, R, G) for i, NC, R, G in _abi.variants('real')}
r = {}
if name == 'c2':
    fn = lambda: e.op(want_resid=False); units = lambda: w.B
elif name == 'c3':
    e.set_jit('off')
    x0 = torch.as_tensor(common.c3_x0(e.op()['x'].cpu().numpy())).cuda()
    def fn():
        r['t'] = e.tran(x0, max_rows=256, **W.C3_TRAN)
    units = lambda: int(r['t']['rows'].sum())
elif name == 'c4':
    x0 = torch.as_tensor(common.c4_x0(circ, w.B)).cuda()
    def fn():
        r['t'] = e.tran(x0, max_rows=3072, **W.C4_TRAN)
    units = lambda: int(r['t']['rows'].sum())
for vid in [-1] + _abi.fitting_variants(e.n, 'real'):
    if only and vid not in only:
        continue
    if vid >= 0 and table[vid][0] == 0 and name != 'c2':
        continue
    for mode in ('off', 'on'):
        try:
            e.set_variant(real=vid); e.set_jit(mode)
            ms = timeit(fn, 2 if name != 'c2' else 5)
            print('%s B=%d variant %2d %-12s jit %-3s %10.4f ms  %.4g units/s' % (
                name, w.B, vid, table.get(vid, 'heuristic'), mode, ms, units() / ms * 1e3), flush=True)
        except EngineError as ex:
            print('%s variant %d jit %s: %s' % (name, vid, mode, str(ex)[:200]), flush=True)
# specialised kernels of any (rows per lane, lanes per instance) shape
e.set_variant(real=-1); e.set_jit('on')
n = e.n
for G in (1, 2, 4, 8, 16):
    R = (n + G - 1) // G
    if R > 16 or (only and -2 not in only):
        continue
    try:
        e.set_jit_shape(R, G)
        ms = timeit(fn, 2 if name != 'c2' else 5)
        print('%s B=%d shape R=%d G=%-2d        jit on  %10.4f ms  %.4g units/s' % (name, w.B, R, G, ms, units() / ms * 1e3), flush=True)
    except EngineError as ex:
        print('%s shape R=%d G=%d: %s' % (name, R, G, str(ex)[:300]), flush=True)
