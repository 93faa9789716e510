"""Dev probe: times every BASELINE workload at full size for each
kernel variant (CUDA events, inputs resident).  Not the bench."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import common  # noqa
from ahkab_b200 import workloads as W
from ahkab_b200.circuit import Circuit
from ahkab_b200.engine import Engine


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    which = sys.argv[1:] or ['c2', 'c1', 'c3', 'c4', 'c5']
    print(torch.cuda.get_device_name(0))
    for name in which:
        t0 = time.time()
        w = W.ALL[name]()
        circ = w.circuit(Circuit)
        e = Engine(circ, w.batch(), opts=w.opts)
        print('== %s B=%d n=%d (setup %.2fs)' % (name, w.B, e.n, time.time() - t0), flush=True)
        from ahkab_b200 import _abi
        kind = 'ac' if name == 'c1' else 'real'
        Gs = [-1] + _abi.fitting_variants(e.n, kind)
        if name == 'c5':
            Gs = [-1]
        info = {i: (NC, R, G_) for i, NC, R, G_ in _abi.variants(kind)}
        for G in Gs:
            e.set_variant(real=G if kind == 'real' else -1, ac=G if kind == 'ac' else -1)
            try:
                if name == 'c2':
                    ms = timeit(lambda: e.op())
                    units = w.B
                elif name == 'c1':
                    om = common.c1_omegas()
                    ms = timeit(lambda: e.ac(om))
                    units = w.B * len(om)
                elif name == 'c3':
                    x0 = torch.as_tensor(common.c3_x0(e.op()['x'].cpu().numpy())).cuda()
                    r = {}
                    def f():
                        r['t'] = e.tran(x0, max_rows=256, **W.C3_TRAN)
                    ms = timeit(f)
                    units = int(r['t']['rows'].sum())
                elif name == 'c4':
                    x0 = torch.as_tensor(common.c4_x0(circ, w.B)).cuda()
                    r = {}
                    def f():
                        r['t'] = e.tran(x0, max_rows=3072, **W.C4_TRAN)
                    ms = timeit(f, reps=1)
                    units = int(r['t']['rows'].sum())
                    print('   rows min/mean/max', int(r['t']['rows'].min()), float(r['t']['rows'].float().mean()), int(r['t']['rows'].max()),
                          'status ok', int((r['t']['status'] & 1).sum()))
                elif name == 'c5':
                    sw = common.c5_sweep()
                    ms = timeit(lambda: e.dc_sweep('V1', sw), reps=2)
                    units = w.B * len(sw)
                print('   variant %2d %-12s %9.3f ms  %.4g units/s' % (G, info.get(G, 'auto'), ms, units / ms * 1e3), flush=True)
            except Exception as ex:
                print('   variant %2d  failed: %s' % (G, ex), flush=True)


if __name__ == '__main__':
    main()
