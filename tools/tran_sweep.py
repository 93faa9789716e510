"""Times the specialised transient kernel of a workload over (rows/lane, lanes, ilp, threads, min_blocks)
shapes: python tools/tran_sweep.py c4 [B] "R,G,ilp,threads,minb;..."  (CUDA events, best of 2)"""
import sys
import torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import common  # noqa
from ahkab_b200 import workloads as W
from ahkab_b200.circuit import Circuit
from ahkab_b200.engine import Engine, EngineError

name = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else None
shapes = [tuple(int(v) for v in q.split(',')) for q in sys.argv[3].split(';')]
w = W.ALL[name](B) if B else W.ALL[name]()
circ = w.circuit(Circuit)
e = Engine(circ, w.batch(), opts=w.opts)
r = {}
if name == 'c3':
    e.set_jit('off')
    x0 = torch.as_tensor(common.c3_x0(e.op()['x'].cpu().numpy())).cuda()
    fn = lambda: r.__setitem__('t', e.tran(x0, max_rows=256, **W.C3_TRAN))
else:
    x0 = torch.as_tensor(common.c4_x0(circ, w.B)).cuda()
    fn = lambda: r.__setitem__('t', e.tran(x0, max_rows=3072, **W.C4_TRAN))
e.set_jit('on')
ref = None
for R, G, ilp, thr, minb in shapes:
    try:
        e.set_jit_shape(R, G); e.set_jit_tuning(ilp, thr, minb)
        fn(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(2):
            a, b = torch.cuda.Event(True), torch.cuda.Event(True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        t = r['t']
        rows = int(t['rows'].sum()); ok = int((t['status'] & 1).sum())
        cnt = t['counters'].sum(0).tolist()
        if ref is None:
            ref = t['rows'].clone()
        drow = int((t['rows'] - ref).abs().max())
        print('%s B=%d R=%d G=%d ilp=%d thr=%d minb=%d  %9.3f ms  %.4g steps/s  ok %d/%d rows %d cnt %s max|drows| %d' % (
            name, w.B, R, G, ilp, thr, minb, best, rows / best * 1e3, ok, w.B, rows, cnt, drow), flush=True)
    except EngineError as ex:
        print('%s R=%d G=%d ilp=%d thr=%d minb=%d: %s' % (name, R, G, ilp, thr, minb, str(ex)[:400]), flush=True)
