"""Per-call wall time of the drop-in API on the C3 transient: python tools/api_tran_timing.py"""
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
import ahkab_b200 as ahkab
from ahkab_b200 import workloads as W
from ahkab_b200.circuit import Circuit
w = W.c3_colpitts(16384)
for k, v in w.opts.items():
    setattr(ahkab.options, k, v)
circ = w.circuit(Circuit); pb = w.batch()
out = None
for i in range(12):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = ahkab.run(circ, [ahkab.new_op(), ahkab.new_tran(x0='op+ic', **W.C3_TRAN)], batch=pb)['tran']
    t1 = time.perf_counter()
    t, x, offs = r.packed()
    t2 = time.perf_counter()
    out = (t, x, r.rows, r.status)
    del r
    torch.cuda.synchronize(); t3 = time.perf_counter()
    eng = ahkab.get_engine(circ, pb)
    print('call %d: run %.2f ms, packed %.2f ms, total %.2f ms, pool %d' % (i, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t0) * 1e3, len(eng.__dict__.get('_host_pool', []))), flush=True)
