"""C1 end to end: packed copies through the pipeline against Engine.ac_host (zero-copy):
python tools/ac_zero_copy.py [steps]"""
import sys
import torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import common  # noqa
from ahkab_b200 import workloads as W
from ahkab_b200.circuit import Circuit
from ahkab_b200.engine import Engine
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
w = W.c1_butterworth()
e = Engine(w.circuit(Circuit), w.batch(), opts=w.opts)
om = common.c1_omegas()
vary_host = torch.empty(e.vary.shape, dtype=torch.float64, pin_memory=True); vary_host.copy_(e.vary)
torch.cuda.synchronize()
flush = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')


def timeit(fn):
    fn(); fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def copies():
    e.set_vary(vary_host)
    return e.to_host(e.ac(om))


print('c1 kernel only        min %.3f median %.3f ms' % timeit(lambda: e.ac(om)))
print('c1 packed copies e2e  min %.3f median %.3f ms' % timeit(copies))
print('c1 zero-copy e2e      min %.3f median %.3f ms' % timeit(lambda: e.ac_host(om, vary_host=vary_host)))
