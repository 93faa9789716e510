import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from ahkab_b200 import workloads as W
from ahkab_b200.circuit import Circuit
from ahkab_b200.engine import Engine
w = W.c2_diode_op(65536)
e = Engine(w.circuit(Circuit), w.batch(), opts=w.opts)
flush = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')
for _ in range(5): e.op()
torch.cuda.synchronize()
ts = []
for _ in range(50):
    flush.zero_()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record(); e.op(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ts.sort()
print('median %.4f ms  min %.4f  mean %.4f' % (ts[25], ts[0], sum(ts)/50))
