"""C2 kernel time per step, L2 flushed: distribution over 200 steps, with and without a
spin kernel ahead of the start event (which lets the CPU run ahead of the GPU, so that the
interval between the two events cannot contain CPU launch latency)."""
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
for spin in (0, 400000):
    ts = []
    for _ in range(200):
        flush.zero_()
        if spin:
            torch.cuda._sleep(spin)
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); e.op(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print('spin %6d: median %.4f ms  min %.4f  mean %.4f  p90 %.4f  top5 %s' % (
        spin, ts[100], ts[0], sum(ts) / 200, ts[180], ['%.3f' % t for t in ts[-5:]]))
