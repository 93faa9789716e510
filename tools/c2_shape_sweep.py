"""C2 (OP) specialised-kernel shape vs batch size: python tools/c2_shape_sweep.py"""
import sys
import torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from ahkab_b200 import workloads as W
from ahkab_b200.circuit import Circuit
from ahkab_b200.engine import Engine, EngineError
flush = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')
for B in (2048, 8192, 16384, 32768, 65536):
    w = W.c2_diode_op(B)
    circ = w.circuit(Circuit)
    for shape in ((0, 0), (3, 1), (1, 4), (2, 2)):
        try:
            e = Engine(circ, w.batch(), opts=w.opts)
            e.set_jit('on')
            e.set_jit_shape(*shape)
            e.op(); torch.cuda.synchronize()
            ts = []
            for _ in range(7):
                flush.zero_()
                a, b = torch.cuda.Event(True), torch.cuda.Event(True)
                a.record(); e.op(); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ts.sort()
            print('c2 B=%d shape R,G=%s -> %s  median %.4f ms' % (B, shape, e.last_launch(), ts[len(ts) // 2]), flush=True)
        except EngineError as ex:
            print('c2 B=%d shape %s: %s' % (B, shape, str(ex)[:200]), flush=True)
