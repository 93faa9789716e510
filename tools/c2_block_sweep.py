"""C2 (OP, one lane per instance) specialised kernel: block size / register cap at the BASELINE batch:
python tools/c2_block_sweep.py"""
import sys
import torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from ahkab_b200 import workloads as W
from ahkab_b200.circuit import Circuit
from ahkab_b200.engine import Engine, EngineError
flush = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')
w = W.c2_diode_op(65536)
circ = w.circuit(Circuit)
for thr, minb in ((0, 0), (32, 16), (64, 8), (64, 4), (96, 5), (128, 4), (128, 3), (128, 2), (256, 2), (192, 2), (160, 3)):
    try:
        e = Engine(circ, w.batch(), opts=w.opts)
        e.set_jit('on'); e.set_jit_shape(3, 1); e.set_jit_tuning(0, thr, minb)
        e.op(); torch.cuda.synchronize()
        ts = []
        for _ in range(15):
            flush.zero_()
            a, b = torch.cuda.Event(True), torch.cuda.Event(True)
            a.record(); e.op(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        print('c2 threads=%d minb=%d -> %s  median %.4f ms min %.4f' % (thr, minb, e.last_launch(), ts[len(ts) // 2], ts[0]), flush=True)
    except EngineError as ex:
        print('c2 threads=%d minb=%d: %s' % (thr, minb, str(ex)[:200]), flush=True)
