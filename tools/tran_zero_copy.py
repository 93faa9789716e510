"""Experiment: transient results written by the kernel straight into pinned host buffers (zero-copy,
as ahk_op_host with nchunks = 0) against the packed device->host copy of the bench's end-to-end
step: python tools/tran_zero_copy.py c3|c4 [steps]"""
import ctypes as C
import sys
import time
import torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import common  # noqa
from ahkab_b200 import _abi, workloads as W
from ahkab_b200.circuit import Circuit
from ahkab_b200.engine import Engine, _ptr

name = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
w = W.ALL[name]()
circ = w.circuit(Circuit)
e = Engine(circ, w.batch(), opts=w.opts)
B, n = e.B, e.n
if name == 'c3':
    x0 = torch.as_tensor(common.c3_x0(e.op()['x'].cpu().numpy())).cuda()
    tr, max_rows = dict(W.C3_TRAN), 256
else:
    x0 = torch.as_tensor(common.c4_x0(circ, w.B)).cuda()
    tr, max_rows = dict(W.C4_TRAN), 3072
method = tr.pop('method', 'TRAP'); usc = tr.pop('use_step_control', True)
vary_host = torch.empty(e.vary.shape, dtype=torch.float64, pin_memory=True); vary_host.copy_(e.vary)
x0_host = torch.empty(x0.shape, dtype=torch.float64, pin_memory=True); x0_host.copy_(x0)
torch.cuda.synchronize()


def ev():
    return torch.cuda.Event(enable_timing=True)


def timeit(fn):
    fn(); fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        a, b = ev(), ev()
        a.record(); r = fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts), r


def packed():
    e.set_vary(vary_host)
    x0.copy_(x0_host, non_blocking=True)
    raw = e.tran(x0, max_rows=max_rows, method=method, use_step_control=usc, **tr)
    return e.to_host(e.pack_tran(raw))


t = torch.empty((B, max_rows), dtype=torch.float64, pin_memory=True)
x = torch.empty((B, max_rows, n), dtype=torch.float64, pin_memory=True)
rows = torch.empty((B,), dtype=torch.int32, pin_memory=True)
cnt = torch.empty((B, 4), dtype=torch.int32, pin_memory=True)
st = torch.empty((B,), dtype=torch.int32, pin_memory=True)


def zero_copy():
    b = _abi.ahk_batch()
    b.B = B; b.n_vary = len(e._slots)
    b.vary_slots = e._slots.ctypes.data_as(C.POINTER(C.c_int32))
    b.vary = vary_host.data_ptr()
    rc = e.lib.ahk_tran(e._h, C.byref(b), _ptr(x0_host), float(tr['tstart']), float(tr['tstep']), float(tr['tstop']),
                        _abi.METHODS[method], int(bool(usc)), int(max_rows), _ptr(t), _ptr(x), _ptr(rows), _ptr(cnt),
                        _ptr(st), e._stream())
    assert rc == 0, e._err()
    return dict(t=t, x=x, rows=rows)


a = timeit(packed)
print(name, 'packed copy e2e: min %.3f ms avg %.3f ms' % a[:2])
b = timeit(zero_copy)
print(name, 'zero-copy e2e:   min %.3f ms avg %.3f ms' % b[:2], 'rows', int(rows.sum()))
