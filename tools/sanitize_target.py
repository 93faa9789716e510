"""Small runs of every specialised kernel family for compute-sanitizer:
compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import sys
import torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import common  # noqa
from ahkab_b200 import workloads as W
from ahkab_b200.circuit import Circuit
from ahkab_b200.engine import Engine

# C2: one-lane OP kernel with cooperative result rows, ragged last block, zero-copy host path
w = W.c2_diode_op(1000)
e = Engine(w.circuit(Circuit), w.batch(), opts=w.opts); e.set_jit('on')
r = e.op(); h = e.op_host(nchunks=0); torch.cuda.synchronize()
print('c2', int((r['status'] & 1).sum()), int((h['status'] & 1).sum()))
# C3: static-schedule transient (replicated rows), calibration
w = W.c3_colpitts(70)
e = Engine(w.circuit(Circuit), w.batch(), opts=w.opts); e.set_jit('on')
x0 = torch.as_tensor(common.c3_x0(e.op()['x'].cpu().numpy())).cuda()
tr = dict(W.C3_TRAN); tr['tstop'] = tr['tstop'] / 10
r = e.tran(x0, max_rows=64, **tr); torch.cuda.synchronize()
print('c3', r['rows'][:4].tolist(), e.slu_stats())
# C4: LTE transient
w = W.c4_ring3(40)
circ = w.circuit(Circuit)
e = Engine(circ, w.batch(), opts=w.opts); e.set_jit('on')
x0 = torch.as_tensor(common.c4_x0(circ, w.B)).cuda()
tr = dict(W.C4_TRAN); tr['tstop'] = tr['tstop'] / 20
r = e.tran(x0, max_rows=512, **tr); torch.cuda.synchronize()
print('c4', r['rows'][:4].tolist(), e.slu_stats())
# C1: lane-per-chunk AC, ragged batch
w = W.c1_butterworth(37)
e = Engine(w.circuit(Circuit), w.batch(), opts=w.opts); e.set_jit('on')
r = e.ac(common.c1_omegas()); torch.cuda.synchronize()
print('c1', int((r['status'] & 1).sum()), e.slu_stats(), e.last_launch())
# C5: compiled schedule, batch not a multiple of 32
w = W.c5_r2r(37)
e = Engine(w.circuit(Circuit), w.batch(), opts=w.opts); e.set_jit('on')
r = e.dc_sweep('V1', common.c5_sweep()); torch.cuda.synchronize()
print('c5', int((r['status'] & 1).sum()))
