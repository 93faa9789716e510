"""Short single-workload run for ncu captures: python tools/ncu_target.py c2 [variant] [reps] [B]"""
import os
import sys
import torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import common  # noqa
from ahkab_b200 import workloads as W
from ahkab_b200.circuit import Circuit
from ahkab_b200.engine import Engine

name = sys.argv[1] if len(sys.argv) > 1 else 'c2'
G = int(sys.argv[2]) if len(sys.argv) > 2 else -1   # kernel variant id, -1 = heuristic
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
B = int(sys.argv[4]) if len(sys.argv) > 4 else None
w = W.ALL[name](B) if B else W.ALL[name]()
circ = w.circuit(Circuit)
e = Engine(circ, w.batch(), opts=w.opts)
e.set_variant(real=G if name != 'c1' else -1, ac=G if name == 'c1' else -1)
e.set_jit(os.environ.get('AHK_JIT', 'auto'))   # circuit-specialised kernels: auto / on / off
if os.environ.get('AHK_SHAPE'):                # "R,G,ilp,threads,minb" of the specialised kernel
    R_, G_, ilp_, thr_, minb_ = (int(v) for v in os.environ['AHK_SHAPE'].split(','))
    e.set_jit_shape(R_, G_); e.set_jit_tuning(ilp_, thr_, minb_)
for _ in range(reps):
    if name == 'c2':
        r = e.op()
    elif name == 'c1':
        r = e.ac(common.c1_omegas())
    elif name == 'c3':
        x0 = torch.as_tensor(common.c3_x0(e.op()['x'].cpu().numpy())).cuda()
        r = e.tran(x0, max_rows=256, **W.C3_TRAN)
    elif name == 'c4':
        x0 = torch.as_tensor(common.c4_x0(circ, w.B)).cuda()
        r = e.tran(x0, max_rows=3072, **W.C4_TRAN)
    elif name == 'c5':
        r = e.dc_sweep('V1', common.c5_sweep())
torch.cuda.synchronize()
print('done', name, {k: tuple(v.shape) for k, v in r.items() if hasattr(v, 'shape')})
if 'counters' in r:
    print('rows', int(r['rows'].sum()), 'solves/nr/rej/cut', r['counters'].sum(0).tolist())
