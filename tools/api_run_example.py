import sys, time
sys.path.insert(0, '.')
import numpy as np
import ahkab_b200 as ahkab
rng = np.random.default_rng(0)
def once():
    cir = ahkab.Circuit('test diode')
    cir.add_model('diode', 'di', dict(IS=1e-14, N=1.0))
    cir.add_vsource('v1', '1', '0', dc_value=1.0)
    cir.add_resistor('r1', '1', '2', 1.0)
    cir.add_diode('d1', '2', '0', 'di')
    batch = ahkab.ParamBatch(65536)
    batch.set('di', 'IS', 1e-14 * 10 ** rng.uniform(-1, 1, 65536))
    t0 = time.time()
    res = ahkab.run(cir, ahkab.new_op(), batch=batch)
    v = res['op']['V2']
    return time.time() - t0, v.shape, float(v.mean())
for i in range(3):
    print(once())
from ahkab_b200 import _abi
print('nvrtc builds', _abi.load_library().ahk_jit_build_count(), 'jit launches', _abi.load_library().ahk_jit_launch_count())
