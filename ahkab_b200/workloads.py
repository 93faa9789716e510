"""The five BASELINE.json workloads as (circuit builder, synthetic batch).

Each workload is written against the `Circuit.add_*` API only, so the same
`build(Circuit, p)` function constructs either an `ahkab_b200.Circuit` (one
topology + a `ParamBatch`) or — in `oracle/make_golden.py`, in the authoring
container — one reference `ahkab.Circuit` per instance.  Inputs follow
SURVEY.md §8(d): numpy `default_rng(seed)`, float64, Gaussian factors clipped
to +-3 sigma, nominal values from the cited reference files."""
import numpy as np

from .batch import ParamBatch


def _gauss(rng, B, rel):
    return 1.0 + rel * np.clip(rng.standard_normal(B), -3.0, 3.0)


class Workload(object):
    """name, build(Circuit, p), params{name: (B,) array}, keys{name: (owner,
    attr)}, nominal{name: float}, options overrides, description."""

    def __init__(self, name, build, nominal, params, keys, opts=None,
                 descr=""):
        self.name, self.build = name, build
        self.nominal, self.params, self.keys = nominal, params, keys
        self.opts = opts or {}
        self.descr = descr
        self.B = len(next(iter(params.values()))) if params else 1

    def circuit(self, Circuit):
        return self.build(Circuit, self.nominal)

    def instance_params(self, i):
        p = dict(self.nominal)
        for k, v in self.params.items():
            p[k] = float(v[i])
        return p

    def batch(self, lo=0, hi=None):
        hi = self.B if hi is None else hi
        pb = ParamBatch(hi - lo)
        for k, v in self.params.items():
            owner, attr = self.keys[k]
            pb.set(owner, attr, v[lo:hi])
        return pb


# ---- C1: README 5th-order Butterworth band-pass, AC ------------------------
def c1_butterworth(B=4096, seed=0):
    """/root/reference/README.md:61-80; new_ac(.97e3, 1.03e3, 1e2, x0=None)."""
    nominal = dict(R1=50., R2=50., L1=0.245894, L2=9.83652e-05, L3=0.795775,
                   L4=9.83652e-05, L5=0.245894, C1=1.03013e-07,
                   C2=0.000257513, C3=3.1831e-08, C4=0.000257513,
                   C5=1.03013e-07)

    def build(Circuit, p):
        cir = Circuit('Butterworth 1kHz band-pass filter')
        cir.add_vsource('V1', 'n1', cir.gnd, dc_value=0., ac_value=1.)
        cir.add_resistor('R1', 'n1', 'n2', p['R1'])
        cir.add_inductor('L1', 'n2', 'n3', p['L1'])
        cir.add_capacitor('C1', 'n3', 'n4', p['C1'])
        cir.add_inductor('L2', 'n4', cir.gnd, p['L2'])
        cir.add_capacitor('C2', 'n4', cir.gnd, p['C2'])
        cir.add_inductor('L3', 'n4', 'n5', p['L3'])
        cir.add_capacitor('C3', 'n5', 'n6', p['C3'])
        cir.add_inductor('L4', 'n6', cir.gnd, p['L4'])
        cir.add_capacitor('C4', 'n6', cir.gnd, p['C4'])
        cir.add_capacitor('C5', 'n7', 'n8', p['C5'])
        cir.add_inductor('L5', 'n6', 'n7', p['L5'])
        cir.add_resistor('R2', 'n8', cir.gnd, p['R2'])
        return cir

    rng = np.random.default_rng(seed)
    params = {k: nominal[k] * _gauss(rng, B, 0.05) for k in nominal}
    keys = {k: (k, 'value') for k in nominal}
    return Workload('c1_butterworth_ac', build, nominal, params, keys,
                    descr="README Butterworth AC, 100 log pts .97k-1.03k")


C1_AC = dict(start=.97e3, stop=1.03e3, points=1e2)


# ---- C2: tests/diode_op ------------------------------------------------------
def c2_diode_op(B=65536, seed=1):
    """/root/reference/tests/diode_op/diode_op.ckt; new_op()."""
    nominal = dict(IS=1e-14, N=1.0, R1=1.0)

    def build(Circuit, p):
        cir = Circuit('test diode')
        cir.add_model('diode', 'di', dict(IS=p['IS'], N=p['N']))
        cir.add_vsource('v1', '1', '0', dc_value=1.0)
        cir.add_resistor('r1', '1', '2', p['R1'])
        cir.add_diode('d1', '2', '0', 'di')
        return cir

    rng = np.random.default_rng(seed)
    params = dict(IS=1e-14 * 10.0 ** rng.uniform(-1, 1, B),
                  N=rng.uniform(1, 2, B),
                  R1=1.0 * _gauss(rng, B, 0.05))
    keys = dict(IS=('di', 'IS'), N=('di', 'N'), R1=('r1', 'value'))
    return Workload('c2_diode_op', build, nominal, params, keys,
                    descr="tests/diode_op OP")


# ---- C3: tests/colpitts ------------------------------------------------------
def c3_colpitts(B=16384, seed=2):
    """/root/reference/tests/colpitts/colpitts.ckt with vea=iea=1e-6
    (tests/colpitts/test_colpitts.py:8-10); OP, then TRAP fixed step
    0.1 ns to 25 ns from the ic-modified OP (uic=2)."""
    nominal = dict(L1=5e-9, C1=1.12e-12, C2=1.12e-12, R0=3.5e3, IB=1.3e-3,
                   VTO=.4, KP=10e-6)

    def build(Circuit, p):
        cir = Circuit('MOS COLPITTS OSCILLATOR')
        cir.add_model('ekv', 'nmos', dict(TYPE='n', VTO=p['VTO'], KP=p['KP']))
        cir.add_vsource('vdd', 'dd', '0', dc_value=2.5)
        cir.add_inductor('l1', 'dd', 'nd', p['L1'], ic=-1e-9)
        cir.add_resistor('r0', 'nd', 'dd', p['R0'])
        cir.add_capacitor('c1', 'nd', 'ns', p['C1'], ic=2.5)
        cir.add_capacitor('c2', 'ns', '0', p['C2'])
        cir.add_mos('m1', 'nd1', 'bias', 'ns', 'ns', w=200e-6, l=1e-6,
                    model_label='nmos')
        cir.add_vsource('vtest', 'nd', 'nd1', dc_value=0.)
        cir.add_vsource('vbias', 'bias', '0', dc_value=2.)
        cir.add_isource('ib', 'ns', '0', dc_value=p['IB'])
        return cir

    rng = np.random.default_rng(seed)
    params = dict(L1=5e-9 * _gauss(rng, B, 0.05),
                  C1=1.12e-12 * _gauss(rng, B, 0.05),
                  C2=1.12e-12 * _gauss(rng, B, 0.05),
                  R0=3.5e3 * _gauss(rng, B, 0.05),
                  IB=1.3e-3 * _gauss(rng, B, 0.02),
                  VTO=.4 + 0.01 * np.clip(rng.standard_normal(B), -3, 3),
                  KP=10e-6 * _gauss(rng, B, 0.05))
    keys = dict(L1=('l1', 'value'), C1=('c1', 'value'), C2=('c2', 'value'),
                R0=('r0', 'value'), IB=('ib', 'dc_value'),
                VTO=('nmos', 'VTO'), KP=('nmos', 'KP'))
    return Workload('c3_colpitts_tran', build, nominal, params, keys,
                    opts=dict(vea=1e-6, iea=1e-6),
                    descr="tests/colpitts OP + TRAP fixed 0.1n x 250")


C3_TRAN = dict(tstart=0., tstop=25e-9, tstep=.1e-9, method='TRAP',
               use_step_control=False)


# ---- C4: tests/ring3 ---------------------------------------------------------
def c4_ring3(B=16384, seed=3):
    """/root/reference/tests/ring3/ring3.ckt (subcircuits expanded in the
    reference's order: X1, X2, X3 then v2); x0 from the .ic directive;
    GEAR2 with LTE step control, tstop 50 ns, HMAX 1 ns."""
    nominal = dict(NVTO=.4, NKP=40e-6, PVTO=-.4, PKP=12e-6, CA=10e-15,
                   CB=10e-15, CC=10e-15)

    def build(Circuit, p):
        cir = Circuit('RING OSCILLATOR with 3 inverters')
        cir.add_model('ekv', 'nch', dict(TYPE='n', VTO=p['NVTO'],
                                         KP=p['NKP']))
        cir.add_model('ekv', 'pch', dict(TYPE='p', VTO=p['PVTO'],
                                         KP=p['PKP']))
        for k, (i, o, cap) in enumerate((('1', '3', 'CA'), ('3', '4', 'CB'),
                                         ('4', '1', 'CC'))):
            cir.add_mos('m1x%d' % (k + 1), o, i, '2', '2', w=3e-6, l=1e-6,
                        model_label='pch')
            cir.add_mos('m2x%d' % (k + 1), o, i, '0', '0', w=1e-6, l=1e-6,
                        model_label='nch')
            cir.add_capacitor('c1x%d' % (k + 1), o, '0', p[cap])
        cir.add_vsource('v2', '2', '0', dc_value=2.5)
        return cir

    rng = np.random.default_rng(seed)
    params = dict(NVTO=.4 + 0.01 * np.clip(rng.standard_normal(B), -3, 3),
                  NKP=40e-6 * _gauss(rng, B, 0.05),
                  PVTO=-.4 + 0.01 * np.clip(rng.standard_normal(B), -3, 3),
                  PKP=12e-6 * _gauss(rng, B, 0.05),
                  CA=10e-15 * _gauss(rng, B, 0.05),
                  CB=10e-15 * _gauss(rng, B, 0.05),
                  CC=10e-15 * _gauss(rng, B, 0.05))
    keys = dict(NVTO=('nch', 'VTO'), NKP=('nch', 'KP'), PVTO=('pch', 'VTO'),
                PKP=('pch', 'KP'), CA=('c1x1', 'value'),
                CB=('c1x2', 'value'), CC=('c1x3', 'value'))
    return Workload('c4_ring3_tran', build, nominal, params, keys,
                    descr="tests/ring3 GEAR2 + LTE, 50 ns")


C4_IC = {'v(1)': 0, 'v(2)': 2.5, 'v(3)': 2.5, 'v(4)': 0}
C4_TRAN = dict(tstart=0., tstop=50e-9, tstep=1e-9, method='GEAR2',
               use_step_control=True)


# ---- C5: tests/r2r ladder, 256 nodes ----------------------------------------
def c5_r2r(B=4096, seed=4, nodes=256):
    """R-2R ladder as generated by tests/r2r/test_r2r.py:71-94;
    new_dc(1e3, 10e3, 10, 'V1')."""
    names = []
    nominal = {}
    for n in range(1, nodes):
        names += ['R%dh' % n, 'R%dv' % n]
        nominal['R%dh' % n] = 1.3e3
        nominal['R%dv' % n] = 2.6e3
    names.append('R%dve' % nodes)
    nominal['R%dve' % nodes] = 2.6e3

    def build(Circuit, p):
        cir = Circuit('MATRIX SIZE TEST: R-2R ladder')
        cir.add_vsource('V1', '1', '0', dc_value=10e3)
        for n in range(1, nodes):
            cir.add_resistor('R%dh' % n, str(n), str(n + 1), p['R%dh' % n])
            cir.add_resistor('R%dv' % n, str(n + 1), '0', p['R%dv' % n])
        cir.add_resistor('R%dve' % nodes, str(nodes), '0',
                         p['R%dve' % nodes])
        return cir

    rng = np.random.default_rng(seed)
    params = {k: nominal[k] * _gauss(rng, B, 0.01) for k in names}
    keys = {k: (k, 'value') for k in names}
    return Workload('c5_r2r_dc', build, nominal, params, keys,
                    descr="R-2R %d nodes, DC sweep V1 1k..10k, 10 pts"
                    % nodes)


C5_DC = dict(start=1e3, stop=10e3, points=10, source='V1')

ALL = dict(c1=c1_butterworth, c2=c2_diode_op, c3=c3_colpitts, c4=c4_ring3,
           c5=c5_r2r)
