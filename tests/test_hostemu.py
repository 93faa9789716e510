"""CPU: the CUDA engine's per-instance source (core.cuh / entry.cuh / ac.cuh /
devices.cuh) compiled for the host with one lane per instance, against the
oracle on the same seeded inputs.  Covers the generic shared-memory variant and
the register-resident thread-per-instance variants (the multi-lane variants need
warp shuffles and are covered by the -m gpu tests)."""
import numpy as np
import pytest

import common
import emu
import oracle
from ahkab_b200 import workloads as W
from ahkab_b200.circuit import Circuit


def _pair(w, variant):
    emu.set_variant(variant)
    circ = w.circuit(Circuit)
    return (emu.Emu(circ, w.batch(), opts=w.opts),
            oracle.Oracle(circ, w.batch(), opts=w.opts), circ)


@pytest.fixture(autouse=True)
def _reset_variant():
    yield
    emu.set_variant(0)


@pytest.mark.parametrize('variant', [0, 1, 2])
def test_c2_diode_op(variant):
    e, o, _ = _pair(W.c2_diode_op(2048), variant)
    r, q = e.op(), o.op()
    assert (r['status'] == q['status']).all() and (r['status'] & 1).all()
    assert np.abs(r['x'] - q['x']).max() < 1e-12
    # Gauss-Jordan vs LU rounding may flip a convergence test on a rare instance
    assert (r['iters'] != q['iters']).mean() < 0.002


@pytest.mark.parametrize('variant', [0])
def test_c3_colpitts_op_tran(variant):
    e, o, _ = _pair(W.c3_colpitts(8), variant)
    rop, qop = e.op(), o.op()
    assert np.abs(rop['iters'] - qop['iters']).max() <= 2
    assert common.parity_tol(rop['x'], qop['x'], 5, 1e-6, 1e-6) < 1e-3
    x0 = common.c3_x0(qop['x'])
    a = (x0, 0., W.C3_TRAN['tstep'], W.C3_TRAN['tstop'], 'TRAP', False, 256)
    rt, qt = e.tran(*a), o.tran(*a)
    assert (rt['rows'] == 250).all() and (qt['rows'] == 250).all()
    assert np.array_equal(rt['t'][:, :250], qt['t'][:, :250])
    assert common.parity_tol(rt['x'][:, :250], qt['x'][:, :250], 5, 1e-6, 1e-6) < 1.0
    assert np.abs(rt['counters'][:, 1] - qt['counters'][:, 1]).max() <= 8


@pytest.mark.parametrize('variant', [0, 2])
def test_c4_ring3_gear2_lte(variant):
    w = W.c4_ring3(4)
    e, o, circ = _pair(w, variant)
    x0 = common.c4_x0(circ, w.B)
    a = (x0, 0., 1e-9, 50e-9, 'GEAR2', True, 4096)
    rt, qt = e.tran(*a), o.tran(*a)
    assert (rt['status'] & 1).all()
    assert np.abs(rt['rows'] - qt['rows']).max() <= 0.01 * qt['rows'].max()
    k = 200     # early window, before LTE-controlled grids drift apart
    assert common.parity_tol(rt['x'][:, :k], qt['x'][:, :k], 4) < 50.0


@pytest.mark.parametrize('variant', [0])
def test_c1_butterworth_ac(variant):
    e, o, _ = _pair(W.c1_butterworth(8), variant)
    om = common.c1_omegas()
    r, q = e.ac(om), o.ac(om)
    scale = np.abs(q['x']).max(axis=(1,), keepdims=True)
    assert (np.abs(r['x'] - q['x']) / (1e-9 + 1e-6 * scale)).max() < 1.0


@pytest.mark.parametrize('variant', [0, 1])
def test_small_rc_ac_register_variant(variant):
    """n = 3 linear RC low-pass: the AC register variant Var<4,3,1> on the host."""
    def build(Circuit, p):
        cir = Circuit('rc')
        cir.add_vsource('V1', 'in', cir.gnd, dc_value=0., ac_value=1.)
        cir.add_resistor('R1', 'in', 'out', p['R1'])
        cir.add_capacitor('C1', 'out', cir.gnd, p['C1'])
        return cir
    rng = np.random.default_rng(0)
    B = 64
    params = dict(R1=1e3 * (1 + .1 * rng.standard_normal(B)),
                  C1=1e-6 * (1 + .1 * rng.standard_normal(B)))
    w = W.Workload('rc', build, dict(R1=1e3, C1=1e-6), params,
                   dict(R1=('R1', 'value'), C1=('C1', 'value')))
    e, o, _ = _pair(w, variant)
    om = 2 * np.pi * np.logspace(1, 4, 20)
    r, q = e.ac(om), o.ac(om)
    assert np.abs(r['x'] - q['x']).max() < 1e-12
    # analytic: |Vout| = 1/sqrt(1 + (w R C)^2) (Gmin = 1e-12 S on the nodes shifts it by ~R*Gmin)
    ana = 1 / np.sqrt(1 + (om[None, :] * (params['R1'] * params['C1'])[:, None]) ** 2)
    assert np.abs(np.abs(r['x'][:, :, 1]) - ana).max() < 1e-8


def test_c5_small_ladder_dc_sweep():
    w = W.c5_r2r(8, nodes=30)
    e, o, _ = _pair(w, 0)
    sw = common.c5_sweep()
    r, q = e.dc_sweep(0, sw), o.dc_sweep(0, sw)
    assert np.abs(r['x'] - q['x']).max() / np.abs(q['x']).max() < 1e-13


def test_variant_that_does_not_fit_is_refused():
    emu.set_variant(1)
    w = W.c3_colpitts(2)          # n = 9 > 3
    circ = w.circuit(Circuit)
    with pytest.raises(ValueError):
        emu.Emu(circ, w.batch(), opts=w.opts).op()


def test_static_schedule_path_for_large_linear_circuits():
    """biglin3.cuh on the host: the schedule built from the nominal instance reproduces the
    oracle's sweep of the 257-unknown ladder, and an instance with other pivots is flagged
    for the dynamic kernel instead of being solved wrongly."""
    import oracle
    w = W.c5_r2r(16)
    w.params['R1h'] = w.params['R1h'].copy()
    w.params['R1h'][5] = 0.5                # g = 2 S > the source row's 1: this instance pivots differently
    circ = w.circuit(Circuit)
    sw = common.c5_sweep()
    o = oracle.Oracle(circ, w.batch(), opts=w.opts).dc_sweep(0, sw)
    r = emu.Emu(circ, w.batch(), w.opts).big3_sweep(0, sw)
    assert r['dynamic'] >= 0                # a schedule was built
    served = r['flags'] == 0
    assert served.sum() == 15 and r['flags'][5] == 2 and r['dynamic'] == 1
    scale = np.abs(o['x']).max()
    assert np.abs(r['x'][served] - o['x'][served]).max() / scale < 1e-14
    assert (r['status'][served] == o['status'][served]).all() and (r['iters'][served] == 2).all()


class _Big3CK(__import__('ctypes').Structure):
    import ctypes as _C
    _fields_ = [('vary', _C.c_void_p), ('B', _C.c_longlong), ('src_elem', _C.c_int), ('sweep', _C.c_void_p),
                ('npts', _C.c_int), ('pt0', _C.c_int), ('ptc', _C.c_int), ('x_out', _C.c_void_p),
                ('iters', _C.c_void_p), ('status', _C.c_void_p), ('resid', _C.c_void_p), ('scr', _C.c_void_p), ('scr_stride', _C.c_longlong),
                ('F', _C.c_void_p),
                ('Mv', _C.c_void_p), ('flagm', _C.c_void_p), ('ovf', _C.c_void_p)]


@pytest.mark.parametrize('nodes', [256, 40])
def test_compiled_schedule_host_build_matches_interpreted(nodes):
    """biglin3_gen.hpp: the static elimination schedule emitted as straight-line source (what
    NVRTC compiles for the GPU), built here for the host, gives the bits of the interpreted
    schedule — incl. the hand-over of an instance whose pivots differ from the nominal ones."""
    import ctypes as C
    import subprocess
    import tempfile
    w = W.c5_r2r(6, nodes=nodes)
    w.params['R1h'] = w.params['R1h'].copy()
    w.params['R1h'][3] = 0.5                              # other pivot in column 0 -> dynamic kernel
    e = emu.Emu(w.circuit(Circuit), w.batch(), w.opts)
    lib = emu.lib()
    c6 = e._common()
    nf, npos = C.c_int32(0), C.c_int32(0)
    args = (c6[0], c6[1], c6[3], c6[4])
    n = lib.emu_big3_gen_source(*args, None, C.c_int64(0), C.byref(nf), C.byref(npos))
    assert n > 0
    buf = C.create_string_buffer(n + 1)
    lib.emu_big3_gen_source(*args, buf, C.c_int64(n + 1), C.byref(nf), C.byref(npos))
    d = tempfile.mkdtemp(prefix='ahk_b3gen_')
    open(d + '/b3.cpp', 'wb').write(buf.value)
    subprocess.check_call(['g++', '-O1', '-std=c++17', '-fPIC', '-shared', '-DAHK_B3_HOST',
                           '-o', d + '/b3.so', d + '/b3.cpp'])
    so = C.CDLL(d + '/b3.so')
    sw = common.c5_sweep()
    B, nn, P = e.B, e.n, len(sw)
    x = np.zeros((B, P, nn)); res = np.zeros((B, P, nn))
    it = np.zeros((B, P), np.int32); st = np.zeros((B, P), np.int32)
    B32 = (B + 31) // 32 * 32
    F = np.zeros(2 * nf.value * B32); Mv = np.zeros(npos.value * B32)   # tiled by 32 lanes
    fl = np.zeros((2, B), np.int32); ovf = np.zeros(B + 1, np.int32)
    vary = np.ascontiguousarray(e.vary)
    scr = np.zeros(2 * nn * 32)
    k = _Big3CK(vary.ctypes.data, B, 0, sw.ctypes.data, P, 0, P, x.ctypes.data, it.ctypes.data,
                st.ctypes.data, res.ctypes.data, scr.ctypes.data, 1, F.ctypes.data, Mv.ctypes.data, fl.ctypes.data,
                ovf.ctypes.data)
    so.b3_host_run(C.byref(k))
    ref = e.big3_sweep(0, sw)
    ok = ref['flags'] == 0
    assert ok.sum() == B - 1 and (fl[0] == ref['flags']).all()
    assert ovf[0] == 1 and ovf[1] == 3                    # handed to the dynamic kernel
    for key, got in (('x', x), ('resid', res), ('iters', it), ('status', st)):
        assert np.array_equal(got[ok], ref[key][ok]), key
