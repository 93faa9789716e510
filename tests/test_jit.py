"""Circuit-specialised kernels (ahkab_b200/csrc/jit_gen.hpp, jit.hpp).

CPU: the generated constants + the engine's per-instance source compile for sm_100a with
NVRTC (no GPU needed), and the same specialised source built for the host gives results
bit-identical to the interpreted host build.  GPU: specialised == interpreted kernels on
the workloads, through the C-ABI."""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest
torch = pytest.importorskip("torch")

import common  # noqa: F401  (sys.path set-up)
from ahkab_b200 import _abi, workloads as W
from ahkab_b200.circuit import Circuit

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'hostemu'))


def _emu(w):
    import emu
    return emu.Emu(w.circuit(Circuit), w.batch(), w.opts)


def _nvrtc_available():
    for p in (os.environ.get('AHK_NVRTC', ''), 'libnvrtc.so.12',
              '/usr/local/cuda/lib64/libnvrtc.so.12'):
        if not p:
            continue
        try:
            C.CDLL(p)
            return True
        except OSError:
            pass
    return False


CASES = [('c2', W.c2_diode_op, 0, 'TRAP', 0, (2, 3, 4)),
         ('c3', W.c3_colpitts, 1, 'TRAP', 0, (7,)),
         ('c3op', W.c3_colpitts, 0, 'TRAP', 0, (7,)),
         ('c4', W.c4_ring3, 1, 'GEAR2', 1, (3, 5)),
         ('c4ie', W.c4_ring3, 1, 'IMPLICIT_EULER', 1, (5,)),
         ('c1ac', W.c1_butterworth, 2, 'TRAP', 0, (4, 5)),       # AC table: Var<16,1,16>, Var<24,1,32>
         ('c3ac', W.c3_colpitts, 2, 'TRAP', 0, (4,))]


@pytest.mark.skipif(not _nvrtc_available(), reason="NVRTC not installed")
@pytest.mark.parametrize('name,mk,tran,method,usc,vids', CASES, ids=[c[0] for c in CASES])
def test_specialised_source_compiles_for_sm_100a(name, mk, tran, method, usc, vids):
    import oracle
    lib = _abi.load_library()
    w = mk(B=16)
    o = oracle.Oracle(w.circuit(Circuit), w.batch(), w.opts)
    for vid in vids:
        nb = C.c_int64(0)
        rc = lib.ahk_jit_compile(C.byref(o.desc.c), C.byref(o.opts), len(o.slots),
                                 o.slots.ctypes.data_as(C.POINTER(C.c_int32)), tran,
                                 _abi.METHODS[method], usc, vid, C.byref(nb))
        assert rc == 0, lib.ahk_last_error().decode()[:2000]
        assert nb.value > 10000      # an sm_100a cubin came back


def _build_host_jit(spec, tag, defs=()):
    d = tempfile.mkdtemp(prefix='ahk_jit_%s_' % tag)
    open(os.path.join(d, 'jit_spec.h'), 'w').write(spec)
    so = os.path.join(d, 'libemuj.so')
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-I', d] + list(defs) + [
                           '-I', os.path.join(ROOT, 'ahkab_b200', 'csrc'), '-o', so,
                           os.path.join(ROOT, 'tests', 'hostemu', 'hostemu_jit.cpp'), '-lm'])
    return C.CDLL(so)


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def test_specialised_host_build_is_bit_identical_op():
    import emu
    w = W.c2_diode_op(B=256)
    e = _emu(w)
    emu.set_variant(1)                                    # Var<4,3,1>
    try:
        ref = e.op()
    finally:
        emu.set_variant(0)
    lib = _build_host_jit(e.jit_source(0, NC=4, R=3, G=1), 'c2')
    B, n = e.B, e.n
    x = np.zeros((B, n)); res = np.zeros((B, n))
    it = np.zeros(B, np.int32); st = np.zeros(B, np.int32)
    rc = lib.emuj_dc(C.c_int64(B), _dp(e.vary), C.c_int(-1), None, C.c_int(1), None, C.c_int(1),
                     _dp(x), _dp(res), _ip(it), _ip(st))
    assert rc == 0
    assert (st == ref['status']).all() and (it == ref['iters']).all()
    assert np.array_equal(x, ref['x']) and np.array_equal(res, ref['resid'])


@pytest.mark.parametrize('ilp,always', [(1, False), (3, False), (2, True)])
def test_specialised_host_build_is_bit_identical_tran(ilp, always):
    """ilp > 1: the devices go through the interleaved ekv_eval_v; always: every predicated
    section of the flat transient loop runs in every pass (AHK_EMU_ALWAYS), as on a warp whose
    other groups are in other phases — a badly predicated state update would change results."""
    import emu
    w = W.c4_ring3(B=2)
    e = _emu(w)
    tr = dict(W.C4_TRAN)
    rows_cap = 4096
    x0 = np.zeros((e.B, e.n))
    emu.set_variant(2)                                    # Var<6,5,1>
    try:
        ref = e.tran(x0, tr['tstart'], tr['tstep'], tr['tstop'] / 10, 'GEAR2', True, rows_cap)
    finally:
        emu.set_variant(0)
    lib = _build_host_jit(e.jit_source(1, method='GEAR2', use_step_control=True, NC=6, R=5, G=1, ilp=ilp),
                          'c4', ['-DAHK_EMU_ALWAYS'] if always else [])
    B, n = e.B, e.n
    t = np.full((B, rows_cap), np.nan); x = np.full((B, rows_cap, n), np.nan)
    rows = np.zeros(B, np.int32); cnt = np.zeros((B, 4), np.int32); st = np.zeros(B, np.int32)
    rc = lib.emuj_tran(C.c_int64(B), _dp(e.vary), _dp(x0), C.c_double(tr['tstart']),
                       C.c_double(tr['tstep']), C.c_double(tr['tstop'] / 10), C.c_int64(rows_cap),
                       _dp(t), _dp(x), _ip(rows), _ip(cnt), _ip(st))
    assert rc == 0
    assert (rows == ref['rows']).all() and (rows > 20).all()
    assert (st == ref['status']).all() and (cnt == ref['counters']).all()
    assert np.array_equal(t, ref['t'], equal_nan=True) and np.array_equal(x, ref['x'], equal_nan=True)


@pytest.mark.parametrize('seq,what', [('0 1 4 3 4', 'the sequence dgesv takes'),
                                      ('2 2 4 3 4', 'another valid sequence: every solve falls back'),
                                      ('2 2 4 3 4;0 1 4 3 4;0 2 4 3 4', 'a trie of three sequences: the code branches on the pivot'),
                                      ('4 1 2 3 4', 'a sequence with a structurally zero pivot: no schedule')])
def test_static_schedule_lu_host_build_is_bit_identical(seq, what, monkeypatch):
    """core.cuh slu_solve_thread / jit_gen.hpp slu_codegen: the LU eliminated symbolically along a
    pivot sequence, as straight-line code on the structural non-zeros, gives the bits of the
    pivoting one-lane LU — while the pivots agree because the skipped operations are on exact
    zeros, and otherwise because the solve is assembled again and goes through the pivoting LU."""
    import emu
    w = W.c4_ring3(B=2)
    e = _emu(w)
    tr = dict(W.C4_TRAN)
    rows_cap = 4096
    x0 = np.zeros((e.B, e.n))
    emu.set_variant(2)                                    # Var<6,5,1>
    try:
        ref = e.tran(x0, tr['tstart'], tr['tstep'], tr['tstop'] / 10, 'GEAR2', True, rows_cap)
    finally:
        emu.set_variant(0)
    plain = e.jit_source(1, method='GEAR2', use_step_control=True, NC=6, R=5, G=1)
    monkeypatch.setenv('AHK_SLU_SEQ', seq)
    spec = e.jit_source(1, method='GEAR2', use_step_control=True, NC=6, R=5, G=1)
    assert ('AHK_JIT_SLU_BODY' in spec) == (seq != '4 1 2 3 4'), what
    assert 'AHK_JIT_SLU_BODY' not in plain
    if 'AHK_JIT_SLU_BODY' not in spec:
        return
    lib = _build_host_jit(spec, 'c4slu')
    B, n = e.B, e.n
    t = np.full((B, rows_cap), np.nan); x = np.full((B, rows_cap, n), np.nan)
    rows = np.zeros(B, np.int32); cnt = np.zeros((B, 4), np.int32); st = np.zeros(B, np.int32)
    rc = lib.emuj_tran(C.c_int64(B), _dp(e.vary), _dp(x0), C.c_double(tr['tstart']),
                       C.c_double(tr['tstep']), C.c_double(tr['tstop'] / 10), C.c_int64(rows_cap),
                       _dp(t), _dp(x), _ip(rows), _ip(cnt), _ip(st))
    assert rc == 0
    assert (rows == ref['rows']).all() and (rows > 20).all()
    assert (st == ref['status']).all() and (cnt == ref['counters']).all()
    assert np.array_equal(t, ref['t'], equal_nan=True) and np.array_equal(x, ref['x'], equal_nan=True)


# ---- GPU: specialised kernels == interpreted kernels, through the C-ABI --------------------
def _engine(w):
    from ahkab_b200.engine import Engine
    return Engine(w.circuit(Circuit), w.batch(), opts=w.opts)


def _np(t):
    return t.detach().cpu().numpy()


def _same(a, b, what):
    """Same source, same compiler: expected bit-identical; rounding-level slack is allowed
    because constant folding may change which products are contracted into FMAs."""
    a, b = _np(a), _np(b)
    if a.dtype.kind == 'f':
        fin = np.isfinite(a) & np.isfinite(b)
        assert (np.isfinite(a) == np.isfinite(b)).all(), what
        assert np.abs(a[fin] - b[fin]).max(initial=0.0) <= 1e-9 * (1 + np.abs(b[fin]).max(initial=0.0)), what
    else:
        assert (a != b).mean() < 0.002, what


@pytest.mark.gpu
@pytest.mark.parametrize('vid', [-1, 2, 3, 4])
def test_gpu_specialised_op_and_sweep(vid):
    w = W.c2_diode_op(8192)
    e = _engine(w)
    e.set_variant(real=vid)
    n0 = e.jit_launch_count()
    e.set_jit('off'); a = e.op(); sa = e.dc_sweep('V1', np.linspace(0.2, 1.4, 7))
    assert e.jit_launch_count() == n0
    e.set_jit('on'); b = e.op(); sb = e.dc_sweep('V1', np.linspace(0.2, 1.4, 7))
    assert e.jit_launch_count() == n0 + 2          # the specialised kernels really ran
    for k in ('x', 'resid', 'iters', 'status'):
        _same(a[k], b[k], 'op ' + k)
    for k in ('x', 'iters', 'status'):
        _same(sa[k], sb[k], 'sweep ' + k)
    assert (_np(b['status']) & 1).all()
    # automatic policy: this batch is above the default threshold, a small one is not
    e.set_jit('auto'); e.op()
    assert e.jit_launch_count() == n0 + 3
    small = _engine(W.c2_diode_op(64)); small.op()
    assert small.jit_launch_count() == n0 + 3


@pytest.mark.gpu
@pytest.mark.parametrize('vid', [-1, 7, 9])
def test_gpu_specialised_fixed_step_transient(vid):
    w = W.c3_colpitts(256)
    e = _engine(w)
    e.set_variant(real=vid)
    e.set_jit('off'); opa = e.op()
    x0 = common.c3_x0(_np(opa['x']))
    ta = e.tran(x0, max_rows=256, **W.C3_TRAN)
    n0 = e.jit_launch_count()
    e.set_jit('on'); opb = e.op(); tb = e.tran(x0, max_rows=256, **W.C3_TRAN)
    assert e.jit_launch_count() == n0 + 2
    _same(opa['x'], opb['x'], 'op x')
    assert (_np(tb['rows']) == 250).all() and (_np(tb['status']) & 1).all()
    for k in ('rows', 'status', 'counters'):
        _same(ta[k], tb[k], 'tran ' + k)
    assert np.array_equal(_np(ta['t'])[:, :250], _np(tb['t'])[:, :250])
    xa, xb = _np(ta['x'])[:, :250], _np(tb['x'])[:, :250]
    assert np.abs(xa - xb).max() < 1e-6        # 250 steps of an oscillator amplify rounding


@pytest.mark.gpu
@pytest.mark.parametrize('vid', [-1, 3, 5])
def test_gpu_specialised_lte_transient(vid):
    w = W.c4_ring3(128)
    e = _engine(w)
    e.set_variant(real=vid)
    tr = dict(W.C4_TRAN); tr['tstop'] = 10e-9
    e.set_jit('off'); ta = e.tran(None, max_rows=2048, **tr)
    n0 = e.jit_launch_count()
    e.set_jit('on'); tb = e.tran(None, max_rows=2048, **tr)
    assert e.jit_launch_count() == n0 + 1
    ra, rb = _np(ta['rows']), _np(tb['rows'])
    assert (_np(tb['status']) & 1).all()
    # LTE step control amplifies last-bit differences: step counts agree within 2 %
    assert np.abs(ra - rb).max() <= np.maximum(3, 0.02 * ra).max()
    same = ra == rb
    assert same.mean() > 0.5
    xa, xb = _np(ta['x']), _np(tb['x'])
    for i in np.nonzero(same)[0][:16]:
        assert np.abs(xa[i, :ra[i]] - xb[i, :rb[i]]).max() < 1e-2


@pytest.mark.skipif(not _nvrtc_available(), reason="NVRTC not installed")
def test_disk_cache_skips_nvrtc(tmp_path, monkeypatch):
    """The cubin of a (circuit, options, variant) is stored under $AHK_JIT_CACHE and reused."""
    import oracle
    lib = _abi.load_library()
    monkeypatch.setenv('AHK_JIT_CACHE', str(tmp_path / 'cache'))
    w = W.c2_diode_op(B=16)
    o = oracle.Oracle(w.circuit(Circuit), w.batch(), w.opts)

    def compile_once():
        nb = C.c_int64(0)
        rc = lib.ahk_jit_compile(C.byref(o.desc.c), C.byref(o.opts), len(o.slots),
                                 o.slots.ctypes.data_as(C.POINTER(C.c_int32)), 0, 0, 0, 2, C.byref(nb))
        assert rc == 0, lib.ahk_last_error().decode()[:2000]
        return nb.value

    b0 = lib.ahk_jit_build_count()
    n1 = compile_once()
    assert lib.ahk_jit_build_count() == b0 + 1
    files = list((tmp_path / 'cache').glob('*.cubin'))
    assert len(files) == 1 and files[0].stat().st_size == n1
    n2 = compile_once()                                   # served from the file: no NVRTC run
    assert n2 == n1 and lib.ahk_jit_build_count() == b0 + 1
    monkeypatch.setenv('AHK_JIT_CACHE', 'off')
    compile_once()
    assert lib.ahk_jit_build_count() == b0 + 2


@pytest.mark.gpu
@pytest.mark.parametrize('vid', [-1, 4, 5])
def test_gpu_specialised_ac(vid):
    w = W.c1_butterworth(512)
    e = _engine(w)
    e.set_variant(ac=vid)
    om = common.c1_omegas()
    e.set_jit('off'); a = e.ac(om)
    n0 = e.jit_launch_count()
    e.set_jit('on'); b = e.ac(om)
    assert e.jit_launch_count() == n0 + 1
    xa, xb = _np(a['x']), _np(b['x'])
    assert (_np(b['status']) & 1).all() and (_np(a['status']) == _np(b['status'])).all()
    assert np.abs(xa - xb).max() <= 1e-10 * np.abs(xa).max()


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['c3', 'c4'])
def test_gpu_static_schedule_lu_matches_pivoting_lu(name):
    """Specialised transient kernels with the static-schedule LU (calibrated on the first call)
    against the same kernel shape with the pivoting LU: the scheduled elimination only skips
    operations on exact zeros, so rows, counters and values agree; the calibration must end on a
    schedule that instance 0 never leaves."""
    if name == 'c3':
        w = W.c3_colpitts(2048)
        e = _engine(w)
        e.set_jit('off')
        x0 = torch.as_tensor(common.c3_x0(_np(e.op()['x']))).cuda()
        run = lambda: e.tran(x0, max_rows=256, **W.C3_TRAN)
    else:
        w = W.c4_ring3(1024)
        e = _engine(w)
        x0 = torch.as_tensor(common.c4_x0(w.circuit(Circuit), w.B)).cuda()
        tr = dict(W.C4_TRAN); tr['tstop'] = tr['tstop'] / 5
        run = lambda: e.tran(x0, max_rows=1024, **tr)
    e.set_jit('on')
    e.set_jit_shape(9 if name == 'c3' else 5, 2)          # replicated rows, two lanes: the same shape for both
    e.set_slu(0); a = run(); sa = e.last_launch()
    e.set_slu(1); b = run(); sb = e.last_launch()
    st = e.slu_stats()
    assert sa['jit'] == 1 and sb['jit'] == 1 and (sa['R'], sa['G']) == (sb['R'], sb['G']), (sa, sb)
    assert st['calibrations'] >= 1 and st['last_solves'] > 100 and st['last_off'] == 0, st
    b2 = run()                                            # calibrated: no further calibration launch
    assert e.slu_stats()['calibrations'] == st['calibrations']
    for r in (b, b2):
        assert (_np(a['rows']) == _np(r['rows'])).all() and (_np(a['status']) == _np(r['status'])).all()
        assert (_np(a['counters']) == _np(r['counters'])).all()
        rows = _np(a['rows'])
        valid = np.arange(_np(a['t']).shape[1])[None, :] < rows[:, None]      # rows each instance wrote
        assert np.array_equal(_np(a['t'])[valid], _np(r['t'])[valid])
        xa, xr = _np(a['x'])[valid], _np(r['x'])[valid]
        assert np.abs(xa - xr).max() <= 1e-9 * (1 + np.abs(xa).max())


@pytest.mark.gpu
def test_gpu_static_schedule_lu_instances_that_pivot_differently_fall_back():
    """C2's column 0 holds 1 / R1 next to the voltage source's 1: which one dgesv takes flips with
    the instance's R1, so a schedule learned from instance 0 fits only part of the batch (and
    the homotopy stages of the OP change it too).  With the static-schedule LU forced onto the
    OP kernel (mode 2), every solve that leaves it goes through the pivoting LU: same results
    as the kernel without a schedule, bit for bit (the scheduled form only skips exact zeros)."""
    w = W.c2_diode_op(8192)
    e = _engine(w)
    e.set_jit('on')
    e.set_variant(real=2)                                 # Var<4,3,1>: one lane per instance
    e.set_slu(0); a = e.op()
    e.set_slu(2); b = e.op()
    assert e.slu_stats()['calibrations'] >= 1
    assert e.last_launch()['jit'] == 1
    assert (_np(a['status']) == _np(b['status'])).all() and (_np(a['iters']) == _np(b['iters'])).all()
    assert np.array_equal(_np(a['x']), _np(b['x']))


@pytest.mark.gpu
def test_gpu_ac_host_zero_copy_matches_ac():
    """Engine.ac_host: the lane-per-chunk AC kernel reading its parameters from and writing its
    result rows into pinned host memory gives what the device-buffer call gives, bit for bit
    (ragged batch: the last block is partial)."""
    w = W.c1_butterworth(301)
    e = _engine(w)
    e.set_jit('on')
    om = common.c1_omegas()
    a = e.ac(om)
    h = e.ac_host(om)
    torch.cuda.synchronize()
    assert e.last_launch()['jit'] == 1 and e.last_launch()['R'] == e.n       # the static-schedule kernel
    assert (h['status'].numpy() == _np(a['status'])).all() and (h['status'].numpy() & 1).all()
    assert np.array_equal(h['x'].numpy(), _np(a['x']))


def test_generated_constants_do_not_depend_on_the_batch_values():
    """The text that keys the compiled-kernel cache must be the same for two batches that
    vary the same parameters with different values (else every run() call recompiles)."""
    def spec(seed):
        w = W.c2_diode_op(B=32, seed=seed)
        return _emu(w).jit_source(0, NC=4, R=3, G=1)
    a, b = spec(1), spec(2)
    assert a == b
    w = W.c2_diode_op(B=32)
    assert _emu(w).jit_source(0, NC=4, R=3, G=1) != _emu(w).jit_source(0, NC=4, R=1, G=4)
