"""Edge cases of the batched path: ragged / single-instance batches, the smallest
circuit, singular matrices, homotopy fall-backs, row-buffer overflow, step-size
underflow and argument errors.  CPU: host build of the engine vs the oracle (status
words included); -m gpu: the CUDA engine vs the oracle."""
import numpy as np
import pytest

import common
import oracle
from ahkab_b200 import _abi
from ahkab_b200 import time_functions as TF
from ahkab_b200 import workloads as W
from ahkab_b200.batch import ParamBatch
from ahkab_b200.circuit import Circuit


# ---- circuits -----------------------------------------------------------------------
def one_node():
    """n = 1: a current source into a resistor."""
    c = Circuit('one node')
    c.add_isource('I1', c.gnd, 'a', dc_value=2e-3)
    c.add_resistor('R1', 'a', c.gnd, 500.)
    return c


def floating_node():
    """Two capacitors in series: the middle node floats at DC — regular with Gmin,
    singular without it (pass 2 of the OP fails, the Gmin solution is kept)."""
    c = Circuit('floating')
    c.add_vsource('V1', 'a', c.gnd, dc_value=1.)
    c.add_capacitor('C1', 'a', 'm', 1e-9)
    c.add_capacitor('C2', 'm', c.gnd, 1e-9)
    return c


def step_source():
    """A 1 ns edge into an RC: with options.hmin raised to 20 ns the LTE control asks for a
    step below hmin — the reference raises ValueError('Step size too small')
    (transient.py:467-469), the batch reports AHK_ST_HMIN for the instance."""
    c = Circuit('edge')
    c.add_vsource('V1', 'in', c.gnd, dc_value=0.,
                  function=TF.pulse(0., 1., 1e-6, 1e-9, 2e-6, 1e-9, 5e-6))
    c.add_resistor('R1', 'in', 'out', 1e3)
    c.add_capacitor('C1', 'out', c.gnd, 1e-9)
    return c


class _Gpu(object):
    def __init__(self, circ, batch=None, opts=None, jit='auto'):
        from ahkab_b200.engine import Engine
        self.e = Engine(circ, batch, opts=opts)
        self.e.set_jit(jit)

    @staticmethod
    def _np(r):
        import torch
        return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else v) for k, v in r.items()
                if k != '_packed'}

    def op(self, **kw):
        return self._np(self.e.op(**kw))

    def tran(self, x0, tstart, tstep, tstop, method, usc, max_rows):
        return self._np(self.e.tran(x0, tstart, tstep, tstop, method=method,
                                    use_step_control=usc, max_rows=max_rows))


def _sims(circ, batch=None, opts=None, gpu=False):
    o = oracle.Oracle(circ, batch, opts=opts)
    if gpu:
        return _Gpu(circ, batch, opts, jit='on' if gpu == 'jit' else 'off'), o
    import emu
    emu.set_variant(0)
    return emu.Emu(circ, batch, opts=opts), o


BACKENDS = [pytest.param(False, id='host-build'), pytest.param(True, id='cuda', marks=pytest.mark.gpu),
            pytest.param('jit', id='cuda-specialised', marks=pytest.mark.gpu)]


@pytest.mark.parametrize('gpu', BACKENDS)
@pytest.mark.parametrize('B', [1, 2, 33, 129, 1000])
def test_ragged_batch_sizes(gpu, B):
    w = W.c2_diode_op(B)
    s, o = _sims(w.circuit(Circuit), w.batch(), gpu=gpu)
    r, q = s.op(), o.op()
    assert r['x'].shape == (B, 3)
    assert np.abs(r['x'] - q['x']).max() < 1e-12 and (r['status'] == q['status']).all()


@pytest.mark.parametrize('gpu', BACKENDS)
def test_smallest_circuit_and_no_batch(gpu):
    s, o = _sims(one_node(), gpu=gpu)            # no ParamBatch: one nominal instance
    r, q = s.op(), o.op()
    assert r['x'].shape == (1, 1) and abs(r['x'][0, 0] - 1.0) < 1e-9     # 2 mA * 500 Ohm
    assert np.array_equal(r['status'], q['status']) and int(r['iters'][0]) == int(q['iters'][0]) == 2


@pytest.mark.parametrize('gpu', BACKENDS)
def test_singular_pass2_keeps_the_gmin_solution(gpu):
    s, o = _sims(floating_node(), gpu=gpu)
    r, q = s.op(), o.op()
    st = int(r['status'][0])
    assert st == int(q['status'][0])
    # the no-Gmin pass meets an exactly singular matrix (LinAlgError in the reference,
    # dc_analysis.py:279-288) and falls back to gmin stepping, which ends at 1e-10 S
    assert st & _abi.ST_OK and st & _abi.ST_SINGULAR
    assert (st & _abi.ST_METHOD_MASK) >> _abi.ST_METHOD_SHIFT == 1
    assert np.allclose(r['x'], q['x'], rtol=1e-9, atol=1e-12)
    assert abs(r['x'][0, 1]) < 1e-9                # only Gmin ties the floating node: 0 V


@pytest.mark.parametrize('gpu', BACKENDS)
def test_homotopy_fallbacks_match_the_oracle(gpu):
    """dc_max_nr_iter = 5 makes the standard solve of tests/diode_op fail (it needs 11
    iterations): dc_solve moves on to gmin stepping, then source stepping
    (dc_analysis.py:128-320, 354-408).  Most instances converge through the homotopy
    (20-35 iterations in total instead of 12), a few exhaust every method: status without
    AHK_ST_OK, method bits = source stepping, x = NaN (the reference returns None)."""
    w = W.c2_diode_op(256)
    s, o = _sims(w.circuit(Circuit), w.batch(), opts=dict(dc_max_nr_iter=5), gpu=gpu)
    r, q = s.op(), o.op()
    assert (r['status'] == q['status']).all()
    assert (r['iters'] == q['iters']).all() and (r['iters'] > 12).any()
    ok = (r['status'] & 1) != 0
    assert ok.any() and (~ok).any()
    assert np.allclose(r['x'][ok], q['x'][ok], rtol=1e-9, atol=1e-12)
    assert (((r['status'][~ok] & _abi.ST_METHOD_MASK) >> _abi.ST_METHOD_SHIFT) == 2).all()
    assert np.isnan(r['x'][~ok]).all() and np.isnan(q['x'][~ok]).all()


@pytest.mark.parametrize('gpu', BACKENDS)
def test_row_buffer_overflow_is_reported(gpu):
    w = W.c3_colpitts(4)
    s, o = _sims(w.circuit(Circuit), w.batch(), opts=w.opts, gpu=gpu)
    x0 = common.c3_x0(o.op()['x'])
    a = (x0, 0., W.C3_TRAN['tstep'], W.C3_TRAN['tstop'], 'TRAP', False, 100)   # needs 250 rows
    r, q = s.tran(*a), o.tran(*a)
    assert (r['rows'] == 100).all() and (q['rows'] == 100).all()
    assert ((r['status'] & _abi.ST_ROWS_FULL) != 0).all() and ((r['status'] & 1) == 0).all()
    assert (r['status'] == q['status']).all()


@pytest.mark.parametrize('gpu', BACKENDS)
def test_step_size_underflow_sets_hmin(gpu):
    s, o = _sims(step_source(), opts=dict(hmin=2e-8), gpu=gpu)
    x0 = np.zeros((1, 3))
    a = (x0, 0., 1e-7, 5e-6, 'TRAP', True, 8192)
    r, q = s.tran(*a), o.tran(*a)
    assert int(r['status'][0]) & _abi.ST_HMIN and not int(r['status'][0]) & 1
    assert int(q['status'][0]) == int(r['status'][0]) and int(q['rows'][0]) == int(r['rows'][0])
    # the very first LTE-controlled step, (tstop - tstart) / 9999 = 0.5 ns, is already
    # below hmin: both stop after the bootstrap step
    assert int(r['rows'][0]) <= 2


@pytest.mark.gpu
def test_argument_errors_raise():
    from ahkab_b200.engine import Engine, EngineError
    w = W.c2_diode_op(8)
    e = Engine(w.circuit(Circuit), w.batch())
    with pytest.raises(EngineError):
        e.tran(None, 1.0, 1e-3, 0.5)              # tstart > tstop (transient.py:153-155)
    with pytest.raises(EngineError):
        e.tran(None, 0.0, -1e-3, 0.5)             # tstep < 0 (transient.py:156-158)
    with pytest.raises(ValueError):
        e.tran(None, 0.0, 1e-3, 0.5, method='RK4')
    with pytest.raises(EngineError):
        e.dc_sweep('r1', np.array([1.0, 2.0]))    # only V / I sources can be swept
    with pytest.raises(EngineError):
        e.dc_sweep('v1', np.array([]))
    with pytest.raises(EngineError):
        e.set_variant(real=99)                    # no such kernel variant
    e.set_variant(real=-1)
    w9 = W.c3_colpitts(2)
    e9 = Engine(w9.circuit(Circuit), w9.batch())
    with pytest.raises(EngineError):
        e9.set_variant(real=2)                    # Var<4,3,1> cannot hold n = 9


@pytest.mark.gpu
def test_near_tie_pivots_of_the_multi_lane_solve():
    """The multi-lane Gauss-Jordan compares pivot candidates through a key that treats
    magnitudes equal to 2^-47 as ties (lowest row wins), where LAPACK's idamax takes the
    exact maximum: here column 0 holds 1 - 2^-50 in rows 0, 1 and exactly 1 in the KVL row
    2 (pass 2, no Gmin), so the two rules pick DIFFERENT pivot rows.  Both are valid
    pivots: the solutions must agree to rounding."""
    import oracle
    from ahkab_b200 import _abi
    from ahkab_b200.circuit import Circuit
    from ahkab_b200.engine import Engine
    c = Circuit('near tie')
    c.add_vsource('V1', 'a', c.gnd, dc_value=1.)
    c.add_resistor('R1', 'a', 'b', 1.0 + 2.0 ** -50)       # 1/R1 = 1 - 2^-50
    c.add_resistor('R2', 'b', c.gnd, 1.0)
    assert 1.0 / (1.0 + 2.0 ** -50) == 1.0 - 2.0 ** -50
    o = oracle.Oracle(c).op()
    table = {i: (NC, R, G) for i, NC, R, G in _abi.variants('real')}
    multi = [i for i in _abi.fitting_variants(3, 'real') if table[i][2] > 1]
    assert multi
    for v in [-1] + multi:
        for jit in ('off', 'on'):
            e = Engine(c)
            e.set_variant(real=v)
            e.set_jit(jit)
            r = e.op()
            x = r['x'].cpu().numpy()
            assert (r['status'].cpu().numpy() & 1).all()
            assert np.abs(x - o['x']).max() <= 1e-14, (v, jit, x, o['x'])
