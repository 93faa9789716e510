"""The general large-n engine (64 < n <= 512; dc_analysis.py:796-822 solves any size densely):
non-linear circuits beyond the shared-memory engine, against the oracle on seeded batches — the
host build of the engine on CPU, the CUDA engine with -m gpu.  (tests/test_zoo.py holds the
same circuits at n = 102 / 74 against the LIVE reference, all four analyses.)"""
import numpy as np
import pytest

import common
import zoo
from ahkab_b200 import time_functions as TF
from ahkab_b200.batch import ParamBatch
from ahkab_b200.circuit import Circuit


def _batch(circ, B, seed=0, k=8):
    pb = ParamBatch(B)
    rng = np.random.default_rng(seed)
    for e in [e for e in circ if e.part_id.lower().startswith('r')][:k]:
        pb.set(e.part_id, 'value', e.value * (1 + 0.05 * rng.standard_normal(B)))
    return pb


CASES = [('ladder', zoo.diode_ladder, 250), ('mesh', zoo.diode_mesh, 150), ('mesh', zoo.diode_mesh, 83)]


@pytest.mark.parametrize('name,build,N', CASES)
def test_host_build_matches_oracle_on_large_circuits(name, build, N):
    import emu
    import oracle
    emu.set_variant(0)
    circ = build(Circuit, TF, N)
    pb = _batch(circ, 3)
    o = oracle.Oracle(circ, pb).op()
    e = emu.Emu(circ, pb).op()
    assert (o['status'] & 1).all() and np.array_equal(e['status'], o['status'])
    assert np.array_equal(e['iters'], o['iters'])
    nv1 = circ.get_nodes_number() - 1
    assert common.parity_tol(e['x'], o['x'], nv1, rel=1e-9) < 1.0


@pytest.mark.gpu
@pytest.mark.parametrize('name,build,N', CASES)
def test_cuda_engine_matches_oracle_on_large_circuits(name, build, N):
    """OP on a seeded batch: sparse ladder (bitmap-tracked elimination) and dense meshes (blocked
    panels, FP64 tensor-core trailing update), one warp per instance, matrix in global memory."""
    import oracle
    from ahkab_b200.engine import Engine
    circ = build(Circuit, TF, N)
    B = 40
    pb = _batch(circ, B)
    o = oracle.Oracle(circ, pb).op()
    eng = Engine(circ, pb)
    r = eng.op()
    x = r['x'].cpu().numpy()
    st = r['status'].cpu().numpy()
    assert eng.n > 64 and eng.last_launch()['G'] == 32 and eng.last_launch()['NC'] == 0
    assert (st & 1).all() and np.array_equal(st, o['status'])
    assert (r['iters'].cpu().numpy() != o['iters']).mean() <= 0.05
    nv1 = circ.get_nodes_number() - 1
    assert common.parity_tol(x, o['x'], nv1, rel=1e-6) < 1.0


@pytest.mark.gpu
def test_large_linear_circuit_through_the_general_engine_matches_the_linear_path():
    """C5 (linear, n = 257) forced through the general large-n engine == the factor-once path."""
    from ahkab_b200 import workloads as W
    from ahkab_b200.engine import Engine
    w = W.c5_r2r(B=16)
    circ = w.circuit(Circuit)
    sw = common.c5_sweep()[:3]
    a = Engine(circ, w.batch(), opts=w.opts)
    ra = a.dc_sweep('V1', sw)
    b = Engine(circ, w.batch(), opts=w.opts)
    b.set_variant(real=1)
    rb = b.dc_sweep('V1', sw)
    assert b.last_launch()['G'] == 32
    xa, xb = ra['x'].cpu().numpy(), rb['x'].cpu().numpy()
    assert np.abs(xa - xb).max() <= 1e-9 * np.abs(xa).max()
    assert (rb['status'].cpu().numpy() & 1).all()


@pytest.mark.gpu
def test_large_transient_matches_oracle():
    import oracle
    from ahkab_b200.engine import Engine
    circ = zoo.diode_ladder(Circuit, TF, 120)
    B = 6
    pb = _batch(circ, B)
    orc = oracle.Oracle(circ, pb)
    x0 = orc.op()['x']
    ot = orc.tran(x0, 0., 1e-7, 4e-6, 'TRAP', True, 512)
    eng = Engine(circ, pb)
    rt = eng.tran(x0, 0., 1e-7, 4e-6, method='TRAP', use_step_control=True, max_rows=512)
    rows, orows = rt['rows'].cpu().numpy(), ot['rows']
    assert (np.abs(rows - orows) <= np.maximum(2, 0.02 * orows)).all()
    t, x = rt['t'].cpu().numpy(), rt['x'].cpu().numpy()
    nv1 = circ.get_nodes_number() - 1
    for b in range(B):
        tg = ot['t'][b, :orows[b]]
        for v in (0, 1, nv1 // 2, nv1 - 1):
            xi = np.interp(tg, t[b, :rows[b]], x[b, :rows[b], v])
            ref = ot['x'][b, :orows[b], v]
            assert np.abs(xi - ref).max() <= 1e-6 + 2e-3 * np.abs(ref).max()
