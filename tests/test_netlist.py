"""Netlist front-end reuse: circuits parsed by the REFERENCE's own netlist parser
(netlist_parser.py:145, :1428) converted with ahkab_b200.netlist.from_reference, then run through
the oracle (CPU) / the CUDA engine (-m gpu) and compared with what the reference itself computes
for the parsed circuit.  Skipped where the reference package is absent (it is pure Python:
baseline/_ref travels to the GPU box)."""
import os
import sys

import numpy as np
import pytest

import common
import ahkab_b200
from ahkab_b200 import utilities

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', 'oracle'))
NETLISTS = ('diode_bridge', 'lc_filter', 'ekv_stage')


def _reference_run(name):
    import ref_loop
    if ref_loop.reference_path() is None:
        pytest.skip("reference package not on this machine")
    ref = ref_loop._load()
    from ahkab import netlist_parser
    rc, directives, _pp = netlist_parser.parse_circuit(os.path.join(HERE, 'netlists', name + '.ckt'))
    ans = netlist_parser.parse_analysis(rc, directives)
    for a in ans:
        a['verbose'] = 0
        a['outfile'] = None
    ans = [ref.new_op(verbose=0) if a['type'] == 'op' else a for a in ans]
    for a in ans:                      # (parse_analysis leaves outfile to the CLI: temp files like new_*)
        if a.get('outfile') is None and a['type'] != 'op':
            import tempfile
            a['outfile'] = os.path.join(tempfile.mkdtemp(prefix='ahk_nl_'), name + '.' + a['type'])
    res = ref.run(rc, ans)
    return rc, ans, res


def _sim_results(sim, circ, ans):
    out = {}
    op = sim.op()
    out['op'] = op
    for a in ans:
        if a['type'] == 'dc':
            sw = np.array(list(utilities.lin_axis_iterator(a['start'], a['stop'],
                                                           (a['stop'] - a['start']) / a['step'] + 1)))
            src = [i for i, e in enumerate(circ) if e.part_id.lower() == a['source'].lower()][0]
            out['dc'] = sim.dc_sweep(src, sw)
        elif a['type'] == 'ac':
            om = np.array(list(utilities.log_axis_iterator(2 * np.pi * a['start'], 2 * np.pi * a['stop'],
                                                           a['points'])))
            out['ac'] = sim.ac(om, x_op=op['x'] if circ.is_nonlinear() else None)
        elif a['type'] == 'tran':
            # x0 None: the transient starts from zeros (transient.py:189-191); 'op': from the OP
            x0 = op['x'] if a.get('x0') is not None else np.zeros_like(op['x'])
            out['tran'] = sim.tran(x0, a['tstart'], a['tstep'], a['tstop'], a['method'].upper(),
                                   a.get('use_step_control', True), 8192)
    return out


def _compare(circ, ref_res, got):
    nv1 = circ.get_nodes_number() - 1
    assert common.parity_tol(got['op']['x'][0], ref_res['op'].asarray().ravel(), nv1) < 1e-3
    if 'dc' in got:
        a = ref_res['dc'].asarray()
        assert common.parity_tol(got['dc']['x'][0], a[1:].T, nv1) < 1.0
    if 'ac' in got:
        a = ref_res['ac'].asarray()
        g = a[1:].T
        scale = np.abs(g).max(axis=0, keepdims=True)
        assert (np.abs(got['ac']['x'][0] - g) / (1e-12 + 1e-8 * scale)).max() < 1.0
    if 'tran' in got:
        a = ref_res['tran'].asarray()
        tg, xg = a[0], a[1:].T
        rows = int(got['tran']['rows'][0])
        assert abs(rows - len(tg)) <= max(2, 0.02 * len(tg))
        t, x = got['tran']['t'][0, :rows], got['tran']['x'][0, :rows]
        span = np.abs(xg).max(axis=0) + 1e-12
        for v in range(x.shape[1]):
            xi = np.interp(tg, t, x[:, v])
            assert np.abs(xi - xg[:, v]).max() <= 2e-2 * span[v] + (1e-6 if v < nv1 else 1e-9), v


@pytest.mark.parametrize('name', NETLISTS)
def test_parsed_netlists_through_the_oracle_match_the_reference(name):
    import oracle
    rc, ans, ref_res = _reference_run(name)
    circ = ahkab_b200.netlist.from_reference(rc)
    # same unknowns, same order, same names
    assert circ.get_nodes_number() == rc.get_nodes_number()
    assert [e.part_id for e in circ] == [e.part_id for e in rc]
    mine = ahkab_b200.netlist.convert_analyses(ans)
    assert [a['type'] for a in mine] == [a['type'] for a in ans]
    _compare(circ, ref_res, _sim_results(oracle.Oracle(circ), circ, ans))


def test_converted_models_are_frozen():
    rc, _ans, _res = _reference_run('ekv_stage')
    circ = ahkab_b200.netlist.from_reference(rc)
    pb = ahkab_b200.ParamBatch(4)
    pb.set('rd', 'value', np.array([9e3, 1e4, 1.1e4, 1.2e4]))
    from ahkab_b200 import compile as cp
    flat = cp.flatten(circ, pb)                    # element values vary ...
    assert flat.vary.shape == (1, 4)
    pb.set('nch', 'VTO', np.full(4, .45))
    with pytest.raises(KeyError):                  # ... derived model parameters do not
        cp.flatten(circ, pb)


@pytest.mark.gpu
@pytest.mark.parametrize('name', NETLISTS)
def test_parsed_netlists_on_the_gpu_match_the_reference(name):
    from test_zoo import _Gpu
    rc, ans, ref_res = _reference_run(name)
    circ = ahkab_b200.netlist.from_reference(rc)
    for jit in ('off', 'on'):
        _compare(circ, ref_res, _sim_results(_Gpu(circ, jit=jit), circ, ans))
