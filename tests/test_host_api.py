"""CPU: host-side mirror of the reference API — Circuit construction, the
topology compiler (flatten), analysis descriptors, sweep iterators, source
waveforms, x0 helpers and the results objects (fed with numpy stand-ins for
the device tensors).  No GPU, no compute through the C-ABI."""
import ctypes as C

import numpy as np
import pytest

import ahkab_b200 as ahkab
import oracle
from ahkab_b200 import compile as cp
from ahkab_b200 import results, utilities, time_functions as TF
from ahkab_b200 import workloads as W
from ahkab_b200.circuit import Circuit


def test_new_analysis_descriptors_keep_reference_signatures():
    op = ahkab.new_op()
    assert op == {'type': 'op', 'guess': True, 'x0': None, 'outfile': None, 'verbose': 0}
    dc = ahkab.new_dc(1e3, 10e3, 10, 'V1')
    assert dc['step'] == (10e3 - 1e3) / 9 and dc['sweep_type'] == 'LINEAR'   # ahkab.py:317
    tr = ahkab.new_tran(0, 25e-9, .1e-9, method='TRAP', use_step_control=False)
    assert tr['x0'] == 'op' and tr['tstep'] == .1e-9
    ac = ahkab.new_ac(.97e3, 1.03e3, 1e2, x0=None)
    assert ac['sweep_type'] == 'LOG' and ac['points'] == 1e2


def test_sweep_iterators_are_cumulative_like_the_reference():
    """utilities.py:246-378: products / sums, so f[99] = 1029.9999999999918, not 1030."""
    # the AC sweep iterates over omega = 2 pi f and reports omega / 2 / pi (ac.py:283-309)
    om = np.array(list(utilities.log_axis_iterator(2 * np.pi * .97e3, 2 * np.pi * 1.03e3, 1e2)))
    f = om / np.pi / 2
    assert f[0] == 970 and f[1] == 970.58823353488754 and f[99] == 1029.9999999999918
    s = np.array(list(utilities.lin_axis_iterator(1e3, 10e3, 10)))
    assert np.allclose(s, np.arange(1, 11) * 1e3, rtol=1e-15) and len(s) == 10


def test_node_numbering_and_variable_names_follow_creation_order():
    w = W.c4_ring3(2)
    circ = w.circuit(Circuit)
    # internal numbering by creation order: '3','1','2','4' 
    assert [circ.nodes_dict[i] for i in range(1, 5)] == ['3', '1', '2', '4']
    raw = dict(x=np.zeros((2, 5)), resid=np.zeros((2, 5)), iters=np.zeros(2, np.int32),
               status=np.ones(2, np.int32))
    sol = results.op_solution(circ, raw)
    assert sol.keys() == ['V3', 'V1', 'V2', 'V4', 'I(V2)']
    assert 'v1' in sol and sol['i(v2)'].shape == (2,)
    c3 = W.c3_colpitts(2).circuit(Circuit)
    names = results.tran_solution(c3, dict(t=np.zeros((2, 4)), x=np.zeros((2, 4, 9)),
                                           rows=np.array([4, 2]), status=np.ones(2, np.int32)),
                                  0, 1, 'TRAP').keys()
    assert names == ['T', 'VDD', 'VND', 'VNS', 'VND1', 'VBIAS', 'I(VDD)', 'I(L1)', 'I(VTEST)',
                     'I(VBIAS)']


def test_tran_solution_masks_rows_beyond_the_accepted_count():
    c3 = W.c3_colpitts(2).circuit(Circuit)
    t = np.arange(8.).reshape(2, 4)
    sol = results.tran_solution(c3, dict(t=t, x=np.ones((2, 4, 9)), rows=np.array([4, 2]),
                                         status=np.ones(2, np.int32)), 0, 1, 'TRAP')
    T = sol['T']
    assert np.array_equal(T[0], t[0]) and np.isnan(T[1, 2:]).all() and not np.isnan(T[1, :2]).any()
    assert sol.instance(1).shape == (10, 2)


def test_flatten_builds_the_abi_tables():
    w = W.c2_diode_op(8)
    flat = cp.flatten(w.circuit(Circuit), w.batch())
    assert flat.n == 3 and flat.n_nodes == 3 and flat.n_vde == 1 and flat.B == 8
    assert flat.vary.shape == (3, 8) and len(flat.vary_slots) == 3
    # slot-major vary rows hold the batch values of the overridden slots
    nom = np.array(flat.nominal)
    for k, slot in enumerate(flat.vary_slots):
        assert flat.vary[k].shape == (8,) and np.isfinite(nom[slot])
    types = [int(e['type']) for e in flat.elems]
    assert types == [5, 1, 11]                       # AHK_V, AHK_R, AHK_D
    # a V source without a waveform carries 3 slots, not 11
    assert int(flat.elems[1]['pbase']) - int(flat.elems[0]['pbase']) == 3
    # the OP-guess operator reproduces dc_guess.get_dc_guess: x0 = [1, 0.425, 0]
    T = np.array([flat.guess_const[j] + (flat.guess_scale[j] * nom[flat.guess_slot[j]]
                                         if flat.guess_slot[j] >= 0 else 0.0)
                  for j in range(flat.n_guess)])
    x0 = np.asarray(flat.guess_G).reshape(flat.n, flat.n_guess) @ T
    assert np.allclose(x0, [1.0, 0.425, 0.0], atol=1e-12)


def test_unknown_batch_owner_is_rejected():
    w = W.c2_diode_op(4)
    pb = w.batch()
    pb.set('nosuch', 'value', np.ones(4))
    with pytest.raises((KeyError, ValueError)):
        cp.flatten(w.circuit(Circuit), pb)
    with pytest.raises(ValueError):
        ahkab.ParamBatch(4).set('r1', 'value', np.ones(5))


def test_time_functions_match_the_device_encoding():
    """host __call__ == the C evaluation of the encoded 8-slot block (same code the
    kernels run, through the oracle's exported tf_eval)."""
    L = oracle.lib()
    fns = [TF.pulse(0., 1., 1e-9, 2e-9, 5e-9, 3e-9, 20e-9), TF.sin(0.5, 2., 1e8, 1e-9, 1e7, 30.),
           TF.exp(0., 1., 1e-9, 2e-9, 8e-9, 3e-9), TF.sffm(0., 1., 1e8, 2., 1e7, 1e-9),
           TF.am(1., 1e8, 1e7, 0.5, 1e-9),
           TF.pwl((0., 1e-9, 5e-9, 6e-9), (0., 1., 1., -2.), repeat=True, repeat_time=1e-9, td=2e-9),
           TF.pwl((1e-9, 4e-9), (0.5, 2.))]
    for fn in fns:
        kind, prm = TF.encode(fn)
        q = np.zeros(max(8, len(prm)))
        q[:len(prm)] = prm
        for t in np.linspace(0, 45e-9, 91):
            c = L.ahko_tf_eval(C.c_int(kind), q.ctypes.data_as(C.POINTER(C.c_double)),
                               C.c_double(float(t)))
            assert abs(c - fn(float(t))) <= 1e-12 * max(1.0, abs(c)), (type(fn).__name__, t)


def test_x0_helpers():
    w = W.c4_ring3(3)
    circ = w.circuit(Circuit)
    x0 = ahkab.new_x0(circ, W.C4_IC)
    assert x0.shape == (5,)
    assert x0[circ.ext_node_to_int('2') - 1] == 2.5 and x0[circ.ext_node_to_int('1') - 1] == 0
    c3 = W.c3_colpitts(2).circuit(Circuit)
    xm = ahkab.icmodified_x0(c3, np.zeros((2, 9))).numpy()
    assert np.allclose(xm[:, 1], 2.5) and np.allclose(xm[:, 6], -1e-9)   # c1 ic, l1 ic


def test_running_an_analysis_without_a_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    w = W.c2_diode_op(4)
    with pytest.raises(ahkab.EngineError):
        ahkab.run(w.circuit(ahkab.Circuit), ahkab.new_op(), batch=w.batch())


def test_write_csv_uses_the_reference_file_format(tmp_path):
    """csvlib.write_csv layout: '#' + tab-separated names, then one '%.18e' row per sample."""
    c3 = W.c3_colpitts(2).circuit(Circuit)
    t = np.array([[1e-10, 2e-10, 3e-10, 0.0], [1e-10, 2e-10, 0.0, 0.0]])
    x = np.random.default_rng(0).standard_normal((2, 4, 9))
    sol = results.tran_solution(c3, dict(t=t, x=x, rows=np.array([3, 2]),
                                         status=np.ones(2, np.int32)), 0, 1, 'TRAP')
    f = tmp_path / 'a.tran'
    sol.write_csv(str(f), b=0)
    lines = open(f).read().splitlines()
    assert lines[0] == '#' + ' ' + '\t'.join(sol.keys()) or lines[0] == '#' + '\t'.join(sol.keys())
    assert len(lines) == 1 + 3
    back = np.loadtxt(f, delimiter='\t', comments='#')
    assert np.array_equal(back[:, 0], t[0, :3]) and np.array_equal(back[:, 1:], x[0, :3])   # lossless
    op = results.op_solution(c3, dict(x=x[:, 0], resid=x[:, 1], iters=np.zeros(2, np.int32),
                                      status=np.ones(2, np.int32)))
    op.write_csv(str(tmp_path / 'a.op'), b=1)
    back = np.loadtxt(tmp_path / 'a.op', delimiter='\t', comments='#')
    assert np.array_equal(back, x[1, 0])


def test_flatten_is_cached_per_circuit_signature_and_batch_version():
    """run() flattens once per (circuit, batch): the cache lives on the batch and is dropped by
    ParamBatch.set() and by any change to the circuit's elements (plain mutable objects)."""
    w = W.c2_diode_op(8)
    circ, pb = w.circuit(Circuit), w.batch()
    f1 = cp.flatten(circ, pb)
    assert cp.flatten(circ, pb) is f1
    # the same circuit built again has the same signature (models are distinct objects -> new flattening)
    r = [e for e in circ if e.part_id.lower().startswith('v')][0]
    old = r.dc_value
    r.dc_value = old * 2
    f2 = cp.flatten(circ, pb)
    assert f2 is not f1 and not np.array_equal(f2.nominal, f1.nominal)
    r.dc_value = old
    f3 = cp.flatten(circ, pb)
    assert np.array_equal(f3.nominal, f1.nominal) and np.array_equal(f3.vary, f1.vary)
    # new values in the batch: new flattening with the new rows
    (owner, attr), v = next(iter(pb.items()))
    pb.set(owner, attr, v * 1.5)
    f4 = cp.flatten(circ, pb)
    assert f4 is not f3 and not np.array_equal(f4.vary, f3.vary)
    assert cp.topology_key(f4) == cp.topology_key(f3)       # same engine serves both


def test_sources_with_a_waveform_enter_op_with_their_dc_value():
    """VSource.V(time=None) returns dc_value when there is one, else the waveform at t = 0
    (VSource.py:92-97): the staged dc slot holds exactly that."""
    for dc, want in ((0.25, 0.25), (None, 0.5)):
        c = Circuit('exp')
        c.add_vsource('V1', 'a', c.gnd, dc_value=dc, function=TF.exp(.5, -.05, 0., 2e-5, 4e-4, 2e-5))
        c.add_resistor('R1', 'a', c.gnd, 1e3)
        flat = cp.flatten(c)
        assert flat.nominal[int(flat.elems[0]['pbase'])] == want
        x = oracle.Oracle(c).op()['x'][0]
        assert abs(x[0] - want) < 1e-12


def test_csv_files_load_with_the_reference_loader(tmp_path):
    """The optional exporter's files read back bit-exactly through the REFERENCE's own
    csvlib.load_csv (ahkab/csvlib.py:213-290) — skipped where the reference package is absent."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'oracle'))
    import ref_loop
    if ref_loop.reference_path() is None:
        pytest.skip("reference package not on this machine")
    ref = ref_loop._load()
    from ahkab import csvlib
    c3 = W.c3_colpitts(2).circuit(Circuit)
    rng = np.random.default_rng(1)
    t = np.cumsum(rng.uniform(1e-10, 2e-10, (2, 6)), axis=1)
    x = rng.standard_normal((2, 6, 9))
    sol = results.tran_solution(c3, dict(t=t, x=x, rows=np.array([6, 4]),
                                         status=np.ones(2, np.int32)), 0, 1, 'TRAP')
    f = str(tmp_path / 'b.tran')
    sol.write_csv(f, b=1)
    data, headers, pos, eof = csvlib.load_csv(f, load_headers=[], verbose=0)
    assert [h.upper() for h in headers] == [k.upper() for k in sol.keys()]
    assert data.shape == (10, 4)
    assert np.array_equal(data[0], t[1, :4]) and np.array_equal(data[1:], x[1, :4].T)
    # AC: |x| and arg(x) columns (results.py:640-682)
    c1 = W.c1_butterworth(2).circuit(Circuit)
    n = c1.get_nodes_number() - 1 + sum(1 for e in c1 if getattr(e, '_ahk_type', 0) in (3, 5, 7, 9))
    xa = rng.standard_normal((2, 5, n)) + 1j * rng.standard_normal((2, 5, n))
    om = 2 * np.pi * np.logspace(2, 3, 5)
    ac = results.ac_solution(c1, dict(x=xa, omegas=om, status=np.ones(2, np.int32)), 'LOG')
    f = str(tmp_path / 'b.ac')
    ac.write_csv(f, b=0)
    data, headers, pos, eof = csvlib.load_csv(f, load_headers=[], verbose=0)
    assert headers[0].lower() in ('f', '#f') and len(headers) == 1 + 2 * n
    assert np.allclose(data[0], om / 2 / np.pi, rtol=1e-15)
    assert np.array_equal(data[1::2], np.abs(xa[0]).T) and np.array_equal(data[2::2], np.angle(xa[0]).T)
    del ref


def test_opinfo_report_matches_the_reference_report(tmp_path):
    """`.opinfo` (results.py:441-475): same RESULTS table, resistor / source tables and total
    power as the file the reference writes for the same circuit — values fed from the oracle's
    OP.  Skipped where the reference package is absent."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'oracle'))
    import ref_loop
    if ref_loop.reference_path() is None:
        pytest.skip("reference package not on this machine")
    ref = ref_loop._load()
    import zoo

    def build(Circuit):
        c = Circuit('opinfo check')
        c.add_vsource('V1', 'a', c.gnd, dc_value=2.)
        c.add_resistor('R1', 'a', 'b', 1e3)
        c.add_resistor('R2', 'b', c.gnd, 3e3)
        c.add_isource('I1', 'b', c.gnd, dc_value=1e-3)
        return c
    rc = build(ref.Circuit)
    f_ref = str(tmp_path / 'ref.op')
    ref.run(rc, ref.new_op(outfile=f_ref, verbose=0))
    ours = build(Circuit)
    o = oracle.Oracle(ours).op()
    sol = results.op_solution(ours, dict(x=o['x'], resid=o['resid'], iters=o['iters'], status=o['status']))
    f_our = str(tmp_path / 'our.opinfo')
    sol.write_opinfo(f_our, ours)
    a, b = open(f_ref + '.opinfo', encoding='utf-8').read(), open(f_our, encoding='utf-8').read()   # new_op appends '.op', write_to_file 'info'

    def section(txt, start, stop):
        return txt[txt.index(start):txt.index(stop)]
    # RESULTS table: same variables, units and values (the Error column is the solver residual)
    ra = [ln.split() for ln in section(a, 'RESULTS:', 'ELEMENTS OP').splitlines() if ln[:1] in 'VI' and '(' in ln or ln[:2] in ('VA', 'VB')]
    rb = [ln.split() for ln in section(b, 'RESULTS:', 'ELEMENTS OP').splitlines() if ln[:1] in 'VI' and '(' in ln or ln[:2] in ('VA', 'VB')]
    assert [r[:2] for r in ra] == [r[:2] for r in rb] and len(ra) == 3
    assert np.allclose([float(r[2]) for r in ra], [float(r[2]) for r in rb], rtol=1e-5, atol=1e-12)
    # element tables and total power, number by number
    import re
    num = re.compile(r'-?\d+\.?\d*(?:e[-+]?\d+)?')
    ta = section(a, 'ELEMENTS OP INFORMATION', 'Total power')
    tb = section(b, 'ELEMENTS OP INFORMATION', 'Total power')
    for pid in ('R1', 'R2', 'V1', 'I1'):
        la = [ln for ln in ta.splitlines() if ln.startswith(pid)][0]
        lb = [ln for ln in tb.splitlines() if ln.startswith(pid)][0]
        assert np.allclose([float(t) for t in num.findall(la)[1:]], [float(t) for t in num.findall(lb)[1:]], rtol=1e-5, atol=1e-12), (la, lb)
    pa = float(num.findall(a[a.index('Total power dissipation'):])[0])
    pb = float(num.findall(b[b.index('Total power dissipation'):])[0])
    assert abs(pa - pb) <= 1e-5 * abs(pa)
    assert 'Convergence reached in' in b and 'vea =' in b
