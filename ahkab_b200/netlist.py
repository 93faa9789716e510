"""Netlist front-end reuse.

The reference's SPICE-like netlist parser (`ahkab/netlist_parser.py:145 parse_circuit`,
`:1428 parse_analysis`) is front-end code far from the hot path: it is not rebuilt here.  A
user who has the reference installed parses with IT and hands the resulting circuit to the
batched engine through `from_reference`: the reference's `Circuit` (a list of element objects
plus node maps, `ahkab/circuit.py:136`) is converted element by element into an
`ahkab_b200.Circuit` with the same node numbering, element order and part ids, so results
carry the same variable names in the same order.  Nothing of the reference is imported here:
the conversion goes by class names and attributes.

    import ahkab, ahkab_b200
    from ahkab import netlist_parser
    rc, directives, _ = netlist_parser.parse_circuit('filter.ckt')
    circ = ahkab_b200.netlist.from_reference(rc)
    ans = ahkab_b200.netlist.convert_analyses(netlist_parser.parse_analysis(rc, directives))
    res = ahkab_b200.run(circ, ans, batch=ahkab_b200.ParamBatch(4096).set('R1', 'value', r1_values))

Models arrive with their DERIVED parameters (the reference's constructors have already run),
so a converted model is frozen: a `ParamBatch` may vary element values (R, C, L, source values,
W / L / M / N, AREA) of a converted circuit, not its model parameters — build the circuit with
`Circuit.add_model` for that."""
import cmath

from . import circuit as _c
from . import constants
from . import time_functions as TF


def _frozen(cls, name, values, npmos=None):
    m = cls.__new__(cls)
    m.name = name
    if npmos is not None:
        m.NPMOS = int(npmos)
    m._kw = {}
    frozen = {k: float(values[k]) for k in cls.RUNTIME}

    def runtime(overrides, frozen=frozen, name=name):
        if overrides:
            raise KeyError("model %s was converted from a reference circuit: its derived parameters "
                           "are fixed (vary them on a circuit built with Circuit.add_model)" % name)
        return dict(frozen)
    m.runtime = runtime
    for k, v in frozen.items():
        setattr(m, k, v)
    return m


def _model(rm):
    kind = type(rm).__name__
    if kind == 'ekv_mos_model':
        vals = {k: getattr(rm, k) for k in _c.ekv_mos_model.RUNTIME}
        return _frozen(_c.ekv_mos_model, rm.name, vals, rm.NPMOS)
    if kind == 'mosq_mos_model':
        vals = {k: getattr(rm, k) for k in _c.mosq_mos_model.RUNTIME}
        return _frozen(_c.mosq_mos_model, rm.name, vals, rm.NPMOS)
    if kind == 'diode_model':
        vals = {k: getattr(rm, k) for k in _c.diode_model.RUNTIME if k != 'VT'}
        vals['VT'] = constants.Vth(26.85 + 273.15)
        return _frozen(_c.diode_model, rm.name, vals)
    if kind == 'vswitch_model':
        vals = {k: getattr(rm, k) for k in _c.vswitch_model.RUNTIME}
        return _frozen(_c.vswitch_model, rm.name, vals)
    raise NotImplementedError("model class %s" % kind)


_TF_ARGS = {'pulse': ('v1', 'v2', 'td', 'tr', 'pw', 'tf', 'per'),
            'sin': ('vo', 'va', 'freq', 'td', 'theta', 'phi'),
            'exp': ('v1', 'v2', 'td1', 'tau1', 'td2', 'tau2'),
            'sffm': ('vo', 'va', 'fc', 'mdi', 'fs', 'td'),
            'am': ('sa', 'fc', 'fm', 'oc', 'td'),
            'pwl': ('x', 'y', 'repeat', 'repeat_time', 'td')}


def _time_function(f):
    if f is None:
        return None
    kind = type(f).__name__
    if kind not in _TF_ARGS:
        raise NotImplementedError("time function %s" % kind)
    return getattr(TF, kind)(**{k: getattr(f, k) for k in _TF_ARGS[kind]})


def from_reference(rc):
    """ahkab.Circuit -> ahkab_b200.Circuit (same nodes, element order, part ids)."""
    c = _c.Circuit(rc.title, getattr(rc, 'filename', None))
    for i in range(rc.get_nodes_number()):
        c.add_node(rc.nodes_dict[i])
    c.internal_nodes = getattr(rc, 'internal_nodes', 0)
    ext = rc.nodes_dict
    models = {}

    def model_of(rm):
        if id(rm) not in models:
            models[id(rm)] = _model(rm)
            c.models[rm.name] = models[id(rm)]
        return models[id(rm)]
    for label, rm in (rc.models or {}).items():
        c.models[label] = model_of(rm)
    for e in rc:
        kind = type(e).__name__
        pid = e.part_id
        if kind == 'Resistor':
            c.add_resistor(pid, ext[e.n1], ext[e.n2], e.value)
        elif kind == 'Capacitor':
            c.add_capacitor(pid, ext[e.n1], ext[e.n2], e.value, ic=e.ic)
        elif kind == 'Inductor':
            c.add_inductor(pid, ext[e.n1], ext[e.n2], e.value, ic=e.ic)
        elif kind == 'InductorCoupling':
            c.add_inductor_coupling(pid, e.L1, e.L2, e.K)
        elif kind in ('VSource', 'ISource'):
            ac = 0
            if getattr(e, 'abs_ac', None) is not None:
                ac = e.abs_ac * cmath.exp(1j * (e.arg_ac or 0.0))
                if abs(ac.imag) == 0.0:
                    ac = ac.real
            fn = _time_function(e._time_function) if e.is_timedependent else None
            add = c.add_vsource if kind == 'VSource' else c.add_isource
            add(pid, ext[e.n1], ext[e.n2], dc_value=e.dc_value, ac_value=ac, function=fn)
            if getattr(e, 'abs_ac', None) is None:
                c[-1].abs_ac = None
        elif kind == 'EVSource':
            c.add_vcvs(pid, ext[e.n1], ext[e.n2], ext[e.sn1], ext[e.sn2], e.alpha)
        elif kind == 'GISource':
            c.add_vccs(pid, ext[e.n1], ext[e.n2], ext[e.sn1], ext[e.sn2], e.alpha)
        elif kind == 'HVSource':
            c.add_ccvs(pid, ext[e.n1], ext[e.n2], e.source_id, e.alpha)
        elif kind == 'FISource':
            c.add_cccs(pid, ext[e.n1], ext[e.n2], e.source_id, e.alpha)
        elif kind == 'diode':
            m = model_of(e.model)
            c.add_diode(pid, ext[e.n1], ext[e.n2], m.name, models={m.name: m}, Area=e.device.AREA,
                        T=None, ic=e.ic, off=getattr(e, 'off', False))
        elif kind in ('ekv_device', 'mosq_device'):
            rm = e.ekv_model if kind == 'ekv_device' else e.mosq_model
            m = model_of(rm)
            c.add_mos(pid, ext[e.n1], ext[e.ng], ext[e.n2], ext[e.nb], e.device.W, e.device.L, m.name,
                      models={m.name: m}, m=e.device.M, n=e.device.N)
        elif kind == 'switch_device':
            m = model_of(e.model)
            c.add_switch(pid, ext[e.n1], ext[e.n2], ext[e.sn1], ext[e.sn2], bool(e.device.is_on), m.name,
                         models={m.name: m})
        else:
            raise NotImplementedError("element %s of class %s cannot run on the batched engine "
                                      "(user-defined Python elements are out of scope)" % (pid, kind))
    return c


_AN_KEYS = {'op': ('guess', 'x0', 'outfile', 'verbose'),
            'dc': ('start', 'stop', 'step', 'source', 'sweep_type', 'guess', 'x0', 'outfile', 'verbose'),
            'tran': ('tstart', 'tstop', 'tstep', 'x0', 'method', 'use_step_control', 'outfile', 'verbose'),
            'ac': ('start', 'stop', 'points', 'sweep_type', 'x0', 'outfile', 'verbose')}


def convert_analyses(an_list):
    """The analysis dicts of the reference's `parse_analysis` / `new_*` (ahkab.py:201-448), cut
    down to the four batched analyses and their arguments; other directives (.pz, .pss,
    .symbolic, .temp …) are outside the hot path and are dropped with their names returned."""
    out, dropped = [], []
    for an in an_list:
        t = an.get('type')
        if t not in _AN_KEYS:
            dropped.append(t)
            continue
        d = {'type': t}
        for k in _AN_KEYS[t]:
            if k in an:
                d[k] = an[k]
        d['outfile'] = None
        if isinstance(d.get('x0'), str) or d.get('x0') is None:
            pass
        else:       # an op_solution object of the reference: its values as a plain vector
            x0 = d['x0']
            d['x0'] = x0.asarray().ravel() if hasattr(x0, 'asarray') else x0
        out.append(d)
    convert_analyses.dropped = dropped
    return out
