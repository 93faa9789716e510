"""Topology compiler: `Circuit` (+ `ParamBatch`) -> the flat tables that cross
the C-ABI (include/ahkab_b200.h: ahk_elem[], nominal[], vary_slots/vary,
guess operator).

There is no reference counterpart: ahkab walks the Python element list on
every assembly (dc_analysis.py:949-1069, transient.py:472-561, ac.py:323-443).
Here the walk happens once; the device-side stamp program is derived from the
element table inside the library (csrc/program.cpp)."""
import numpy as np

from . import circuit as C
from . import dc_guess as _guess
from . import time_functions as TF

ELEM_DTYPE = np.dtype([(k, np.int32) for k in (
    'type', 'n1', 'n2', 'n3', 'n4', 'vde', 'ref', 'ref2', 'aux0', 'aux1',
    'pbase', 'mbase', 'flags', 'tf')])

FLAG_HAS_AC = 0x100
FLAG_HAS_IC = 0x200
N_TF = 8


def _is_arr(v):
    return isinstance(v, np.ndarray) and v.ndim == 1


class Flat(object):
    """Result of `flatten`: everything the engine needs for one topology and
    (optionally) one batch.  `vary` ([n_vary][B], slot-major) is stacked from `vary_rows` on
    first use: the engine uploads a fresh batch row by row into its pinned staging instead."""
    _vary = None
    vary_rows = ()
    uploads = 0

    @property
    def vary(self):
        if self._vary is None:
            self._vary = (np.ascontiguousarray(np.stack(self.vary_rows)) if len(self.vary_rows)
                          else np.zeros((0, self.B or 0)))
        return self._vary

    @vary.setter
    def vary(self, v):
        self._vary = v


def _src_block(elem, ov, B):
    """[dc, |ac|, arg(ac), tf0..tf7] for a V/I source."""
    dc = ov.get('dc_value', elem.dc_value)
    if dc is None:
        # VSource.V / ISource.I at time None (OP, DC): dc_value if there is one, else the
        # waveform, which reads None as t = 0 (VSource.py:92-97, time_functions.py:441-442)
        dc = float(elem._time_function(0.)) if elem.is_timedependent else 0.0
    flags = 0
    if 'ac_value' in ov:
        ac = ov['ac_value']
        abs_ac, arg_ac = np.abs(ac), np.angle(ac)
        flags |= FLAG_HAS_AC
    elif elem.abs_ac is not None:
        abs_ac, arg_ac = float(elem.abs_ac), float(elem.arg_ac)
        flags |= FLAG_HAS_AC
    else:
        abs_ac, arg_ac = 0.0, 0.0
    tfk, tfp = TF.encode(elem._time_function) if elem.is_timedependent \
        else (0, [])
    if tfk == TF.TF_PWL:
        tfp = list(tfp)                       # variable length: 4 + 2 * npts
    elif tfk:
        tfp = list(tfp) + [0.0] * (N_TF - len(tfp))
    else:
        tfp = []      # no waveform -> no tf slots (keeps the staged block small)
    return [dc, abs_ac, arg_ac] + tfp, flags, tfk


_GUESS_CACHE = {}
_SCALARS = (int, float, bool, str, type(None), complex)


def _scalars(o):
    """The plain-valued attributes of an element / model / device / waveform object."""
    out = []
    for k, v in vars(o).items():
        if type(v) in _SCALARS or isinstance(v, (np.floating, np.integer)):
            out.append((k, v))
        elif isinstance(v, np.ndarray):
            out.append((k, v.tobytes()))
        elif type(v) in (list, tuple) and all(type(t) in _SCALARS for t in v):
            out.append((k, tuple(v)))
    return tuple(out)


def circuit_signature(circ):
    """Everything `flatten` reads from the circuit, as a hashable value: the elements (in
    order) with their nodes / values, their device blocks, waveforms and models.  Elements
    are plain mutable objects (`r.value = ...` is legal, as in the reference), so a cached
    flattening is only reused while this signature is unchanged."""
    sig = [circ.get_nodes_number()]
    for e in circ:
        item = [type(e).__name__, _scalars(e)]
        for sub in ('device', '_time_function'):
            so = getattr(e, sub, None)
            if so is not None and hasattr(so, '__dict__'):
                item.append(_scalars(so))
        m = getattr(e, 'model', None)
        if m is not None:
            item.append(id(m))
        sig.append(tuple(item))
    for label, m in circ.models.items():
        sig.append((label, id(m), _scalars(m)))
    return tuple(sig)


def flatten(circ, batch=None):
    """Flatten `circ`; with `batch` also build the vary tables.  The result is cached on the
    batch (keyed by the circuit's signature and the batch's version), so `run()` on the same
    circuit and batch — analysis after analysis — flattens once."""
    if batch is None:
        return _flatten(circ, None)
    sig = circuit_signature(circ)
    hit = getattr(batch, '_flat_cache', None)
    if hit is not None and hit[0] == batch._version and hit[1] == sig:
        return hit[2]
    flat = _flatten(circ, batch)
    batch._flat_cache = (batch._version, sig, flat)
    return flat


def _flatten(circ, batch=None):
    B = batch.B if batch is not None else None
    vals = []      # per slot: float or (B,) array
    names = []     # per slot: (owner, name)
    elems = np.zeros(len(circ), dtype=ELEM_DTYPE)
    for f in ('vde', 'ref', 'ref2', 'aux0', 'aux1', 'mbase'):
        elems[f] = -1
    used_owners = set()

    def ovs(owner):
        if batch is None:
            return {}
        d = batch.for_owner(owner)
        if d:
            used_owners.add(owner.lower())
        return d

    def push(owner, block):
        base = len(vals)
        for nm, v in block:
            if _is_arr(v):
                if B is None or v.shape != (B,):
                    raise ValueError("bad batch array for %s.%s" % (owner, nm))
                v = np.asarray(v, dtype=np.float64)     # (no copy: ParamBatch stores float64)
            else:
                v = float(v)
            vals.append(v)
            names.append((owner, nm))
        return base

    # model blocks first (shared by all devices using the model)
    model_base = {}
    for label, m in circ.models.items():
        rt = m.runtime(ovs(label))
        model_base[id(m)] = push(label, [(k, rt[k]) for k in m.RUNTIME])

    def mbase_of(m):
        if id(m) not in model_base:      # model passed via models= argument
            rt = m.runtime(ovs(m.name))
            model_base[id(m)] = push(m.name, [(k, rt[k]) for k in m.RUNTIME])
        return model_base[id(m)]

    vde_index = {}
    k = 0
    for e in circ:
        if C.is_elem_voltage_defined(e):
            vde_index[e.part_id.lower()] = k
            k += 1
    n_vde = k
    pbase_of = {}

    def allowed(ov, keys, pid):
        bad = set(ov) - set(keys)
        if bad:
            raise KeyError("element %s has no batchable parameter(s) %s "
                           "(allowed: %s)" % (pid, sorted(bad), sorted(keys)))

    for i, e in enumerate(circ):
        r = elems[i]
        pid = e.part_id
        ov = ovs(pid)
        r['type'] = e._ahk_type
        if C.is_elem_voltage_defined(e):
            r['vde'] = vde_index[pid.lower()]
        if isinstance(e, C.Resistor):
            allowed(ov, ('value',), pid)
            r['n1'], r['n2'] = e.n1, e.n2
            r['pbase'] = push(pid, [('value', ov.get('value', e.value))])
        elif isinstance(e, (C.Capacitor, C.Inductor)):
            allowed(ov, ('value', 'ic'), pid)
            r['n1'], r['n2'] = e.n1, e.n2
            ic = ov.get('ic', e.ic)
            if ic is not None and (_is_arr(ic) or ic):
                # `elem.ic` truthiness test of dc_analysis.py:1172-1179
                r['flags'] |= FLAG_HAS_IC
            r['pbase'] = push(pid, [('value', ov.get('value', e.value)),
                                    ('ic', 0.0 if ic is None else ic)])
        elif isinstance(e, C.InductorCoupling):
            if 'value' in ov:
                ov['K'] = ov.pop('value')
            allowed(ov, ('K',), pid)
            r['ref'] = vde_index[e.L1.lower()]
            r['ref2'] = vde_index[e.L2.lower()]
            r['pbase'] = push(pid, [('K', ov.get('K', e.K))])
        elif isinstance(e, (C.VSource, C.ISource)):
            allowed(ov, ('dc_value', 'ac_value'), pid)
            r['n1'], r['n2'] = e.n1, e.n2
            blk, fl, tfk = _src_block(e, ov, B)
            r['flags'] |= fl
            r['tf'] = tfk
            nm = ['dc_value', 'abs_ac', 'arg_ac'] + ['tf%d' % j
                                                     for j in range(len(blk))]
            r['pbase'] = push(pid, list(zip(nm[:len(blk)], blk)))
        elif isinstance(e, (C.EVSource, C.GISource)):
            if 'value' in ov:
                ov['alpha'] = ov.pop('value')
            allowed(ov, ('alpha',), pid)
            r['n1'], r['n2'], r['n3'], r['n4'] = e.n1, e.n2, e.sn1, e.sn2
            r['pbase'] = push(pid, [('alpha', ov.get('alpha', e.alpha))])
        elif isinstance(e, (C.HVSource, C.FISource)):
            if 'value' in ov:
                ov['alpha'] = ov.pop('value')
            allowed(ov, ('alpha',), pid)
            r['n1'], r['n2'] = e.n1, e.n2
            r['ref'] = vde_index[e.source_id.lower()]
            r['pbase'] = push(pid, [('alpha', ov.get('alpha', e.alpha))])
        elif isinstance(e, C.diode):
            if 'Area' in ov:
                ov['AREA'] = ov.pop('Area')
            allowed(ov, ('AREA',), pid)
            r['n1'], r['n2'] = e.n1, e.n2
            r['pbase'] = push(pid, [('AREA', ov.get('AREA', e.device.AREA))])
            r['mbase'] = mbase_of(e.model)
        elif isinstance(e, C.ekv_device):
            ov = {k_.upper(): v for k_, v in ov.items()}
            allowed(ov, ('W', 'L', 'M', 'N'), pid)
            r['n1'], r['n2'], r['n3'], r['n4'] = e.n1, e.ng, e.n2, e.nb
            d = e.device
            r['pbase'] = push(pid, [('W', ov.get('W', d.W)),
                                    ('L', ov.get('L', d.L)),
                                    ('M', ov.get('M', d.M)),
                                    ('N', ov.get('N', d.N))])
            r['mbase'] = mbase_of(e.model)
            r['flags'] = e.model.NPMOS
        elif isinstance(e, C.mosq_device):
            ov = {(k_.upper() if len(k_) == 1 else k_): v
                  for k_, v in ov.items()}
            allowed(ov, ('W', 'L', 'M', 'N', 'z_vt', 'z_kp'), pid)
            r['n1'], r['n2'], r['n3'], r['n4'] = e.n1, e.ng, e.n2, e.nb
            d = e.device
            mck = d.mckey if d.mckey else (0.0, 0.0)
            r['pbase'] = push(pid, [('W', ov.get('W', d.W)),
                                    ('L', ov.get('L', d.L)),
                                    ('M', ov.get('M', d.M)),
                                    ('N', ov.get('N', d.N)),
                                    ('z_vt', ov.get('z_vt', mck[0])),
                                    ('z_kp', ov.get('z_kp', mck[1]))])
            r['mbase'] = mbase_of(e.model)
            r['flags'] = e.model.NPMOS
        elif isinstance(e, C.switch_device):
            allowed(ov, (), pid)
            r['n1'], r['n2'], r['n3'], r['n4'] = e.n1, e.n2, e.sn1, e.sn2
            r['pbase'] = push(pid, [('is_on',
                                     1.0 if e.device.is_on else 0.0)])
            r['mbase'] = mbase_of(e.model)
        else:
            raise TypeError("element %r is not supported by the batched "
                            "engine" % (e,))
        pbase_of[pid.lower()] = int(r['pbase'])

    for i, e in enumerate(circ):
        if isinstance(e, C.InductorCoupling):
            elems[i]['aux0'] = pbase_of[e.L1.lower()]
            elems[i]['aux1'] = pbase_of[e.L2.lower()]

    if batch is not None:
        unknown = batch.owners() - used_owners
        if unknown:
            raise KeyError("batch refers to unknown elements/models: %s"
                           % sorted(unknown))

    # ---- OP guess rows, in the reference's order (dc_guess.py:112-147) ----
    hint_rows = []
    for i, e in enumerate(circ):
        if getattr(e, 'dc_guess', None) is None:
            continue
        r = elems[i]
        if e.is_nonlinear:
            if isinstance(e, C.diode):
                g = e.dc_guess if isinstance(e.dc_guess, (list, tuple)) \
                    else [e.dc_guess]
                hints = [(-1, 0.0, float(g[0]))]
            elif isinstance(e, C.ekv_device):
                NP = e.model.NPMOS
                vto = int(r['mbase'])      # VTO is slot 0 of the model block
                hints = [(vto, 0.1 * NP, 0.0), (vto, 1.1 * NP, 0.0),
                         (-1, 0.0, 0.0)]
            elif isinstance(e, C.mosq_device):
                NP = e.model.NPMOS
                vto = int(r['mbase'])
                hints = [(vto, 0.4 * NP, 0.0), (vto, 1.1 * NP, 0.0),
                         (-1, 0.0, 0.0)]
            elif isinstance(e, C.switch_device):
                vt = int(r['mbase'])       # VT is slot 0
                sc = .9 + e.device.is_on * .2
                hints = [(vt, sc, 0.0), (vt, sc, 0.0)]
            port_index = 0
            for (n1, n2) in e.ports:
                if n1 == n2:
                    continue
                s, k_, c0 = hints[port_index]
                hint_rows.append((n1, n2, s, k_, c0))
                port_index += 1
        else:
            if e.n1 == e.n2:
                continue
            hint_rows.append((e.n1, e.n2, int(r['pbase']), 1.0, 0.0))
    # (topology only: cached per hint rows — a fresh batch of the same circuit re-uses it)
    gkey = (circ.get_nodes_number(), tuple(hint_rows), tuple(bool(C.is_elem_voltage_defined(e)) for e in circ))
    hit = _GUESS_CACHE.get(gkey)
    if hit is None:
        hit = _guess.build_guess_operator(circ, hint_rows)
        if len(_GUESS_CACHE) > 64:
            _GUESS_CACHE.clear()
        _GUESS_CACHE[gkey] = hit
    G, kept = hit

    out = Flat()
    out.n_nodes = circ.get_nodes_number()
    out.n_vde = n_vde
    out.n = out.n_nodes - 1 + n_vde
    out.elems = elems
    out.names = names
    out.nonlinear = circ.is_nonlinear()
    nominal = np.empty(len(vals), dtype=np.float64)
    vary_slots, vary_rows = [], []
    for s, v in enumerate(vals):
        if _is_arr(v):
            nominal[s] = v[0]
            vary_slots.append(s)
            vary_rows.append(v)
        else:
            nominal[s] = v
    out.nominal = nominal
    out.vary_slots = np.asarray(vary_slots, dtype=np.int32)
    out.vary_rows = vary_rows
    out.B = B
    out.vary_pinned = None      # pinned torch copy of `vary`, made by the engine on first upload
    if G is None:
        out.n_guess = 0
        out.guess_G = np.zeros((out.n, 0))
        out.guess_slot = np.zeros(0, np.int32)
        out.guess_scale = np.zeros(0)
        out.guess_const = np.zeros(0)
    else:
        rows = [hint_rows[i] for i in kept]
        out.n_guess = len(rows)
        out.guess_G = np.ascontiguousarray(G, dtype=np.float64)
        out.guess_slot = np.asarray([r_[2] for r_ in rows], np.int32)
        out.guess_scale = np.asarray([r_[3] for r_ in rows], np.float64)
        out.guess_const = np.asarray([r_[4] for r_ in rows], np.float64)
        assert out.guess_G.shape == (out.n, out.n_guess), \
            (out.guess_G.shape, out.n, out.n_guess)
    return out


def topology_key(flat):
    """Hashable identity of the topology + nominal values of the fixed slots + vary slot set
    (what an engine handle is specific to)."""
    nom = flat.nominal
    if len(flat.vary_slots):
        # a varied slot's nominal is just the first instance's value (every instance overrides
        # it): it must not make a new batch of the same sweep look like a new circuit
        nom = nom.copy()
        nom[flat.vary_slots] = 0.0
    return (flat.elems.tobytes(), nom.tobytes(),
            flat.vary_slots.tobytes(), flat.guess_G.tobytes())
