"""The reference's analysis API with a batch axis.

`new_op / new_dc / new_tran / new_ac`, `queue`, `run`, `new_x0`,
`icmodified_x0`, `get_op_x0` keep the signatures and semantics of
ahkab/ahkab.py:201-448,660-803,962-975; `run(circ, an_list, batch=ParamBatch)`
executes each analysis for all B instances at once on the GPU through the
C-ABI (include/ahkab_b200.h) instead of `analysis[an_type](circ, **an_item)`
(ahkab/ahkab.py:714,860-863)."""
import copy
import re

import numpy as np
import torch

from . import options
from . import circuit as _c
from . import compile as _compile
from . import results
from . import utilities
from .engine import Engine

_queue = []
_x0s = {None: None}
_engines = {}


def new_op(guess=None, x0=None, outfile=None, verbose=0):
    """ahkab/ahkab.py:201-249."""
    if guess is None:
        guess = options.dc_use_guess
    return {'type': 'op', 'guess': guess, 'x0': x0, 'outfile': outfile,
            'verbose': verbose}


def new_dc(start, stop, points, source, sweep_type='LINEAR', guess=True,
           x0=None, outfile=None, verbose=0):
    """ahkab/ahkab.py:252-319 (points -> step, :317)."""
    step = (stop - start) / (points - 1)
    return {'type': 'dc', 'start': float(start), 'stop': float(stop),
            'step': float(step), 'source': source, 'x0': x0,
            'outfile': outfile, 'guess': guess, 'sweep_type': sweep_type,
            'verbose': verbose}


def new_tran(tstart, tstop, tstep, x0='op', method=None,
             use_step_control=True, outfile=None, verbose=0):
    """ahkab/ahkab.py:322-384."""
    if method is None:
        method = options.default_tran_method
    return {"type": "tran", "tstart": tstart, "tstop": tstop, "tstep": tstep,
            "method": method, "use_step_control": use_step_control, 'x0': x0,
            'outfile': outfile, 'verbose': verbose}


def new_ac(start, stop, points, x0='op', sweep_type='LOG', outfile=None,
           verbose=0):
    """ahkab/ahkab.py:387-448."""
    return {'type': 'ac', 'start': start, 'stop': stop, 'points': points,
            'sweep_type': sweep_type, 'x0': x0, 'outfile': outfile,
            'verbose': verbose}


def queue(*analysis):
    """ahkab/ahkab.py:660-674."""
    global _queue
    for an in analysis:
        _queue += [copy.deepcopy(an)]


def get_engine(circ, batch=None, opts=None):
    """Engines are cached per (topology, nominal values, varied slots)."""
    flat = _compile.flatten(circ, batch)
    key = _compile.topology_key(flat)
    eng = _engines.get(key)
    if eng is None:
        eng = Engine(circ, batch, opts=opts, flat=flat)
        if len(_engines) > 16:
            _engines.clear()
        _engines[key] = eng
    else:
        eng.circ = circ
        eng.flat = flat
        eng.set_vary_flat(flat)
        o = options.snapshot()
        if opts:
            o.update(opts)
        if o != eng.opts:          # unchanged options: keep the engine's cached plans / kernels
            eng.set_options(**o)
    return eng


# ---- x0 helpers -----------------------------------------------------------------
def new_x0(circ, icdict):
    """dc_analysis.build_x0_from_user_supplied_ic (dc_analysis.py:1072-1140).
    Values may be scalars or (B,) arrays -> (n,) or (B, n)."""
    Vregex = re.compile(r"V\s*\(\s*([a-z0-9]+)\s*\)", re.IGNORECASE | re.DOTALL)
    Iregex = re.compile(r"I\s*\(\s*([a-z0-9]+)\s*\)", re.IGNORECASE | re.DOTALL)
    nv = circ.get_nodes_number()
    vde = [e.part_id.lower() for e in circ if _c.is_elem_voltage_defined(e)]
    B = 1
    for v in icdict.values():
        if np.ndim(v) == 1:
            B = len(v)
    x0 = np.zeros((B, nv + len(vde)))
    for label, value in icdict.items():
        if Vregex.search(label):
            x0[:, circ.ext_node_to_int(Vregex.findall(label)[0])] = value
        elif Iregex.search(label):
            x0[:, nv + vde.index(Iregex.findall(label)[0].lower())] = value
        else:
            raise ValueError("Unrecognized label " + label)
    x0 = x0[:, 1:]
    return x0[0] if B == 1 else x0


def icmodified_x0(circ, x0, batch=None):
    """dc_analysis.modify_x0_for_ic (dc_analysis.py:1143-1190), batched.

    x0: op_solution or (B, n) tensor/array.  Capacitors / diodes with `ic` pin
    V(n1) = V(n2) + ic, inductors with `ic` pin their current; index -1 for a
    grounded node wraps to the last unknown exactly like the reference's numpy
    indexing does."""
    if isinstance(x0, results.op_solution):
        x = x0.raw['x'].clone()
    else:
        x = torch.as_tensor(x0, dtype=torch.float64).clone()
    if x.dim() == 1:
        x = x.reshape(1, -1)
    nv = circ.get_nodes_number()
    vde = [e for e in circ if _c.is_elem_voltage_defined(e)]

    def val(e):
        ic = e.ic
        if batch is not None:
            ov = batch.for_owner(e.part_id)
            if 'ic' in ov:
                ic = torch.as_tensor(ov['ic'], dtype=x.dtype, device=x.device)
        return ic

    def truthy(ic):
        return ic is not None and (torch.is_tensor(ic) or bool(ic))
    for e in circ:
        if isinstance(e, (_c.Capacitor, _c.diode)):
            ic = val(e) if isinstance(e, _c.Capacitor) else e.ic
            if truthy(ic):
                x[:, e.n1 - 1] = x[:, e.n2 - 1] + ic
    for e in vde:
        if isinstance(e, _c.Inductor):
            ic = val(e)
            if truthy(ic):
                x[:, nv - 1 + vde.index(e)] = ic
    return x


def get_op_x0(circ, batch=None):
    """ahkab/ahkab.py:781-795."""
    return run(circ, [new_op()], batch=batch)


# ---- run -----------------------------------------------------------------------------
def _resolve_x0(x0, eng):
    if x0 is None:
        return None
    if isinstance(x0, results.op_solution):
        return x0.raw['x']
    return x0


def _run_op(circ, eng, guess=True, x0=None, outfile=None, verbose=0):
    """dc_analysis.op_analysis (dc_analysis.py:598-687)."""
    if not options.dc_use_guess:
        guess = False
    raw = eng.op(x0=_resolve_x0(x0, eng), use_guess=guess)
    return results.op_solution(circ, raw, engine=eng)


def _run_dc(circ, eng, start, stop, step, source, sweep_type='LINEAR',
            guess=True, x0=None, outfile=None, verbose=0):
    """dc_analysis.dc_analysis (dc_analysis.py:461-595)."""
    elem_type, elem_descr = source[0].lower(), source.lower()
    sweep_label = elem_type[0].upper() + elem_descr[1:]
    if sweep_type == options.dc_log_step and stop - start < 0:
        raise ValueError("DC analysis has log sweeping and negative stepping.")
    if (stop - start) * step < 0:
        raise ValueError("Unbonded stepping in DC analysis.")
    points = (stop - start) / step + 1
    stype = sweep_type.upper()[:3]
    if stype == options.dc_log_step:
        it = utilities.log_axis_iterator(start, stop, points=points)
    elif stype == options.dc_lin_step:
        it = utilities.lin_axis_iterator(start, stop, points=points)
    else:
        raise ValueError("Unknown sweep type: %s" % (sweep_type,))
    if elem_type != 'v' and elem_type != 'i':
        raise ValueError("Sweeping is possible only with voltage and current "
                         "sources. (" + str(elem_type) + ")")
    src = None
    for i, e in enumerate(circ):
        if e.part_id.lower() == elem_descr and isinstance(
                e, _c.VSource if elem_type == 'v' else _c.ISource):
            src = i
            break
    if src is None:
        raise ValueError(".DC: source %s was not found." % source)
    sweep = np.array(list(it), dtype=np.float64)
    raw = eng.dc_sweep(src, sweep, x0=_resolve_x0(x0, eng), use_guess=guess)
    return results.dc_solution(circ, raw, sweepvar=sweep_label, stype=stype, engine=eng)


def _run_tran(circ, eng, tstart, tstep, tstop, method=None,
              use_step_control=True, x0=None, outfile=None, verbose=0,
              max_rows=None):
    """transient.transient_analysis (transient.py:121-428)."""
    if options.transient_no_step_control:
        use_step_control = False
    method = method.upper() if method is not None else \
        options.default_tran_method
    if tstart > tstop:
        raise ValueError("tstart > tstop")
    if tstep < 0:
        raise ValueError("tstep < 0")
    raw = eng.tran(_resolve_x0(x0, eng), tstart, tstep, tstop, method=method,
                   use_step_control=use_step_control, max_rows=max_rows)
    return results.tran_solution(circ, raw, tstart, tstop, method, engine=eng)


def _run_ac(circ, eng, start, points, stop, sweep_type=None, x0=None,
            outfile=None, verbose=0):
    """ac.ac_analysis (ac.py:146-320)."""
    if sweep_type is None:
        sweep_type = options.ac_log_step
    if not (start > 0 and stop > 0):          # ac.py:203-214
        raise ValueError("AC analysis has start or stop frequency <= 0")
    if stop < start:
        raise ValueError("AC analysis has start > stop")
    if points < 2 and not start == stop:
        raise ValueError("AC analysis has number of points < 2 & start != stop")
    stype = sweep_type.upper()[:3]
    if stype == options.ac_log_step:
        it = utilities.log_axis_iterator(2 * np.pi * start, 2 * np.pi * stop,
                                         points)
    elif stype == options.ac_lin_step:
        it = utilities.lin_axis_iterator(2 * np.pi * start, 2 * np.pi * stop,
                                         points)
    else:
        raise ValueError("Unknown sweep type %s" % sweep_type)
    omegas = np.array(list(it), dtype=np.float64)
    x_op = _resolve_x0(x0, eng)
    if circ.is_nonlinear() and x_op is None:   # ac.py:243-256
        op = _run_op(circ, eng)
        if not bool(op.ok.all()):
            raise RuntimeError("OP analysis failed, no linearization point "
                               "available.")
        x_op = op.raw['x']
    raw = eng.ac(omegas, x_op=x_op if circ.is_nonlinear() else None)
    return results.ac_solution(circ, raw, stype, engine=eng)


_analysis = {'op': _run_op, 'dc': _run_dc, 'tran': _run_tran, 'ac': _run_ac}


def run(circ, an_list=None, batch=None):
    """ahkab/ahkab.py:677-720 with `batch=`.  Returns {'op'|'dc'|'tran'|'ac':
    solution}.  String x0 labels ('op', 'op+ic') resolve only if an OP ran
    earlier in the same list, otherwise they become None (ahkab.py:710-719)."""
    res = {}
    if not an_list:
        an_list = _queue
    else:
        an_list = copy.deepcopy(an_list)
        if type(an_list) == tuple:
            an_list = list(an_list)
        elif type(an_list) == dict:
            an_list = [an_list]
    eng = get_engine(circ, batch)
    x0s = {}
    while len(an_list):
        an_item = an_list.pop(0)
        an_type = an_item.pop('type')
        if 'x0' in an_item and isinstance(an_item['x0'], str):
            label = an_item['x0']
            if label == 'op+ic' and 'op' in x0s and 'op+ic' not in x0s:
                # (built when first asked for: a clone + scatter on the device per OP otherwise)
                x0s['op+ic'] = icmodified_x0(circ, x0s['op'], batch)
                _x0s.update(x0s)
            an_item['x0'] = x0s.get(label)
        r = _analysis[an_type](circ, eng, **an_item)
        res.update({an_type: r})
        if an_type == 'op':
            x0s = {'op': r}
            _x0s.pop('op+ic', None)
            _x0s.update(x0s)
    return res
