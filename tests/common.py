"""Shared helpers for the parity tests (tests only)."""
import numpy as np

from ahkab_b200 import workloads as W
from ahkab_b200.circuit import Circuit
from ahkab_b200 import utilities


def parity_tol(a, b, nv1, vea=1e-6, iea=1e-9, rel=1e-3):
    """SURVEY §8(d) tolerance: |dV| <= vea + rel|V|, |dI| <= iea + rel|I| on
    the last axis (unknowns).  Returns the worst ratio err / allowed."""
    a, b = np.asarray(a), np.asarray(b)
    n = a.shape[-1]
    atol = np.where(np.arange(n) < nv1, vea, iea)
    allowed = atol + rel * np.abs(b)
    with np.errstate(invalid='ignore'):
        r = np.abs(a - b) / allowed
    return float(np.nanmax(r))


def c1_omegas():
    return np.array(list(utilities.log_axis_iterator(
        2 * np.pi * W.C1_AC['start'], 2 * np.pi * W.C1_AC['stop'],
        W.C1_AC['points'])))


def c5_sweep():
    return np.array(list(utilities.lin_axis_iterator(
        W.C5_DC['start'], W.C5_DC['stop'], W.C5_DC['points'])))


def c3_x0(x_op):
    """icmodified_x0 for tests/colpitts: c1 (nd, ns) ic=2.5, l1 ic=-1n."""
    x0 = np.array(x_op, copy=True)
    x0[:, 1] = x0[:, 2] + 2.5
    x0[:, 6] = -1e-9
    return x0


def c4_x0(circ, B):
    x0 = np.zeros((B, 5))
    for lab, val in W.C4_IC.items():
        x0[:, circ.ext_node_to_int(lab[2:-1]) - 1] = val
    return x0
