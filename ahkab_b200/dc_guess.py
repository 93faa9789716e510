"""OP initial guess as a linear operator — host side, once per topology.

Mirror of ahkab/dc_guess.py:40-227.  The reference builds a port-incidence
matrix M (one row per non-degenerate port of every element carrying a
`dc_guess` hint), drops linearly dependent rows and empty columns, and solves
``Rp = pinv(M) . T`` (``inv`` when square).  M depends on the topology only; T
holds the hints, which depend on per-instance parameters (a V source's
dc_value, 0.1*VTO*NP of an EKV device, ...).  So the host computes the operator
G (x0 = G . T) once with numpy — exactly the reference's calls — and the device
evaluates T per instance from the parameter table (include/ahkab_b200.h:
ahk_circuit.guess_*)."""
import numpy as np

from . import circuit as _c


def _remove_row(M, r):
    return np.delete(M, r, axis=0)


def build_guess_operator(circ, hint_rows):
    """hint_rows: list of (n1, n2, slot, scale, const) in the reference's row
    order (see compile.flatten).  Returns (G [n, n_T], kept_rows) or
    (None, None) when the reference would return no guess."""
    if not circ.is_nonlinear() or not hint_rows:
        return None, None
    nv = circ.get_nodes_number()
    v_eq = sum(1 for e in circ if _c.is_elem_voltage_defined(e))
    M = np.zeros((len(hint_rows), nv))
    for i, (n1, n2, _s, _k, _c0) in enumerate(hint_rows):
        M[i, n1] = +1
        M[i, n2] = -1
    kept = list(range(len(hint_rows)))
    # remove the ground column (dc_guess.py:151)
    M = np.delete(M, 0, axis=1)
    # remove linearly dependent rows (dc_guess.py:160-168)
    for i in range(M.shape[0] - 1, -1, -1):
        for j in range(i - 1, -1, -1):
            d1 = M[i, :] - M[j, :]
            d2 = M[i, :] + M[j, :]
            if not d1.any() or not d2.any():
                M = _remove_row(M, i)
                del kept[i]
                break
    # remove empty columns (dc_guess.py:176-181)
    removed_index = []
    for i in range(M.shape[1] - 1, -1, -1):
        if not M[:, i].any():
            M = np.delete(M, i, axis=1)
            removed_index.append(i)
    if M.shape[0] != M.shape[1]:
        Op = np.linalg.pinv(M)
    else:
        if np.linalg.det(M) != 0:
            try:
                Op = np.linalg.inv(M)
            except np.linalg.LinAlgError:
                return None, None
        else:
            return None, None
    # re-insert zero rows for the removed columns, in the reference's order
    # (dc_guess.py:207-211)
    for index in removed_index:
        Op = np.concatenate((Op[:index, :], np.zeros((1, Op.shape[1])),
                             Op[index:, :]), axis=0)
    if v_eq > 0:
        Op = np.concatenate((Op, np.zeros((v_eq, Op.shape[1]))), axis=0)
    return np.ascontiguousarray(Op), kept
