"""TEST INFRASTRUCTURE — ctypes wrapper of oracle/_build/libahk_oracle.so, the
scalar C restatement of the reference's algorithm (oracle/ahk_oracle.c).
numpy in / numpy out, same argument meaning as the C-ABI entry points of the
product so that parity tests compare like with like.  Never imported by the
product package (ahkab_b200/)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from ahkab_b200 import _abi, options as _options  # noqa: E402 (struct layouts)
from ahkab_b200 import compile as _compile        # noqa: E402

LIB = os.path.join(HERE, '_build', 'libahk_oracle.so')
_lib = None


def build():
    subprocess.check_call(['sh', os.path.join(HERE, 'build.sh')],
                          stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.ahko_max_threads.restype = C.c_int
        L.ahko_tf_eval.restype = C.c_double
        L.ahko_tf_eval.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_double]
        _lib = L
    return _lib


def max_threads():
    return lib().ahko_max_threads()


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


class Oracle(object):
    """One topology (+ optional batch), built from our Circuit/ParamBatch."""

    def __init__(self, circ, batch=None, opts=None):
        self.flat = _compile.flatten(circ, batch)
        self.desc = _abi.CircuitDesc(self.flat)
        o = _options.snapshot()
        if opts:
            o.update(opts)
        self.opts = _abi.make_options(o)
        self.n = self.flat.n
        self.B = self.flat.B if batch is not None else 1
        self.vary = np.ascontiguousarray(self.flat.vary, dtype=np.float64)
        self.slots = np.ascontiguousarray(self.flat.vary_slots, np.int32)

    def _common(self):
        return (C.byref(self.desc.c), C.byref(self.opts), C.c_int64(self.B),
                C.c_int(len(self.slots)), _ip(self.slots), _dp(self.vary))

    def op(self, x0=None, use_guess=True, threads=0):
        B, n = self.B, self.n
        x = np.zeros((B, n)); res = np.zeros((B, n))
        it = np.zeros(B, np.int32); st = np.zeros(B, np.int32)
        x0 = None if x0 is None else np.ascontiguousarray(x0, np.float64)
        lib().ahko_op(*self._common(), _dp(x0), C.c_int(bool(use_guess)),
                      _dp(x), _dp(res), _ip(it), _ip(st), C.c_int(threads))
        return dict(x=x, resid=res, iters=it, status=st)

    def dc_sweep(self, src_elem, sweep, x0=None, use_guess=True, threads=0):
        B, n = self.B, self.n
        sweep = np.ascontiguousarray(sweep, np.float64)
        P = len(sweep)
        x = np.zeros((B, P, n))
        it = np.zeros((B, P), np.int32); st = np.zeros((B, P), np.int32)
        x0 = None if x0 is None else np.ascontiguousarray(x0, np.float64)
        lib().ahko_dc_sweep(*self._common(), C.c_int(src_elem), _dp(sweep),
                            C.c_int(P), _dp(x0), C.c_int(bool(use_guess)),
                            _dp(x), _ip(it), _ip(st), C.c_int(threads))
        return dict(x=x, iters=it, status=st)

    def tran(self, x0, tstart, tstep, tstop, method, use_step_control,
             max_rows, threads=0):
        B, n = self.B, self.n
        t = np.full((B, max_rows), np.nan)
        x = np.full((B, max_rows, n), np.nan)
        rows = np.zeros(B, np.int32); cnt = np.zeros((B, 4), np.int32)
        st = np.zeros(B, np.int32)
        x0 = None if x0 is None else np.ascontiguousarray(x0, np.float64)
        lib().ahko_tran(*self._common(), _dp(x0), C.c_double(tstart),
                        C.c_double(tstep), C.c_double(tstop),
                        C.c_int(_abi.METHODS[method.upper()]),
                        C.c_int(bool(use_step_control)), C.c_int64(max_rows),
                        _dp(t), _dp(x), _ip(rows), _ip(cnt), _ip(st),
                        C.c_int(threads))
        return dict(t=t, x=x, rows=rows, counters=cnt, status=st)

    def ac(self, omegas, x_op=None, threads=0, op_first=False, settle=0):
        """op_first: the reference's flow when no linearisation point is given (OP inside, the
        device state of that Newton run carries into J); settle: extra device evaluations at
        the linearisation point before J is taken (see ahko_ac2)."""
        B, n = self.B, self.n
        om = np.ascontiguousarray(omegas, np.float64)
        x = np.zeros((B, len(om), n), np.complex128)
        st = np.zeros(B, np.int32)
        x_op = None if x_op is None else np.ascontiguousarray(x_op, np.float64)
        xo = np.zeros((B, n))
        lib().ahko_ac2(*self._common(), _dp(x_op), _dp(om), C.c_int(len(om)),
                       x.ctypes.data_as(C.POINTER(C.c_double)), _ip(st),
                       C.c_int(threads), C.c_int(int(op_first)), C.c_int(int(settle)), _dp(xo))
        return dict(x=x, status=st, op=xo)

    def ac_op_first(self, omegas, threads=0):
        return self.ac(omegas, threads=threads, op_first=True)
