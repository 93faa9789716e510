"""TEST-ONLY ctypes wrapper of tests/hostemu/_build/libahk_hostemu.so (the CUDA
engine's per-instance source compiled for the host with one lane per instance).
Same call shapes as oracle/oracle.py so tests can diff the two."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

from ahkab_b200 import _abi  # noqa: E402
import oracle as _oracle      # noqa: E402

LIB = os.path.join(HERE, '_build', 'libahk_hostemu.so')
_lib = None


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(HERE, 'hostemu.cpp')] + [
            os.path.join(ROOT, 'ahkab_b200', 'csrc', f)
            for f in os.listdir(os.path.join(ROOT, 'ahkab_b200', 'csrc'))
            if f.endswith(('.cuh', '.hpp', '.h'))]
        if (not os.path.exists(LIB) or
                os.path.getmtime(LIB) < max(os.path.getmtime(s) for s in srcs)):
            subprocess.check_call(['sh', os.path.join(HERE, 'build.sh')],
                                  stdout=subprocess.DEVNULL)
        _lib = C.CDLL(LIB)
    return _lib


_dp, _ip = _oracle._dp, _oracle._ip

# host-testable kernel variants (tests/hostemu/hostemu.cpp)
VARIANTS = {0: 'Var<0,0,1> generic', 1: 'Var<4,3,1>', 2: 'Var<6,5,1>'}


def set_variant(v):
    if lib().emu_set_variant(int(v)) != 0:
        raise ValueError("unknown host variant %r" % (v,))


def _check(rc):
    if rc == -2:
        raise ValueError("circuit does not fit the selected host variant")
    if rc != 0:
        raise RuntimeError("hostemu call failed (%d)" % rc)


class Emu(_oracle.Oracle):
    """Same constructor as Oracle; methods run the emulated engine instead."""
    ac_op_first = None     # (the engine's AC takes the linearisation point from its own OP call)

    def _common6(self):
        return self._common()

    def op(self, x0=None, use_guess=True, threads=0):
        B, n = self.B, self.n
        x = np.zeros((B, n)); res = np.zeros((B, n))
        it = np.zeros(B, np.int32); st = np.zeros(B, np.int32)
        x0 = None if x0 is None else np.ascontiguousarray(x0, np.float64)
        _check(lib().emu_op(*self._common(), _dp(x0), C.c_int(bool(use_guess)),
                     _dp(x), _dp(res), _ip(it), _ip(st)))
        return dict(x=x, resid=res, iters=it, status=st)

    def dc_sweep(self, src_elem, sweep, x0=None, use_guess=True, threads=0):
        B, n = self.B, self.n
        sweep = np.ascontiguousarray(sweep, np.float64)
        P = len(sweep)
        x = np.zeros((B, P, n))
        it = np.zeros((B, P), np.int32); st = np.zeros((B, P), np.int32)
        x0 = None if x0 is None else np.ascontiguousarray(x0, np.float64)
        _check(lib().emu_dc_sweep(*self._common(), C.c_int(src_elem), _dp(sweep),
                           C.c_int(P), _dp(x0), C.c_int(bool(use_guess)),
                           _dp(x), _ip(it), _ip(st)))
        return dict(x=x, iters=it, status=st)

    def tran(self, x0, tstart, tstep, tstop, method, use_step_control,
             max_rows, threads=0):
        B, n = self.B, self.n
        t = np.full((B, max_rows), np.nan)
        x = np.full((B, max_rows, n), np.nan)
        rows = np.zeros(B, np.int32); cnt = np.zeros((B, 4), np.int32)
        st = np.zeros(B, np.int32)
        x0 = None if x0 is None else np.ascontiguousarray(x0, np.float64)
        _check(lib().emu_tran(*self._common(), _dp(x0), C.c_double(tstart),
                       C.c_double(tstep), C.c_double(tstop),
                       C.c_int(_abi.METHODS[method.upper()]),
                       C.c_int(bool(use_step_control)), C.c_int64(max_rows),
                       _dp(t), _dp(x), _ip(rows), _ip(cnt), _ip(st)))
        return dict(t=t, x=x, rows=rows, counters=cnt, status=st)

    def ac(self, omegas, x_op=None, threads=0):
        B, n = self.B, self.n
        om = np.ascontiguousarray(omegas, np.float64)
        x = np.zeros((B, len(om), n), np.complex128)
        st = np.zeros(B, np.int32)
        x_op = None if x_op is None else np.ascontiguousarray(x_op, np.float64)
        _check(lib().emu_ac(*self._common(), _dp(x_op), _dp(om), C.c_int(len(om)),
                     x.ctypes.data_as(C.POINTER(C.c_double)), _ip(st)))
        return dict(x=x, status=st)

    def jit_source(self, mode, method='TRAP', use_step_control=False, NC=4, R=3, G=1,
                   threads=128, min_blocks=4, ilp=1):
        """Text of the generated jit_spec.h (ahkab_b200/csrc/jit_gen.hpp) for this batch."""
        c6 = self._common()
        args = (c6[0], c6[1], c6[3], c6[4], C.c_int(mode),
                C.c_int(_abi.METHODS[method.upper()]), C.c_int(bool(use_step_control)),
                C.c_int(NC), C.c_int(R), C.c_int(G), C.c_int(threads), C.c_int(min_blocks))
        n = lib().emu_jit_source(*args, None, C.c_int64(0), C.c_int(ilp))
        if n < 0:
            raise RuntimeError("emu_jit_source failed")
        buf = C.create_string_buffer(n + 1)
        lib().emu_jit_source(*args, buf, C.c_int64(n + 1), C.c_int(ilp))
        return buf.value.decode()

    def big3_sweep(self, src_elem, sweep):
        """Large linear circuits: the static-schedule path (biglin3.cuh) on the host.
        Returns dict(x, resid, iters, status, flags, dynamic) — `dynamic` = instances whose
        pivot sequence differs from the nominal one (-1: no schedule could be built)."""
        B, n = self.B, self.n
        sweep = np.ascontiguousarray(sweep, np.float64)
        P = len(sweep)
        x = np.zeros((B, P, n)); res = np.zeros((B, P, n))
        it = np.zeros((B, P), np.int32); st = np.zeros((B, P), np.int32)
        fl = np.zeros(B, np.int32)
        dyn = lib().emu_big3_sweep(*self._common(), C.c_int(src_elem), _dp(sweep), C.c_int(P),
                                   _dp(x), _dp(res), _ip(it), _ip(st), _ip(fl))
        return dict(x=x, resid=res, iters=it, status=st, flags=fl, dynamic=dyn)
