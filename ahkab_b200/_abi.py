"""ctypes view of include/ahkab_b200.h and loader of the CUDA library.

`libahkab_b200.so` is built in-tree by `__graft_entry__.build()` /
`ahkab_b200/csrc/build.sh` (nvcc, sm_100a).  There is NO CPU fallback: if the
library is missing or has no usable GPU the engine raises."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'lib', 'libahkab_b200.so')


class ahk_elem(C.Structure):
    _fields_ = [(k, C.c_int32) for k in (
        'type', 'n1', 'n2', 'n3', 'n4', 'vde', 'ref', 'ref2', 'aux0', 'aux1',
        'pbase', 'mbase', 'flags', 'tf')]


class ahk_circuit(C.Structure):
    _fields_ = [('abi_version', C.c_int32), ('n_nodes', C.c_int32),
                ('n_vde', C.c_int32), ('n_elems', C.c_int32),
                ('elems', C.POINTER(ahk_elem)),
                ('n_params', C.c_int32),
                ('nominal', C.POINTER(C.c_double)),
                ('n_guess', C.c_int32),
                ('guess_G', C.POINTER(C.c_double)),
                ('guess_slot', C.POINTER(C.c_int32)),
                ('guess_scale', C.POINTER(C.c_double)),
                ('guess_const', C.POINTER(C.c_double))]


class ahk_options(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        'vea', 'ver', 'iea', 'ier', 'gmin', 'cmin', 'hmin',
        'nl_voltages_lock_factor', 'transient_aposteriori_step_threshold')] + \
        [(k, C.c_int32) for k in (
            'nl_voltages_lock', 'nr_damp_first_iters', 'dc_max_nr_iter',
            'transient_max_nr_iter', 'ac_max_nr_iter',
            'use_standard_solve_method', 'use_gmin_stepping',
            'use_source_stepping', 'transient_prediction_as_x0',
            'transient_use_aposteriori_step_control',
            'transient_max_time_iter', 'reserved')]


class ahk_batch(C.Structure):
    _fields_ = [('B', C.c_int64), ('n_vary', C.c_int32),
                ('vary_slots', C.POINTER(C.c_int32)),
                ('vary', C.c_void_p)]


ABI_VERSION = 1

# status bits
ST_OK, ST_GMIN_CHECK, ST_PASS2_FAIL, ST_HMIN, ST_ROWS_FULL, ST_SINGULAR, \
    ST_NR_FAIL = 0x01, 0x08, 0x10, 0x20, 0x40, 0x80, 0x100
ST_METHOD_MASK, ST_METHOD_SHIFT = 0x06, 1

METHODS = {'IMPLICIT_EULER': 0, 'TRAP': 1, 'GEAR1': 11, 'GEAR2': 12,
           'GEAR3': 13, 'GEAR4': 14, 'GEAR5': 15, 'GEAR6': 16}

EXPORTS = ('ahk_create', 'ahk_destroy', 'ahk_size', 'ahk_set_options',
           'ahk_set_group', 'ahk_last_error', 'ahk_launch_count', 'ahk_op',
           'ahk_dc_sweep', 'ahk_tran', 'ahk_ac', 'ahk_probe_fp64',
           'ahk_set_variant', 'ahk_variant_count', 'ahk_variant_info',
           'ahk_h2d_columns', 'ahk_set_jit', 'ahk_jit_launch_count',
           'ahk_jit_build_count', 'ahk_jit_compile', 'ahk_op_host',
           'ahk_set_jit_shape', 'ahk_set_jit_tuning', 'ahk_set_jit_slu', 'ahk_jit_slu_stats', 'ahk_last_launch', 'ahk_pack_rows', 'ahk_big_stats')


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class CircuitDesc(object):
    """Owns the numpy buffers an `ahk_circuit` points into."""

    def __init__(self, flat):
        self._keep = []

        def keep(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            self._keep.append(a)
            return a
        el = np.ascontiguousarray(flat.elems)
        self._keep.append(el)
        c = ahk_circuit()
        c.abi_version = ABI_VERSION
        c.n_nodes, c.n_vde, c.n_elems = flat.n_nodes, flat.n_vde, len(el)
        c.elems = el.ctypes.data_as(C.POINTER(ahk_elem))
        c.n_params = len(flat.nominal)
        c.nominal = _dp(keep(flat.nominal, np.float64))
        c.n_guess = flat.n_guess
        c.guess_G = _dp(keep(flat.guess_G, np.float64))
        c.guess_slot = _ip(keep(flat.guess_slot, np.int32))
        c.guess_scale = _dp(keep(flat.guess_scale, np.float64))
        c.guess_const = _dp(keep(flat.guess_const, np.float64))
        self.c = c
        self.n = flat.n


def make_options(d):
    o = ahk_options()
    for k, _t in ahk_options._fields_:
        if k == 'reserved':
            continue
        setattr(o, k, d[k])
    return o


_lib = None


def load_library(path=None):
    """Load libahkab_b200.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(
            "CUDA library %s not found: run `python -c 'import "
            "__graft_entry__ as g; g.build()'` (nvcc, sm_100a). There is no "
            "CPU fallback." % path)
    lib = C.CDLL(path)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    lib.ahk_create.argtypes = [C.POINTER(ahk_circuit), C.POINTER(ahk_options),
                               C.POINTER(vp)]
    lib.ahk_create.restype = C.c_int
    lib.ahk_destroy.argtypes = [vp]
    lib.ahk_destroy.restype = None
    lib.ahk_size.argtypes = [vp]
    lib.ahk_size.restype = C.c_int
    lib.ahk_set_options.argtypes = [vp, C.POINTER(ahk_options)]
    lib.ahk_set_options.restype = C.c_int
    lib.ahk_set_group.argtypes = [vp, C.c_int]
    lib.ahk_set_group.restype = C.c_int
    lib.ahk_last_error.argtypes = []
    lib.ahk_last_error.restype = C.c_char_p
    lib.ahk_launch_count.argtypes = []
    lib.ahk_launch_count.restype = i64
    bp = C.POINTER(ahk_batch)
    lib.ahk_op.argtypes = [vp, bp, vp, C.c_int, vp, vp, vp, vp, vp]
    lib.ahk_op.restype = C.c_int
    lib.ahk_dc_sweep.argtypes = [vp, bp, C.c_int, C.POINTER(dbl), C.c_int, vp,
                                 C.c_int, vp, vp, vp, vp]
    lib.ahk_dc_sweep.restype = C.c_int
    lib.ahk_tran.argtypes = [vp, bp, vp, dbl, dbl, dbl, C.c_int, C.c_int, i64,
                             vp, vp, vp, vp, vp, vp]
    lib.ahk_tran.restype = C.c_int
    lib.ahk_ac.argtypes = [vp, bp, vp, C.POINTER(dbl), C.c_int, vp, vp, vp]
    lib.ahk_ac.restype = C.c_int
    lib.ahk_set_variant.argtypes = [vp, C.c_int, C.c_int]
    lib.ahk_set_variant.restype = C.c_int
    lib.ahk_variant_count.argtypes = [C.c_int]
    lib.ahk_variant_count.restype = C.c_int
    lib.ahk_variant_info.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int32)]
    lib.ahk_variant_info.restype = C.c_int
    lib.ahk_h2d_columns.argtypes = [vp, vp, C.c_int, i64, i64, i64, vp]
    lib.ahk_h2d_columns.restype = C.c_int
    lib.ahk_probe_fp64.argtypes = [C.c_int, C.c_int, C.c_int, vp, vp]
    lib.ahk_probe_fp64.restype = C.c_int
    lib.ahk_op_host.argtypes = [vp, i64, C.c_int, C.POINTER(i32), vp, C.c_int, vp, vp, vp,
                                vp, C.c_int, vp]
    lib.ahk_op_host.restype = C.c_int
    lib.ahk_set_jit.argtypes = [vp, C.c_int, i64]
    lib.ahk_set_jit.restype = C.c_int
    lib.ahk_pack_rows.argtypes = [vp, vp, vp, vp, i64, i64, C.c_int, vp, vp, vp]
    lib.ahk_pack_rows.restype = C.c_int
    lib.ahk_set_jit_shape.argtypes = [vp, C.c_int, C.c_int]
    lib.ahk_set_jit_shape.restype = C.c_int
    lib.ahk_last_launch.argtypes = [vp, C.POINTER(C.c_int32)]
    lib.ahk_last_launch.restype = C.c_int
    lib.ahk_big_stats.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.ahk_big_stats.restype = C.c_int
    lib.ahk_set_jit_tuning.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    lib.ahk_set_jit_tuning.restype = C.c_int
    lib.ahk_set_jit_slu.argtypes = [vp, C.c_int]
    lib.ahk_set_jit_slu.restype = C.c_int
    lib.ahk_jit_slu_stats.argtypes = [vp, C.POINTER(i32)]
    lib.ahk_jit_slu_stats.restype = C.c_int
    lib.ahk_jit_launch_count.argtypes = []
    lib.ahk_jit_launch_count.restype = i64
    lib.ahk_jit_build_count.argtypes = []
    lib.ahk_jit_build_count.restype = i64
    lib.ahk_jit_compile.argtypes = [C.POINTER(ahk_circuit), C.POINTER(ahk_options),
                                    C.c_int, C.POINTER(i32), C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.POINTER(i64)]
    lib.ahk_jit_compile.restype = C.c_int
    if path == LIB_PATH:
        _lib = lib
    return lib


def variants(kind='real', lib=None):
    """[(id, NC, R, G)] of the compiled kernel variants; kind = 'real' (OP / DC /
    transient) or 'ac'.  NC = 0 is the generic shared-memory fallback."""
    lib = lib or load_library()
    k = 1 if kind == 'ac' else 0
    out = []
    buf = (C.c_int32 * 3)()
    for i in range(lib.ahk_variant_count(k)):
        lib.ahk_variant_info(k, i, buf)
        out.append((i, int(buf[0]), int(buf[1]), int(buf[2])))
    return out


def fitting_variants(n, kind='real', lib=None):
    """ids of the variants that can run a circuit with n unknowns."""
    return [i for i, NC, R, G in variants(kind, lib)
            if (NC == 0 and (n <= 64 or (G == 32 and n <= 512))) or (NC - 1 >= n and R * G >= n)]
