"""CPU: the C-ABI library loads and exports every symbol include/ahkab_b200.h
declares; struct layouts of the ctypes mirror match the header; the product has
no CPU fallback (no compute call is made here)."""
import ctypes as C
import os
import re

import pytest

from ahkab_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'ahkab_b200.h')


def _declared():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ahk_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    lib = _abi.load_library()
    names = _declared()
    assert len(names) >= 12
    for name in names:
        assert getattr(lib, name) is not None, name
    assert sorted(_abi.EXPORTS) == names      # the Python mirror lists the same set


def test_struct_layouts_match_header():
    assert C.sizeof(_abi.ahk_elem) == 14 * 4
    assert C.sizeof(_abi.ahk_options) == 9 * 8 + 12 * 4
    assert C.sizeof(_abi.ahk_batch) == 8 + 4 + 4 + 8 + 8     # B, n_vary, pad, 2 pointers
    assert _abi.ahk_circuit.elems.offset == 16
    hdr = open(HEADER).read()
    assert '#define AHK_ABI_VERSION %d' % _abi.ABI_VERSION in hdr
    for name, val in (('AHK_ST_OK', _abi.ST_OK), ('AHK_ST_GMIN_CHECK', _abi.ST_GMIN_CHECK),
                      ('AHK_ST_PASS2_FAIL', _abi.ST_PASS2_FAIL), ('AHK_ST_HMIN', _abi.ST_HMIN),
                      ('AHK_ST_ROWS_FULL', _abi.ST_ROWS_FULL), ('AHK_ST_SINGULAR', _abi.ST_SINGULAR),
                      ('AHK_ST_NR_FAIL', _abi.ST_NR_FAIL)):
        m = re.search(r'#define %s\s+(0x[0-9a-fA-F]+)' % name, hdr)
        assert m and int(m.group(1), 16) == val, name


def test_kernel_variant_table():
    real = _abi.variants('real')
    ac = _abi.variants('ac')
    assert any(NC == 0 for _i, NC, _r, _g in real) and any(NC == 0 for _i, NC, _r, _g in ac)
    for _i, NC, R, G in real + ac:
        assert G in (1, 2, 4, 8, 16, 32)
        if NC:
            assert R * G >= 1 and NC >= 2
    # every n up to 64 has at least one kernel; the five workloads have a register variant
    for n in range(1, 65):
        assert _abi.fitting_variants(n, 'real') and _abi.fitting_variants(n, 'ac')
    table = {i: NC for i, NC, _r, _g in real}
    for n in (3, 5, 9, 14):
        assert any(table[i] > 0 for i in _abi.fitting_variants(n, 'real'))


def test_no_cpu_fallback():
    """Without a CUDA device the engine refuses to construct (and create fails)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ahkab_b200 import workloads as W
    from ahkab_b200.circuit import Circuit
    from ahkab_b200.engine import Engine, EngineError
    w = W.c2_diode_op(4)
    with pytest.raises(EngineError):
        Engine(w.circuit(Circuit), w.batch())
    # the C-ABI itself also refuses: ahk_create fails with no device
    from ahkab_b200 import compile as cp, options
    flat = cp.flatten(w.circuit(Circuit), w.batch())
    desc = _abi.CircuitDesc(flat)
    opts = _abi.make_options(options.snapshot())
    h = C.c_void_p()
    lib = _abi.load_library()
    assert lib.ahk_create(C.byref(desc.c), C.byref(opts), C.byref(h)) != 0
    assert b'no CUDA device' in lib.ahk_last_error()


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, 'ahkab_b200')
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.hpp', '.h', '.sh')):
                txt = open(os.path.join(dirpath, f), errors='ignore').read()
                assert 'ahk_oracle' not in txt and 'import oracle' not in txt and \
                    'hostemu' not in txt.replace('tests/hostemu', ''), os.path.join(dirpath, f)
