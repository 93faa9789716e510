"""TEST INFRASTRUCTURE — loader for the *live* reference (authoring container
only: /root/reference does not exist on the GPU box).

Imports the untouched ahkab package from /root/reference after installing the
in-process compatibility shims listed in SURVEY.md §8(c) (removed `imp`,
`scipy.misc.factorial`, `np.complex`, ...).  Used only by
oracle/make_golden.py to generate tests/golden/*.npz and to pin the C
restatement (oracle/ahk_oracle.c).  Never imported by the product package."""
import os
import sys
import types

REF = os.environ.get('AHKAB_REFERENCE', '/root/reference')


def load():
    if 'ahkab' in sys.modules and hasattr(sys.modules['ahkab'], 'run'):
        return sys.modules['ahkab']
    if not os.path.isdir(os.path.join(REF, 'ahkab')):
        raise RuntimeError("reference not available at %s" % REF)
    import numpy as np
    import scipy
    import scipy.special
    import scipy.signal
    import scipy.signal.windows as _w
    if 'imp' not in sys.modules:
        sys.modules['imp'] = types.ModuleType('imp')
    if not hasattr(scipy, 'misc') or not hasattr(scipy.misc, 'factorial'):
        m = types.ModuleType('scipy.misc')
        m.factorial = scipy.special.factorial
        scipy.misc = m
        sys.modules['scipy.misc'] = m
    for nm in ('bartlett', 'hann', 'hamming', 'blackman', 'blackmanharris',
               'gaussian', 'kaiser', 'boxcar'):
        if not hasattr(scipy.signal, nm) and hasattr(_w, nm):
            setattr(scipy.signal, nm, getattr(_w, nm))
    if not hasattr(np, 'complex'):
        np.complex = complex
    if not hasattr(np, 'complex_'):
        np.complex_ = np.complex128
    if not hasattr(np, 'alltrue'):
        np.alltrue = np.all
    if not hasattr(np.linalg, 'linalg'):
        np.linalg.linalg = np.linalg
    if 'matplotlib' not in sys.modules:
        try:
            import matplotlib  # noqa
        except Exception:
            pass
    sys.path.insert(0, REF)
    try:
        import ahkab
    finally:
        sys.path.remove(REF)
    return ahkab
