"""TEST INFRASTRUCTURE — generates tests/golden/zoo.npz by running the LIVE, unmodified
reference (/root/reference through oracle/refshim.py) on the circuits of tests/zoo.py.
Authoring container only:   python oracle/make_golden.py misc"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
GOLD = os.path.join(ROOT, 'tests', 'golden')


def run_one(ahkab, name, spec):
    from ahkab import time_functions as tf
    circ = spec['build'](ahkab.Circuit, tf)
    an = dict(spec['an'])
    kind = an.pop('kind')
    out = {}
    if kind == 'op':
        r = ahkab.run(circ, ahkab.new_op(verbose=0))['op']
        out['x'] = r.asarray().ravel()
        out['iters'] = r.iterations
        out['keys'] = np.array(r.keys())
    elif kind == 'dc':
        r = ahkab.run(circ, ahkab.new_dc(verbose=0, **an))['dc']
        a = r.asarray()
        out['sweep'], out['x'] = a[0], a[1:].T
        out['keys'] = np.array(r.keys())
    elif kind == 'ac':
        # x0='op' (non-linear circuits): the reference computes the linearisation point itself
        # when none is given (ac.py:237-262)
        # when none is given (ac.py:237-262).  The EKV device objects carry the warm start of
        # their inner solve from call to call (ekv.py:575-586), so the AC run gets a circuit
        # whose devices have seen nothing but that one op_analysis.
        if an.pop('x0', None) == 'op':
            out['op'] = ahkab.run(circ, ahkab.new_op(verbose=0))['op'].asarray().ravel()
            circ = spec['build'](ahkab.Circuit, tf)
        r = ahkab.run(circ, ahkab.new_ac(x0=None, verbose=0, **an))['ac']
        a = r.asarray()
        out['f'], out['x'] = a[0].real, a[1:].T
        out['keys'] = np.array(r.keys())
    elif kind == 'tran':
        x0 = an.pop('x0')
        ans = [ahkab.new_op(verbose=0), ahkab.new_tran(x0=x0, verbose=0, **an)]
        r = ahkab.run(circ, ans)
        a = r['tran'].asarray()
        out['op'] = r['op'].asarray().ravel()
        out['t'], out['x'] = a[0], a[1:].T
        out['keys'] = np.array(r['tran'].keys())
    return out


def main():
    import warnings
    warnings.simplefilter('ignore')
    import refshim
    ahkab = refshim.load()
    # mosq's OP *reporting* (not the solve) raises under numpy 2 (mosq.py:254 compares a
    # ragged tuple): replace the report of that one device class; the Newton
    # path is untouched.
    from ahkab import mosq
    mosq.mosq_device.get_op_info = lambda self, ports_v: ([], [])
    import zoo
    out = {}
    for name, spec in zoo.ZOO.items():
        res = run_one(ahkab, name, spec)
        for k, v in res.items():
            out['%s__%s' % (name, k)] = np.asarray(v)
        print(name, {k: np.asarray(v).shape for k, v in res.items()})
    np.savez_compressed(os.path.join(GOLD, 'zoo.npz'), **out)


if __name__ == '__main__':
    main()
