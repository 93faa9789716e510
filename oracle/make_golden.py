"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the LIVE,
unmodified reference (/root/reference, through oracle/refshim.py) on the
first K instances of each BASELINE workload and on a few of the reference's
own test circuits.  Run in the authoring container only:

    python oracle/make_golden.py [c1 c2 c3 c4 c5 dev misc]

The fixtures pin the C restatement (oracle/ahk_oracle.c) and, through it and
directly, the CUDA engine.  Also records the measured wall time of the
reference's Python loop per instance (1 core) for DESIGN.md/BASELINE context.
"""
import os
import sys
import time
import multiprocessing as mp

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, 'tests', 'golden')

K = dict(c1=16, c2=512, c3=8, c4=8, c5=2)


def _ref():
    import warnings
    warnings.simplefilter('ignore')
    import refshim
    return refshim.load()


class _Counter(object):
    """Counts dc_solve calls / NR iterations by wrapping (not editing) the
    reference's dc_analysis.dc_solve."""
    def __init__(self, ahkab):
        self.calls = 0
        self.iters = 0
        self.dca = ahkab.dc_analysis
        self.orig = self.dca.dc_solve

    def __enter__(self):
        def wrapped(*a, **kw):
            r = self.orig(*a, **kw)
            self.calls += 1
            self.iters += r[3]
            return r
        self.dca.dc_solve = wrapped
        return self

    def __exit__(self, *a):
        self.dca.dc_solve = self.orig


def _set_opts(ahkab, opts):
    saved = {k: getattr(ahkab.options, k) for k in opts}
    for k, v in opts.items():
        setattr(ahkab.options, k, v)
    return saved


def _one(args):
    name, i = args
    ahkab = _ref()
    from ahkab_b200 import workloads as W
    w = W.ALL[name](B=K[name])
    saved = _set_opts(ahkab, w.opts)
    out = {}
    try:
        t0 = time.perf_counter()
        circ = w.build(ahkab.Circuit, w.instance_params(i))
        with _Counter(ahkab) as cnt:
            if name == 'c1':
                r = ahkab.run(circ, ahkab.new_ac(x0=None, verbose=0,
                                                 **W.C1_AC))['ac']
                a = r.asarray()
                out = dict(f=a[0].real, x=a[1:].T)
            elif name == 'c2':
                r = ahkab.run(circ, ahkab.new_op(verbose=0))['op']
                out = dict(x=r.asarray().ravel(), iters=r.iterations)
            elif name == 'c3':
                an = [ahkab.new_op(verbose=0),
                      ahkab.new_tran(x0='op+ic', verbose=0, **W.C3_TRAN)]
                r = ahkab.run(circ, an)
                a = r['tran'].asarray()
                out = dict(op=r['op'].asarray().ravel(),
                           op_iters=r['op'].iterations, t=a[0], x=a[1:].T)
            elif name == 'c4':
                x0 = ahkab.new_x0(circ, W.C4_IC)
                r = ahkab.run(circ, ahkab.new_tran(x0=x0, verbose=0,
                                                   **W.C4_TRAN))['tran']
                a = r.asarray()
                out = dict(t=a[0], x=a[1:].T)
            elif name == 'c5':
                r = ahkab.run(circ, ahkab.new_dc(verbose=0, **W.C5_DC))['dc']
                a = r.asarray()
                out = dict(sweep=a[0], x=a[1:].T)
        out['dc_solves'] = cnt.calls
        out['nr_iters'] = cnt.iters
        out['wall'] = time.perf_counter() - t0
    finally:
        _set_opts(ahkab, saved)
    return out


def gen_workload(name):
    n = K[name]
    with mp.Pool(min(8, n)) as pool:
        res = pool.map(_one, [(name, i) for i in range(n)], chunksize=1)
    keys = res[0].keys()
    out = {}
    for k in keys:
        vals = [np.asarray(r[k]) for r in res]
        if vals[0].ndim >= 1 and len(set(v.shape for v in vals)) > 1:
            # ragged (LTE-controlled transient): pad, keep row counts
            m = max(v.shape[0] for v in vals)
            pad = np.full((n, m) + vals[0].shape[1:], np.nan)
            for j, v in enumerate(vals):
                pad[j, :v.shape[0]] = v
            out[k] = pad
            out[k + '_rows'] = np.asarray([v.shape[0] for v in vals])
        else:
            out[k] = np.stack(vals)
    out['K'] = n
    np.savez_compressed(os.path.join(GOLD, name + '.npz'), **out)
    print(name, {k: getattr(v, 'shape', v) for k, v in out.items()},
          'mean wall/instance %.4f s' % out['wall'].mean())


# ---- single-device evaluations (SURVEY Appendix A2 style) ------------------
def gen_devices():
    ahkab = _ref()
    from ahkab import ekv, mosq, diode, switch
    rng = np.random.default_rng(7)
    out = {}
    # EKV: fresh device per line, call order i, g0, g1, g2
    rows = []
    for TYPE, VTO, KP, W, L in (('n', .4, 10e-6, 200e-6, 1e-6),
                                ('n', .4, 40e-6, 1e-6, 1e-6),
                                ('p', -.4, 12e-6, 3e-6, 1e-6)):
        for _ in range(40):
            v = rng.uniform(-0.5, 3.0, 3)
            if TYPE == 'p':
                v = -v
            if rng.uniform() < 0.15:
                v[2] = 0.0
            m = ekv.ekv_mos_model(TYPE=TYPE, VTO=VTO, KP=KP)
            d = ekv.ekv_device('m1', 1, 2, 3, 4, W, L, m)
            pv = tuple(float(t) for t in v)
            i = d.i(0, pv)
            g0, g1, g2 = d.g(0, pv, 0), d.g(0, pv, 1), d.g(0, pv, 2)
            rows.append([1 if TYPE == 'n' else -1, VTO, KP, W, L, *pv, i, g0,
                         g1, g2, d.opdict['ifn'], d.opdict['irn']])
    out['ekv'] = np.asarray(rows)
    # mosq
    rows = []
    for TYPE, VTO, KP, W, L in (('n', .4, 40e-6, 1e-6, 1e-6),
                                ('p', -.4, 12e-6, 3e-6, 1e-6)):
        for _ in range(40):
            v = rng.uniform(-1.0, 3.0, 3)
            v[2] = rng.uniform(-1.0, 0.5)
            if TYPE == 'p':
                v = -v
            m = mosq.mosq_mos_model(TYPE=TYPE, VTO=VTO, KP=KP)
            d = mosq.mosq_device('m1', 1, 2, 3, 4, W, L, m)
            pv = tuple(np.float64(t) for t in v)   # mosq.py:382 needs numpy scalars
            _ix, ist = d.istamp(pv, reduced=False)
            _ix, gst = d.gstamp(pv, reduced=False)
            rows.append([1 if TYPE == 'n' else -1, VTO, KP, W, L, *pv,
                         ist[0], *np.asarray(gst).reshape(-1)])
    out['mosq'] = np.asarray(rows)
    # diode (RS = 0 and RS != 0, BV finite)
    rows = []
    for kw in (dict(), dict(IS=2e-14, N=1.5, ISR=1e-12, BV=5.0, IKF=1e-2),
               dict(RS=2.0)):
        for _ in range(30):
            v = float(rng.uniform(-6, 1.2))
            m = diode.diode_model('d', **kw)
            d = diode.diode('d1', 1, 2, m, AREA=1.5)
            i = d.i(0, (v,))
            g = d.g(0, (v,), 0)
            mm = dict(IS=m.IS, N=m.N, NBV=m.NBV, ISR=m.ISR, NR=m.NR, RS=m.RS,
                      BV=m.BV, IKF=m.IKF)
            rows.append([mm['IS'], mm['N'], mm['NBV'], mm['ISR'], mm['NR'],
                         mm['RS'], mm['BV'], mm['IKF'], 1.5, v, i, g])
    out['diode'] = np.asarray(rows)
    # switch
    rows = []
    for _ in range(40):
        m = switch.vswitch_model('sw', VT=1.0, VH=0.2, RON=10., ROFF=1e6)
        on0 = bool(rng.uniform() < 0.5)
        d = switch.switch_device(1, 2, 3, 4, m, ic=on0)
        v = (float(rng.uniform(-1, 1)), float(rng.uniform(0, 2)))
        i = d.i(0, v)
        go = d.g(0, v, 0)
        gm = d.g(0, v, 1)
        rows.append([1.0, 0.2, 10., 1e6, float(on0), *v, i, go, gm,
                     float(d.device.is_on)])
    out['switch'] = np.asarray(rows)
    np.savez_compressed(os.path.join(GOLD, 'devices.npz'), **out)
    print('devices', {k: v.shape for k, v in out.items()})


def main():
    os.makedirs(GOLD, exist_ok=True)
    what = sys.argv[1:] or ['c2', 'c1', 'c3', 'c5', 'c4', 'dev']
    for w in what:
        if w == 'dev':
            gen_devices()
        elif w == 'misc':
            import make_golden_misc
            make_golden_misc.main()
        else:
            gen_workload(w)


if __name__ == '__main__':
    main()
