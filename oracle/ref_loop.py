"""TEST / BENCH INFRASTRUCTURE — the reference's own CPU path as a timed loop.

Runs the UNMODIFIED reference package (pure Python, ahkab v0.18) on instances of the five
BASELINE workloads, one `ahkab.Circuit` + `ahkab.run` per instance — the loop BASELINE.md §3
names as "the reference" — in this process or in a multiprocessing pool over the host cores.
The package is imported from $AHKAB_REFERENCE, else `baseline/_ref` (the offline
`pip install --target` of /root/reference recorded in DESIGN.md §7; git-ignored, travels to the
GPU box), else /root/reference, through oracle/refshim.py (in-process compatibility shims for
modern numpy / scipy; no reference source is edited or copied).

Only bench.py's `--impl reference` / `cpu_baseline` legs and tests use this module."""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def reference_path():
    for p in (os.environ.get('AHKAB_REFERENCE'), os.path.join(ROOT, 'baseline', '_ref'),
              '/root/reference'):
        if p and os.path.isdir(os.path.join(p, 'ahkab')):
            return p
    return None


def _load():
    import warnings
    warnings.simplefilter('ignore')
    p = reference_path()
    if p is None:
        raise RuntimeError("the reference package is not available (baseline/_ref missing)")
    os.environ['AHKAB_REFERENCE'] = p
    for d in (HERE, ROOT):
        if d not in sys.path:
            sys.path.insert(0, d)
    import refshim
    refshim.REF = p
    # (the reference prints locale / matplotlib warnings on import: keep stdout clean)
    devnull = open(os.devnull, 'w')
    so, se = sys.stdout, sys.stderr
    sys.stdout = sys.stderr = devnull
    try:
        from ahkab_b200 import workloads  # noqa: F401  (pulls in torch: seconds, outside any clock)
        return refshim.load()
    finally:
        sys.stdout, sys.stderr = so, se
        devnull.close()


def run_instance(args):
    """One instance of workload `name` through the reference's public API; returns its unit
    count (instance-solves or accepted time steps)."""
    name, i, B = args
    ahkab = _load()
    from ahkab_b200 import workloads as W
    w = W.ALL[name](B=B)
    saved = {k: getattr(ahkab.options, k) for k in w.opts}
    for k, v in w.opts.items():
        setattr(ahkab.options, k, v)
    try:
        circ = w.build(ahkab.Circuit, w.instance_params(i))
        if name == 'c1':
            r = ahkab.run(circ, ahkab.new_ac(x0=None, verbose=0, **W.C1_AC))['ac']
            return r.asarray().shape[1]
        if name == 'c2':
            ahkab.run(circ, ahkab.new_op(verbose=0))['op']
            return 1
        if name == 'c3':
            an = [ahkab.new_op(verbose=0), ahkab.new_tran(x0='op+ic', verbose=0, **W.C3_TRAN)]
            return ahkab.run(circ, an)['tran'].asarray().shape[1]
        if name == 'c4':
            x0 = ahkab.new_x0(circ, W.C4_IC)
            r = ahkab.run(circ, ahkab.new_tran(x0=x0, verbose=0, **W.C4_TRAN))['tran']
            return r.asarray().shape[1]
        if name == 'c5':
            r = ahkab.run(circ, ahkab.new_dc(verbose=0, **W.C5_DC))['dc']
            return r.asarray().shape[1]
    finally:
        for k, v in saved.items():
            setattr(ahkab.options, k, v)
    raise ValueError(name)


def timed_loop(name, count, B, procs=1):
    """`count` instances (the first ones of the seeded batch of B) -> (units, seconds).
    procs > 1: a multiprocessing pool (the workers import the reference before the clock
    starts)."""
    jobs = [(name, i, B) for i in range(count)]
    if procs <= 1:
        _load()
        run_instance((name, 0, B) if name in ('c1', 'c2') else ('c2', 0, 64))   # first-call costs
        t0 = time.perf_counter()
        units = sum(run_instance(j) for j in jobs)
        return units, time.perf_counter() - t0
    import multiprocessing as mp
    ctx = mp.get_context('fork')
    with ctx.Pool(procs, initializer=_load) as pool:
        pool.map(run_instance, [('c2', 0, 64)] * (4 * procs), chunksize=1)   # first-call costs in every worker
        t0 = time.perf_counter()
        units = sum(pool.map(run_instance, jobs, chunksize=max(1, count // (procs * 8))))
        dt = time.perf_counter() - t0
    return units, dt


if __name__ == '__main__':
    nm = sys.argv[1] if len(sys.argv) > 1 else 'c2'
    cnt = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    pr = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    u, s = timed_loop(nm, cnt, max(cnt, 64), pr)
    print(nm, 'instances', cnt, 'procs', pr, 'units', u, 'seconds %.3f' % s, 'units/s %.1f' % (u / s))
