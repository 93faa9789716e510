#!/usr/bin/env python
"""bench.py — throughput of the batched Newton/MNA hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c2] [--only] [--no-cpu]

One JSON line on stdout (rank 0).  A *step* is one pass of an analysis over one
synthetic batch:

  headline  c2  tests/diode_op operating point, 65536 instances  (BASELINE configs[1])
  also      c1  Butterworth AC 4096 x 100 freq        -> instance-solves/s
            c3  Colpitts OP + TRAP fixed step 16384   -> instance-timesteps/s
            c4  ring3 GEAR2 + LTE 16384               -> instance-timesteps/s
            c5  R-2R n=257 DC sweep 4096 x 10         -> instance-solves/s
  (reported under "configs" in the same line; --only skips them)

`value`   kernel path with the per-instance parameters already resident in HBM,
          CUDA events on the launching stream, L2 flushed between timed steps.
`e2e`     the same step through the Engine API with HOST buffers: pinned
          H2D of the parameters, the analysis, pinned D2H of every result.
N > 1     one process per GPU (torchrun), STRONG scaling: the fixed BASELINE batch is
          cut into contiguous shards (ahkab_b200/dist.py: shard_range), one per rank;
          `value` = units of the whole batch / max-over-ranks kernel time.  The e2e
          step of every rank is pinned H2D of its shard, the analysis, then the ONE
          result collective — every rank's packed result buffer (x / resid / iters /
          status, packed transient rows) to rank 0 over NCCL (grouped send/recv) —
          and rank 0's pinned D2H of the gathered batch.  `e2e.hbm_resident` stops
          the clock when the results sit in rank 0's HBM; `e2e.breakdown_ms` splits a
          step into H2D / kernels / gather / D2H with CUDA events; `weak` is the old
          weak-scaling figure (every rank its own full-size batch), kept as a
          secondary key.
--impl reference   the reference's algorithm on the host cores: the C
          restatement in oracle/ (the reference itself is pure Python and does
          not travel to the GPU box), OpenMP over all cores, rank 0 only.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, 'tests')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

UNITS = dict(c1='instance-solves/s', c2='instance-solves/s',
             c3='instance-timesteps/s', c4='instance-timesteps/s',
             c5='instance-solves/s')
NAMES = dict(
    c1='README Butterworth band-pass AC sweep, 100 log points, B=4096 (n=14 complex128)',
    c2='tests/diode_op operating point, B=65536 (n=3)',
    c3='tests/colpitts OP + TRAP fixed-step transient 250 steps, B=16384 (n=9, 1 EKV)',
    c4='tests/ring3 GEAR2 + LTE step control transient to 50 ns, B=16384 (n=5, 6 EKV)',
    c5='tests/r2r 256-node ladder DC sweep 10 points, B=4096 (n=257)')
FULL_B = dict(c1=4096, c2=65536, c3=16384, c4=16384, c5=4096)
# bounded samples for the CPU arm (instances), sized for a few seconds per step
CPU_SAMPLE = dict(c1=4096, c2=65536, c3=2048, c4=256, c5=64)
C4_MAX_ROWS = 3072
# chunks of the pipelined end-to-end step (copy-in / compute / copy-out overlap)
# Only throughput-bound kernels gain: a latency-bound kernel that fits the machine in one
# or two waves takes as long for a quarter of the batch as for all of it (measured: C3
# 28 -> 44 ms, C4 262 -> 679 ms with 4 / 8 chunks), so those run as one chunk.
E2E_CHUNKS = dict(c1=4, c2=1, c3=1, c4=1, c5=1)   # c5: thread-per-instance kernels are latency bound: a quarter of the batch takes as long as all of it
# The c2 end-to-end step goes through ahk_op_host (host buffers in, host buffers out) in its
# ZERO-COPY form (chunks = 0): the kernel reads the pinned parameters and writes the pinned
# results itself, so both PCIe directions overlap the computation inside the one launch — no
# separate H2D / D2H copies, the same bytes cross the bus.  AHK_E2E_OP_HOST=<chunks> selects the
# chunked copy pipeline inside the library, AHK_E2E_OP_HOST=engine the packed copies through the
# Engine (set_vary + op + to_host).  Measured on B200 (50 steps): zero-copy 0.103 ms; Engine
# 0.127 ms; op_host with 1 / 2 chunks 0.138 / 0.178 ms (every extra DMA costs ~8 us: the kernels
# are too short to hide them).  Zero-copy took 0.414 ms while every lane wrote its own [n] row
# (partial PCIe sectors); the one-lane specialised kernel now writes the block's rows
# cooperatively, 256 contiguous bytes per warp store (entry.cuh: run_op_rows).
OP_HOST = os.environ.get('AHK_E2E_OP_HOST', '0')
if OP_HOST == 'engine':
    OP_HOST = None


def config_of(name):
    """The workload description — identical in both arms (ours / --impl reference)."""
    return {"workload": NAMES[name], "instances": FULL_B[name],
            "l2": "GPU arm: flushed (512 MiB write) before every timed step; CPU arm: not applicable"}


def LU(n):
    return 2.0 / 3.0 * n ** 3 + 2.0 * n ** 2


# algorithmic flop per NR iteration: dense LU+solve, residual
# mat-vec, device evaluations (diode ~120, EKV ~400 flop incl. transcendentals)
NR_FLOP = dict(c2=LU(3) + 2 * 9 + 120, c3=LU(9) + 2 * 81 + 400,
               c4=LU(5) + 2 * 25 + 6 * 400)
AC_FLOP = 4 * LU(14) + 8 * 14 ** 2            # per (instance, frequency)
# c5: two dense factorisations per instance amortised over the 10-point sweep +
# per point two (residual mat-vec + forward/backward substitution)
C5_FLOP = (2 * LU(257) - 2 * 2 * 257 ** 2) / 10.0 + 2 * (2 * 257 ** 2 + 2 * 257 ** 2)


HEADLINE = 'c2'     # --workload
JIT_MODE = 'auto'   # --jit: circuit-specialised kernels (ahk_set_jit); auto = batches >= 1024 here


# ---------------------------------------------------------------------------------
class Runner(object):
    """One workload on one GPU: `step()` with inputs resident, `e2e()` from host."""

    def __init__(self, name, seed_off=0, B=None, rank=0, world=1):
        """world > 1: this rank's contiguous shard of the one seeded batch of B instances
        (strong scaling); seed_off: a different full batch per rank (weak scaling)."""
        import torch
        import common
        from ahkab_b200 import workloads as W
        from ahkab_b200 import dist as ahk_dist
        from ahkab_b200.circuit import Circuit
        from ahkab_b200.engine import Engine
        self.torch, self.common, self.W, self.ahk_dist = torch, common, W, ahk_dist
        self.name = name
        B = B or FULL_B[name]
        seeds = dict(c1=0, c2=1, c3=2, c4=3, c5=4)
        self.w = W.ALL[name](B=B, seed=seeds[name] + 1000 * seed_off)
        self.circ = self.w.circuit(Circuit)
        self.rank, self.world, self.B_total = rank, world, B
        lo, hi = ahk_dist.shard_range(B, world, rank)
        self.lo, self.hi = lo, hi
        B = hi - lo
        self.eng = Engine(self.circ, self.w.batch(lo, hi), opts=self.w.opts)
        # (threshold 1024: the chunks of the pipelined end-to-end paths are sub-batches of
        # 1024 instances and the bench repeats every call, so the compilation amortises)
        # (a rank's shard of a strong-scaling run may be smaller than that — C1 / C5 at 8 GPUs hold
        # 512 instances each: the bench repeats every call, so the compilation amortises there too)
        self.eng.set_jit(JIT_MODE, min_batch=min(1024, max(1, B)))
        self.B = B
        self.pipe = None
        self.vary_host = torch.empty(self.eng.vary.shape, dtype=torch.float64,
                                     pin_memory=True)
        self.vary_host.copy_(self.eng.vary)
        torch.cuda.synchronize()
        self.h2d = self.vary_host.numel() * 8
        if name == 'c1':
            self.om = common.c1_omegas()
        if name == 'c5':
            self.sw = common.c5_sweep()
        self.gbuf = None      # rank 0: destination of the result gather (grow-only)
        if name == 'c4':
            self.x0_host = torch.as_tensor(common.c4_x0(self.circ, B)).pin_memory()
            self.x0 = self.x0_host.cuda()
            self.h2d += self.x0_host.numel() * 8

    def step(self, e=None, x0=None):
        e = e or self.eng
        W, n = self.W, self.name
        if n == 'c2':
            return e.op()
        if n == 'c1':
            return e.ac(self.om)
        if n == 'c3':
            op = e.op(want_resid=False)
            x0 = op['x'].clone()             # icmodified_x0 (dc_analysis.py:1143-1190)
            x0[:, 1] = x0[:, 2] + 2.5        # c1 (nd, ns) ic = 2.5
            x0[:, 6] = -1e-9                 # l1 ic = -1n
            r = e.tran(x0, max_rows=256, **W.C3_TRAN)
            r['op_iters'] = op['iters']
            return r
        if n == 'c4':
            return e.tran(self.x0 if x0 is None else x0, max_rows=C4_MAX_ROWS, **W.C4_TRAN)
        if n == 'c5':
            return e.dc_sweep('V1', self.sw)

    def e2e(self):
        """host params -> device, analysis, every result tensor -> pinned host, cut into
        chunks and pipelined over three streams (ahkab_b200/pipeline.py): the H2D of the
        next chunk and the D2H of the previous one overlap the kernels."""
        if self.name == 'c2' and OP_HOST is not None:
            host = self.eng.op_host(self.vary_host, nchunks=int(OP_HOST))
            return host, host
        if E2E_CHUNKS[self.name] == 1:
            self.eng.set_vary(self.vary_host)
            if self.name == 'c4':
                self.x0.copy_(self.x0_host, non_blocking=True)
            raw = self.step()
            if self.name in ('c3', 'c4'):
                # transients: only the rows that exist go to the host (ahk_pack_rows packs the
                # ragged [B][max_rows] buffers; C4 uses ~1790 of its 3072 rows per instance)
                return raw, self.eng.to_host(self.eng.pack_tran(raw))
            return raw, self.eng.to_host(raw)
        if self.pipe is None:
            from ahkab_b200.pipeline import Pipeline
            self.pipe = Pipeline(self.eng, nchunks=E2E_CHUNKS[self.name])
        extra = {'x0': self.x0_host} if self.name == 'c4' else None
        host = self.pipe.run(lambda sub, lo, hi, x: self.step(sub, x.get('x0')),
                             self.vary_host, extra)
        return host, host

    def api(self, fresh=False):
        """The drop-in call a user of the reference makes: `ahkab.run(circ, analysis,
        batch=ParamBatch)` from host numpy parameters to host numpy results (every result
        array of the analysis), through ahkab_b200.analyses.run and the results objects."""
        import ahkab_b200 as ahkab
        W, n = self.W, self.name
        saved = {k: getattr(ahkab.options, k) for k in self.w.opts}
        for k, v in self.w.opts.items():      # the workload's option overrides, as module globals
            setattr(ahkab.options, k, v)      # (the reference's way: ahkab.options.vea = ...)
        try:
            return self._api(ahkab, W, n, fresh)
        finally:
            for k, v in saved.items():
                setattr(ahkab.options, k, v)

    def _api(self, ahkab, W, n, fresh):
        if fresh or not hasattr(self, 'api_batch'):
            self.api_batch = self.w.batch(self.lo, self.hi)     # ParamBatch of host numpy arrays
        if n == 'c2':
            r = ahkab.run(self.circ, ahkab.new_op(), batch=self.api_batch)['op']
            return r.x, r.iterations, r.status
        if n == 'c1':
            r = ahkab.run(self.circ, ahkab.new_ac(x0=None, **W.C1_AC), batch=self.api_batch)['ac']
            return r._get('x'), r.status
        if n == 'c5':
            r = ahkab.run(self.circ, ahkab.new_dc(**W.C5_DC), batch=self.api_batch)['dc']
            return r._get('x'), r.status
        if n == 'c3':     # OP, then the transient from the ic-modified operating point (tests/colpitts)
            r = ahkab.run(self.circ, [ahkab.new_op(), ahkab.new_tran(x0='op+ic', **W.C3_TRAN)],
                          batch=self.api_batch)['tran']
            t, x, _offs = r.packed()
            return t, x, r.rows, r.status
        if n == 'c4':
            x0 = ahkab.new_x0(self.circ, W.C4_IC)
            r = ahkab.run(self.circ, ahkab.new_tran(x0=x0, **W.C4_TRAN), batch=self.api_batch)['tran']
            t, x, _offs = r.packed()
            return t, x, r.rows, r.status
        return None

    def e2e_dist(self, d2h=True, ev=None):
        """One end-to-end step of a sharded run: pinned H2D of this rank's shard, the
        analysis, the single result collective (every rank's packed result buffer to rank
        0, NCCL), and rank 0's pinned D2H of the gathered batch.  ev: optional list that
        receives CUDA events at the phase boundaries (H2D | kernels | gather | D2H)."""
        torch = self.torch
        def mark():
            if ev is not None:
                e = torch.cuda.Event(enable_timing=True); e.record(); ev.append(e)
        mark()
        self.eng.set_vary(self.vary_host)
        if self.name == 'c4':
            self.x0.copy_(self.x0_host, non_blocking=True)
        mark()
        raw = self.step()
        tran = self.name in ('c3', 'c4')
        pk = self.eng.pack_tran(raw) if tran else raw
        mark()
        big, views = self.ahk_dist.gather_result(pk, ragged=tran, out=self.gbuf)
        mark()
        host = None
        if self.rank == 0:
            self.gbuf = big if (self.gbuf is None or big.numel() > self.gbuf.numel()) else self.gbuf
            self.gathered_bytes = int(big.numel())
            if d2h:
                host = self.eng.pinned_bytes(big.numel())
                host.copy_(big, non_blocking=True)
        mark()
        return raw, host

    def d2h_bytes(self, host):
        return int(sum(v.numel() * v.element_size() for v in host.values()))

    def units(self, raw):
        n = self.name
        if n == 'c2':
            return self.B
        if n == 'c1':
            return self.B * len(self.om)
        if n == 'c5':
            return self.B * len(self.sw)
        return int(raw['rows'].sum().item())   # device tensor or pinned host tensor

    def algo(self, raw):
        """(flop, bytes) one step needs algorithmically."""
        n, B, N = self.name, self.B, self.eng.n
        nv = self.vary_host.shape[0]
        if n == 'c2':
            it = float(raw['iters'].sum().item())
            return it * NR_FLOP['c2'], B * (8 * nv + 8 * N * 2 + 8)
        if n == 'c1':
            u = B * len(self.om)
            return u * AC_FLOP, B * 8 * nv + u * 16 * N + 4 * B
        if n == 'c5':
            u = B * len(self.sw)
            return u * C5_FLOP, B * 8 * nv + u * (8 * N + 8)
        rows = float(raw['rows'].sum().item())
        nr = float(raw['counters'][:, 1].sum().item())
        if n == 'c3':
            nr += float(raw['op_iters'].sum().item())
        return nr * NR_FLOP[n], B * (8 * nv + 8 * N + 28) + rows * 8 * (N + 1)


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while `active` is set."""

    def __init__(self, torch_index):
        threading.Thread.__init__(self, daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.active = threading.Event()
        self.stop = threading.Event()
        self.h = None
        try:
            import pynvml
            import torch
            self.nv = pynvml
            pynvml.nvmlInit()
            try:
                uuid = 'GPU-' + str(torch.cuda.get_device_properties(torch_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                vis = os.environ.get('CUDA_VISIBLE_DEVICES')
                idx = int(vis.split(',')[torch_index]) if vis else torch_index
                self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as ex:          # noqa
            self.err = repr(ex)
            self.h = None

    def run(self):
        if self.h is None:
            return
        nv = self.nv
        names = {0x4: 'sw_power_cap', 0x8: 'hw_slowdown', 0x20: 'sw_thermal_slowdown',
                 0x40: 'hw_thermal_slowdown', 0x80: 'hw_power_brake_slowdown'}
        while not self.stop.is_set():
            if self.active.is_set():
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    try:
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, nm in names.items():
                        if r & bit:
                            self.reasons.add(nm)
                except Exception:
                    pass
                time.sleep(0.002)
            else:
                time.sleep(0.001)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [],
                    "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


class Timer(object):
    def __init__(self, dist_on, device):
        import torch
        self.torch, self.dist_on = torch, dist_on
        self.flush = torch.empty(512 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist_on:
            import torch.distributed as dist
            dist.barrier()
            self.torch.cuda.synchronize()

    def timed(self, fn, steps, warmup, sampler=None):
        """W untimed + K timed steps; L2 flushed before every timed step; returns
        (sum of the K device durations in ms, max over ranks; last result)."""
        torch = self.torch
        r = None
        for _ in range(warmup):
            r = fn()
        self.barrier()
        if sampler is not None:
            sampler.active.set()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(steps)]
        # one untimed step ahead of the K timed ones, in the same stream without a
        # synchronisation in between: the first step after a barrier otherwise reads 0.10-0.20
        # ms for a 0.018 ms kernel (per-step trace, AHK_BENCH_DEBUG=1) because the device waits
        # for the host to enqueue it; from the second step on the host runs ahead of the stream
        # (every step starts with an 85 us L2 flush) and the event pairs measure device time
        self.flush.zero_()
        r = fn()
        for a, b in ev:
            self.flush.zero_()
            a.record()
            r = fn()
            b.record()
        self.barrier()
        if sampler is not None:
            sampler.active.clear()
        ms = sum(a.elapsed_time(b) for a, b in ev)
        if os.environ.get('AHK_BENCH_DEBUG'):
            sys.stderr.write('per-step ms: %s\n' % ' '.join('%.4f' % a.elapsed_time(b) for a, b in ev))
        if self.dist_on:
            import torch.distributed as dist
            t = torch.tensor([ms], dtype=torch.float64, device=self.flush.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, r


def all_sum(v, dist_on, device):
    if not dist_on:
        return v
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(v)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def measure_fp64_peak(eng, timer):
    """Measured FP64-FMA throughput of this GPU (TFLOP/s): best of 5 launches of
    the library's probe kernel, CUDA events."""
    import torch
    best = 0.0
    eng.probe_fp64(iters=1024)
    torch.cuda.synchronize()
    for _ in range(5):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        fl = eng.probe_fp64(iters=16384)
        b.record()
        torch.cuda.synchronize()
        best = max(best, fl / (a.elapsed_time(b) * 1e-3) / 1e12)
    return best


def measured_traffic(name):
    """DRAM bytes per launch of the workload's dominant kernel from the committed ncu
    capture (profiles/traffic.json), or None."""
    try:
        t = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))[name]
        return {"bytes": t["bytes"], "launch": t["launch"], "kernel": t["kernel"]}
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ---- CPU arm: the oracle (C restatement of the reference), OpenMP ---------------------
def host_threads():
    """Host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which
    would silently turn the all-cores CPU arm into a 1-core run)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_step_fn(name, B):
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import oracle
    import common
    from ahkab_b200 import workloads as W
    from ahkab_b200.circuit import Circuit
    seeds = dict(c1=0, c2=1, c3=2, c4=3, c5=4)
    w = W.ALL[name](B=B, seed=seeds[name])
    circ = w.circuit(Circuit)
    o = oracle.Oracle(circ, w.batch(), opts=w.opts)
    NT = host_threads()
    if name == 'c2':
        return (lambda: B), (lambda: o.op(threads=NT)), NT
    if name == 'c1':
        om = common.c1_omegas()
        return (lambda: B * len(om)), (lambda: o.ac(om, threads=NT)), NT
    if name == 'c5':
        sw = common.c5_sweep()
        return (lambda: B * len(sw)), (lambda: o.dc_sweep(0, sw, threads=NT)), NT
    box = {}
    if name == 'c3':
        def f():
            x0 = common.c3_x0(o.op(threads=NT)['x'])
            box['r'] = o.tran(x0, 0., W.C3_TRAN['tstep'], W.C3_TRAN['tstop'], 'TRAP',
                              False, 256, threads=NT)
    else:
        x0 = common.c4_x0(circ, B)

        def f():
            box['r'] = o.tran(x0, 0., W.C4_TRAN['tstep'], W.C4_TRAN['tstop'], 'GEAR2',
                              True, C4_MAX_ROWS, threads=NT)
    return (lambda: int(box['r']['rows'].sum())), f, NT


def cpu_measure(name, steps, warmup):
    B = CPU_SAMPLE[name]
    units, f, cores = cpu_step_fn(name, B)
    for _ in range(warmup):
        f()
    t0 = time.perf_counter()
    for _ in range(steps):
        f()
    dt = time.perf_counter() - t0
    return dict(value=units() * steps / dt, unit=UNITS[name], cores=cores, kind='port',
                sample='%d of %d instances of the same seeded batch, %d step(s), OpenMP '
                       'over %d threads (oracle/ahk_oracle.c)' % (B, FULL_B[name], steps, cores)), \
        dt / steps * 1e3


# instances per host core of the Python-reference sample (one ahkab.run per instance:
# ~3 ms c2, 20 ms c1, 0.35 s c3 / c5, 16 s c4 per instance on one core)
PY_SAMPLE_PER_CORE = dict(c2=256, c1=64, c3=4, c4=1, c5=4)


def python_reference(name, quick=False):
    """The reference itself — ahkab v0.18's pure-Python `ahkab.run`, one Circuit per instance,
    BASELINE.md §3 — on a bounded sample of the same seeded batch: one process, and a pool
    over all host cores.  None when the package is not on this machine (baseline/_ref is the
    offline pip install recorded in DESIGN.md §7; it travels to the GPU box)."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    try:
        import ref_loop
        if ref_loop.reference_path() is None:
            return None
        cores = host_threads()
        per = PY_SAMPLE_PER_CORE[name]
        if quick:
            per = max(1, per // 4)
        n1 = max(1, per // 2) if name != 'c4' else 0          # (c4: 16 s per instance, pool only)
        out = {"kind": "reference", "unit": UNITS[name],
               "what": "ahkab.run loop (pure Python, unmodified reference via oracle/refshim.py), "
                       "one Circuit per instance", "path": ref_loop.reference_path()}
        if n1:
            u, dt = ref_loop.timed_loop(name, n1, FULL_B[name], 1)
            out["one_core"] = {"value": u / dt, "instances": n1, "seconds": dt}
        cnt = per * cores
        u, dt = ref_loop.timed_loop(name, cnt, FULL_B[name], cores)
        out.update({"value": u / dt, "cores": cores, "instances": cnt, "seconds": dt,
                    "sample": "%d of %d instances of the same seeded batch, multiprocessing pool "
                              "over %d processes" % (cnt, FULL_B[name], cores)})
        return out
    except Exception as ex:      # the Python reference is a secondary figure: never fail the arm
        return {"kind": "reference", "unavailable": repr(ex)[:200]}


def run_reference(args, rank):
    if rank != 0:
        return
    names = [args.workload] + ([] if args.only else
                               [c for c in ('c1', 'c3', 'c4', 'c5', 'c2') if c != args.workload])
    out = None
    cfgs = {}
    for i, name in enumerate(names):
        steps = args.steps if i == 0 else min(args.steps, 3)
        warm = args.warmup if i == 0 else 1
        if name in ('c3', 'c4', 'c5'):
            steps, warm = min(steps, 5), min(warm, 1)
        cb, ms = cpu_measure(name, steps, warm)
        line = {"metric": UNITS[name].replace('/s', '/sec'), "value": cb['value'],
                "unit": UNITS[name], "ms_per_step": ms, "steps": steps, "warmup": warm,
                "cpu_baseline": cb,
                "e2e": {"value": cb['value'], "unit": UNITS[name], "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "config": config_of(name)}
        # beside the C port: the real Python reference on a (smaller) sample
        py = python_reference(name, quick=(i != 0))
        if py is not None:
            line["python_reference"] = py
        if i == 0:
            out = line
        else:
            cfgs[name] = line
    out.update({"impl": "reference", "n_gpus": args.gpus, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "gpu_launches": 0})
    if cfgs:
        out["configs"] = cfgs
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------
def bench_one(name, rank, world, timer, sampler, steps, warmup, hbm_peak, hbm_src, fp64_peak,
              dist_on, device, want_cpu):
    import torch
    r = Runner(name, rank=rank, world=world)
    eng = r.eng
    # kernel path, inputs resident
    # kernels of this library per step (counted on one untimed step), times the K timed steps
    r.step()
    l0, j0 = eng.launch_count(), eng.jit_launch_count()
    r.step()
    launches = (eng.launch_count() - l0) * steps
    jit_launches = (eng.jit_launch_count() - j0) * steps
    shape = eng.last_launch() if name != 'c5' else None
    ms, raw = timer.timed(r.step, steps, warmup, sampler)
    units = all_sum(r.units(raw), dist_on, device)
    flop, byts = r.algo(raw)
    value = units * steps / (ms * 1e-3)
    # end to end, host buffers
    hbm_resident = breakdown = None
    if dist_on:
        # (NCCL sets its point-to-point channels up lazily, pair by pair, over the first exchanges:
        # a fixed number of untimed gathers first — the first workload of an 8-rank run otherwise
        # reads 0.45 ms per step for the HBM-resident figure and 0.18 ms for the longer path)
        for _ in range(8):
            r.e2e_dist(d2h=False)
        ms_e, (raw_e, host) = timer.timed(r.e2e_dist, steps, warmup)
        ms_a, _ = timer.timed(lambda: r.e2e_dist(d2h=False), steps, warmup)
        hbm_resident = {"value": units * steps / (ms_a * 1e-3), "ms_per_step": ms_a / steps,
                        "what": "same step, clock stopped when the gathered results are in rank 0's HBM"}
        # phase breakdown of one step on rank 0 (CUDA events on the stream between the phases)
        evs = []
        timer.barrier()
        r.e2e_dist(ev=evs)
        timer.barrier()
        ph = [evs[i].elapsed_time(evs[i + 1]) for i in range(4)]
        breakdown = {"h2d": ph[0], "kernels_and_pack": ph[1], "nccl_gather": ph[2], "d2h_rank0": ph[3],
                     "what": "rank 0, one step, CUDA events between the phases (the gather includes "
                             "waiting for the slowest rank)"}
        d2h_bytes = int(all_sum(getattr(r, 'gathered_bytes', 0), dist_on, device))
        h2d_bytes = int(all_sum(r.h2d, dist_on, device))
    else:
        ms_e, (raw_e, host) = timer.timed(r.e2e, steps, warmup)
        d2h_bytes, h2d_bytes = r.d2h_bytes(host), r.h2d
    units_e = all_sum(r.units(raw_e), dist_on, device)
    e2e_value = units_e * steps / (ms_e * 1e-3)
    e2e_api = None
    if world == 1:
        # the same step through the drop-in API (host numpy in, host numpy out), host wall clock
        # (three untimed calls: the caller holds the previous result while the next call
        # runs, so the engine's pool of pinned result buffers settles at two — pinning the
        # 1.4 GB of a C4 result set costs ~0.3 s once)
        for _ in range(3):
            out_api = r.api()
        torch.cuda.synchronize()
        asteps = steps if name not in ('c3', 'c4') else min(steps, 3)
        t0 = time.perf_counter()
        for _ in range(asteps):
            out_api = r.api()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / asteps
        t0 = time.perf_counter()
        for _ in range(asteps):
            out_api = r.api(fresh=True)
        torch.cuda.synchronize()
        dt_fresh = (time.perf_counter() - t0) / asteps
        e2e_api = {"value": units_e / dt, "unit": UNITS[name], "ms_per_step": dt * 1e3,
                   "vs_e2e": (dt * 1e3) / (ms_e / steps),
                   "what": "ahkab_b200.run(circ, analysis, batch=ParamBatch) -> every result array as "
                           "host numpy, host wall clock; the ParamBatch is built once from host numpy "
                           "arrays, every call copies its (pinned) value matrix host->device",
                   "fresh_batch_ms_per_step": dt_fresh * 1e3,
                   "fresh_batch": "the same with a NEW ParamBatch built from the numpy arrays inside "
                                  "every step (re-flattening, new pinned staging)",
                   "result_bytes": int(sum(a.nbytes for a in out_api))}
        del out_api
    ok = int((raw['status'] & 1).sum().item()) if 'status' in raw else None
    step_s = ms * 1e-3 / steps
    par = ("single GPU" if world == 1 else
           "fixed batch of %d instances in %d contiguous shards (one rank per GPU); one NCCL "
           "result collective per step: every rank's packed results to rank 0 (grouped send/recv)"
           % (r.B_total, world))
    line = {
        "metric": UNITS[name].replace('/s', '/sec'), "value": value,
        "unit": UNITS[name], "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps, "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None, "dtype": "f64" if name != 'c1' else "complex128",
        "data": "synthetic",
        "config": config_of(name),
        "details": {"instances_per_gpu": r.B, "units_per_step": units, "n_unknowns": eng.n,
                    "parallelism": par, "status_ok_rank0": ok, "kernel_shape_rank0": shape,
                    "kernels": ("%d of %d launches per run are circuit-specialised kernels "
                                "(NVRTC, sm_100a; --jit %s)" % (jit_launches, launches, JIT_MODE))},
        "e2e": {"value": e2e_value, "unit": UNITS[name], "ms_per_step": ms_e / steps,
                "chunks": E2E_CHUNKS[name] if world == 1 else 1,
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
        "gpu_launches": int(launches),
        "roofline": {
            "bound": "fp64", "achieved": flop / step_s / 1e12, "peak": fp64_peak,
            "unit": "TFLOP/s", "frac": flop / step_s / 1e12 / fp64_peak if fp64_peak else None,
            "peak_source": "measured in this run: DFMA probe kernel (ahk_probe_fp64), best of 5",
            "algorithmic_flop_per_step": flop,
            "flop_basis": ("SURVEY 8(d): measured iterations / steps / points x [dense LU(n) + residual "
                           "mat-vec + device evaluations] of the reference's algorithm"
                           + ("; the static-schedule kernels of this workload (core.cuh: slu_solve_thread, "
                              "ac.cuh: zslu_solve_thread) execute the LU on the structural non-zeros only, "
                              "i.e. fewer flop than counted here — executed FP64 instruction counts are in "
                              "the ncu summaries under profiles/" if name in ('c1', 'c3', 'c4') else "")),
            "scope": "rank 0's kernels on its shard" if world > 1 else "the whole batch",
            "traffic": (measured_traffic(name) or {}).get("bytes"),
            "traffic_source": measured_traffic(name),
            "hbm": {"bound": "hbm", "achieved": byts / step_s / 1e9, "peak": hbm_peak,
                    "unit": "GB/s", "frac": byts / step_s / 1e9 / hbm_peak,
                    "peak_source": hbm_src, "algorithmic_bytes_per_step": byts}},
    }
    if name == 'c2' and world == 1:
        line["e2e"]["path"] = ("Engine: set_vary (pinned H2D) + ahk_op + to_host (pinned D2H)" if OP_HOST is None else
                               "ahk_op_host, zero-copy: the kernel reads the pinned host parameters and writes the "
                               "pinned host results itself (the same bytes cross PCIe inside the one launch)"
                               if int(OP_HOST) == 0 else
                               "ahk_op_host, %d chunk(s) through the library's copy-in / kernel / copy-out streams" % int(OP_HOST))
    if e2e_api:
        line["e2e_api"] = e2e_api
    if hbm_resident:
        line["e2e"]["hbm_resident"] = hbm_resident
        line["e2e"]["breakdown_ms"] = breakdown
    if name == 'c5':
        line["roofline"].update(c5_roofline_extra(eng, r, step_s, fp64_peak))
    if want_cpu and rank == 0 and world == 1:
        cb, _ms = cpu_measure(name, 2 if name in ('c3', 'c4') else 3, 1)
        line["cpu_baseline"] = cb
        if name == HEADLINE:
            py = python_reference(name)
            if py is not None:
                line["python_reference"] = py
    del r
    torch.cuda.empty_cache()
    if dist_on:
        # secondary: the weak-scaling figure (every rank its own full-size batch, kernel path)
        rw = Runner(name, seed_off=rank)
        ms_w, raw_w = timer.timed(rw.step, min(steps, 3), 3)
        units_w = all_sum(rw.units(raw_w), dist_on, device)
        line["weak"] = {"value": units_w * min(steps, 3) / (ms_w * 1e-3), "ms_per_step": ms_w / min(steps, 3),
                        "instances_per_gpu": rw.B, "scaling": "weak",
                        "what": "kernel path, every rank its own full-size batch, no collective"}
        del rw
        torch.cuda.empty_cache()
    return line


def c5_roofline_extra(eng, r, step_s, fp64_peak):
    """C5 runs a SPARSE static elimination schedule, not a dense LU: `achieved` / `frac` are
    re-stated on the flop that schedule executes (counted on the host from the schedule,
    ahk_big_stats); the dense-equivalent figure (2 LU(257) per instance) is kept beside it as
    what a dense solver would have had to do.  One thread walks a serial dependency chain per
    instance / (instance, point): the bound is latency, not the FP64 pipe."""
    st = eng.big_stats()
    if not st.get('n'):
        return {}
    B, npts = r.B, len(r.sw)
    flop = float(B) * st['factor_flop'] + float(B) * npts * st['point_flop']
    dense = B * npts * C5_FLOP
    # HBM traffic the compiled path needs: parameters in, factor values + matrix values written
    # once and read once per sweep point (L2 permitting), results out
    nv = r.vary_host.shape[0]
    byts = B * 8.0 * nv + B * 8.0 * (2 * st['nf'] + st['npos']) * 2 + B * npts * (8.0 * st['n'] + 8)
    return {"bound": "latency", "achieved": flop / step_s / 1e12,
            "frac": flop / step_s / 1e12 / fp64_peak if fp64_peak else None,
            "algorithmic_flop_per_step": flop,
            "what": "flop executed by the static sparse elimination schedule (host count); "
                    "serial per-thread chains: latency bound, the FP64 pipe fraction is reported for scale only",
            "path": st['path'],
            "dense_equivalent": {"flop_per_step": dense, "achieved": dense / step_s / 1e12,
                                 "frac": dense / step_s / 1e12 / fp64_peak if fp64_peak else None,
                                 "what": "two dense LU(257) per instance + dense substitutions, the work of the reference's dense solver"},
            "schedule_bytes_per_step": byts,
            "schedule_hbm_gbs": byts / step_s / 1e9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c2', choices=sorted(UNITS))
    ap.add_argument('--only', action='store_true', help='headline workload only')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--jit', default='auto', choices=['auto', 'on', 'off'],
                    help='circuit-specialised kernels: auto (batches >= 1024), on, off (interpreted)')
    args = ap.parse_args()
    global JIT_MODE, HEADLINE
    JIT_MODE = args.jit
    HEADLINE = args.workload
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback. "
                         "Use --impl reference for the host-CPU arm.")
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    dist_on = world > 1
    if dist_on:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # stdout must carry exactly the one JSON line: NCCL's version banner (NCCL_DEBUG =
        # VERSION / WARN, the image default) goes to stdout, so it is switched off unless
        # the caller asked for INFO / TRACE output, which is sent to stderr
        if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION', 'WARN'):
            os.environ['NCCL_DEBUG'] = 'NONE'
        else:
            os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        # (every rank must issue the same collectives in the same order: all loop counts below are
        # fixed numbers, never time based; a short timeout turns a mismatch into a quick failure)
        import datetime
        dist.init_process_group('nccl', device_id=device, timeout=datetime.timedelta(seconds=180))
    warmup = max(3, args.warmup)
    timer = Timer(dist_on, device)
    sampler = ClockSampler(local)
    if not os.environ.get('AHK_NO_SAMPLER'):
        sampler.start()
    hbm_peak, hbm_src = peaks()
    probe = Runner('c2', B=1024)
    fp64_peak = measure_fp64_peak(probe.eng, timer)
    del probe
    head = bench_one(args.workload, rank, world, timer, sampler, args.steps, warmup, hbm_peak,
                     hbm_src, fp64_peak, dist_on, device, not args.no_cpu)
    cfgs = {}
    if not args.only:
        for name in ('c1', 'c3', 'c4', 'c5', 'c2'):
            if name == args.workload:
                continue
            heavy = name in ('c3', 'c4', 'c5')
            cfgs[name] = bench_one(name, rank, world, timer, sampler,
                                   min(args.steps, 3 if heavy else 10), 3, hbm_peak, hbm_src,
                                   fp64_peak, dist_on, device, not args.no_cpu)
        head["configs"] = cfgs
    head["clocks"] = sampler.summary()     # sampled during every timed region of this run
    sampler.stop.set()
    head["gpu"] = torch.cuda.get_device_name(local)
    if dist_on:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(head), flush=True)


if __name__ == '__main__':
    main()
