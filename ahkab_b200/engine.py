"""Engine: one compiled topology on one GPU, called through the C-ABI.

PyTorch is the plumbing only — device buffers, pinned staging, streams; the
arithmetic happens in libahkab_b200.so (hand-written sm_100a CUDA).  There is
no CPU path: constructing an Engine without a CUDA device raises."""
import ctypes as C

import numpy as np
import torch

from . import _abi, options as _options
from . import compile as _compile


class EngineError(RuntimeError):
    pass


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Engine(object):
    """engine = Engine(circ, batch); engine.op() / .dc_sweep() / .tran() / .ac()

    Results are dicts of torch tensors on the device, batch axis first."""

    def __init__(self, circ, batch=None, opts=None, device=None, flat=None):
        if not torch.cuda.is_available():
            raise EngineError("ahkab_b200 needs a CUDA device (B200); there is "
                              "no CPU fallback")
        self.lib = _abi.load_library()
        self.device = torch.device(device if device is not None else
                                   'cuda:%d' % torch.cuda.current_device())
        self.flat = flat if flat is not None else _compile.flatten(circ, batch)
        self.circ = circ
        self.n = self.flat.n
        self.B = self.flat.B if self.flat.B is not None else 1
        self._desc = _abi.CircuitDesc(self.flat)
        o = _options.snapshot()
        if opts:
            o.update(opts)
        self.opts = o
        self._opts = _abi.make_options(o)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.ahk_create(C.byref(self._desc.c), C.byref(self._opts),
                                     C.byref(h))
        if rc != 0:
            raise EngineError(self._err())
        self._h = h
        self._slots = np.ascontiguousarray(self.flat.vary_slots, np.int32)
        self._vary_host = None
        self.vary = None
        self.set_vary_flat(self.flat)

    # -- plumbing ---------------------------------------------------------------
    def _err(self):
        return (self.lib.ahk_last_error() or b'').decode()

    def __del__(self):
        h = getattr(self, '_h', None)
        if h is not None and h.value:
            try:
                self.lib.ahk_destroy(h)
            except Exception:
                pass
            self._h = None

    def set_vary(self, vary, non_blocking=True):
        """Upload per-instance values [n_vary][B] (numpy, or a pinned / device
        torch tensor).  This is the H2D copy of an end-to-end step."""
        staged = None
        if isinstance(vary, np.ndarray):
            # two pinned staging buffers used in turn, each guarded by an event recorded
            # after its H2D copy: a buffer is rewritten only once the copy that read it
            # last has completed (queued set_vary / op() loops stay correct)
            st = getattr(self, '_stage', None)
            if st is None or tuple(st[0][0].shape) != vary.shape:
                st = [[torch.empty(vary.shape, dtype=torch.float64, pin_memory=True), None]
                      for _ in range(2)]
                self._stage, self._stage_i = st, 0
            self._stage_i ^= 1
            staged = st[self._stage_i]
            if staged[1] is not None:
                staged[1].synchronize()
            staged[0].numpy()[...] = vary
            vary = self._vary_host = staged[0]
        if vary.device.type != 'cuda':
            if self.vary is None or self.vary.shape != vary.shape:
                self.vary = torch.empty(vary.shape, dtype=torch.float64,
                                        device=self.device)
            self.vary.copy_(vary, non_blocking=non_blocking)
            if staged is not None:
                with torch.cuda.device(self.device):
                    staged[1] = torch.cuda.Event()
                    staged[1].record(torch.cuda.current_stream(self.device))
        else:
            self.vary = vary.contiguous()
        if self.vary.numel():
            self.B = self.vary.shape[1]
        return self

    def set_vary_flat(self, flat):
        """Upload the per-instance values of a flattened (circuit, batch).  The flattening is
        cached on the batch (compile.flatten) and keeps a PINNED copy of its value matrix,
        made here on first use: repeated runs on the same batch copy host->device straight
        from it, with no staging copy on the host."""
        rows = flat.vary_rows
        if not len(rows):
            return self.set_vary(flat.vary)
        if getattr(flat, 'vary_pinned', None) is None:
            flat.uploads += 1
            if flat.uploads < 2:
                # first upload of this batch: straight into the engine's rotating pinned staging,
                # row by row (no stacked copy, no new pinned allocation — a Monte-Carlo loop that
                # builds a fresh batch per call never gets past this branch)
                return self._set_vary_rows(rows)
            flat.vary_pinned = torch.from_numpy(flat.vary).pin_memory()
        return self.set_vary(flat.vary_pinned)

    def _set_vary_rows(self, rows):
        shape = (len(rows), len(rows[0]))
        st = getattr(self, '_stage', None)
        if st is None or tuple(st[0][0].shape) != shape:
            st = [[torch.empty(shape, dtype=torch.float64, pin_memory=True), None] for _ in range(2)]
            self._stage, self._stage_i = st, 0
        self._stage_i ^= 1
        staged = st[self._stage_i]
        if staged[1] is not None:
            staged[1].synchronize()
        host = staged[0].numpy()
        for i, r in enumerate(rows):
            host[i] = r
        if self.vary is None or tuple(self.vary.shape) != shape:
            self.vary = torch.empty(shape, dtype=torch.float64, device=self.device)
        self.vary.copy_(staged[0], non_blocking=True)
        with torch.cuda.device(self.device):
            staged[1] = torch.cuda.Event()
            staged[1].record(torch.cuda.current_stream(self.device))
        self._vary_host = staged[0]
        self.B = shape[1]
        return self

    def set_options(self, **kw):
        self.opts.update(kw)
        self._opts = _abi.make_options(self.opts)
        if self.lib.ahk_set_options(self._h, C.byref(self._opts)) != 0:
            raise EngineError(self._err())

    def set_group(self, lanes):
        if self.lib.ahk_set_group(self._h, int(lanes)) != 0:
            raise EngineError(self._err())

    def set_variant(self, real=-1, ac=-1):
        """Force a kernel variant (ids from _abi.variants()); -1 = heuristic."""
        if self.lib.ahk_set_variant(self._h, int(real), int(ac)) != 0:
            raise EngineError(self._err())

    JIT_MODES = {'off': 0, 'on': 1, 'auto': 2}

    def set_jit(self, mode='auto', min_batch=0):
        """Circuit-specialised kernels (include/ahkab_b200.h: ahk_set_jit): 'off' = always
        the interpreted kernels, 'on' = always a kernel compiled for this circuit (NVRTC,
        seconds on first use), 'auto' = for batches of at least min_batch instances."""
        m = self.JIT_MODES[mode] if isinstance(mode, str) else int(mode)
        if self.lib.ahk_set_jit(self._h, m, int(min_batch)) != 0:
            raise EngineError(self._err())
        return self

    def set_jit_shape(self, rows_per_lane=0, lanes=0):
        """Shape of the specialised OP / DC / transient kernels (ahk_set_jit_shape);
        (0, 0) = the heuristic's variant."""
        if self.lib.ahk_set_jit_shape(self._h, int(rows_per_lane), int(lanes)) != 0:
            raise EngineError(self._err())
        return self

    def set_jit_tuning(self, ilp=0, threads=0, min_blocks=0):
        """ahk_set_jit_tuning: interleaved devices per lane, block size, resident blocks per
        SM of the specialised kernels (0 = heuristic)."""
        if self.lib.ahk_set_jit_tuning(self._h, int(ilp), int(threads), int(min_blocks)) != 0:
            raise EngineError(self._err())
        return self

    def set_slu(self, mode=-1):
        """ahk_set_jit_slu: static-schedule LU of the specialised kernels (-1 default: transients,
        0 off, 1 transients, 2 OP / DC too)."""
        if self.lib.ahk_set_jit_slu(self._h, int(mode)) != 0:
            raise EngineError(self._err())
        return self

    def slu_stats(self):
        """ahk_jit_slu_stats as a dict: calibrations, replaced, last_off, last_solves."""
        out = (C.c_int32 * 4)()
        if self.lib.ahk_jit_slu_stats(self._h, out) != 0:
            raise EngineError(self._err())
        return dict(zip(('calibrations', 'replaced', 'last_off', 'last_solves'), list(out)))

    def last_launch(self):
        """ahk_last_launch as a dict: vid, NC, R, G, threads, blocks, jit, analysis."""
        out = (C.c_int32 * 8)()
        if self.lib.ahk_last_launch(self._h, out) != 0:
            raise EngineError(self._err())
        return dict(zip(('vid', 'NC', 'R', 'G', 'threads', 'blocks', 'jit', 'analysis'), list(out)))

    def big_stats(self):
        """ahk_big_stats as a dict (n > 64 linear circuits: path taken, executed flop counts)."""
        out = (C.c_int64 * 8)()
        if self.lib.ahk_big_stats(self._h, out) != 0:
            raise EngineError(self._err())
        d = dict(zip(('path', 'factor_flop', 'point_flop', 'nf', 'npos', 'nslot', 'n', '_'), list(out)))
        d['path'] = {-1: None, 0: 'dense workspace', 1: 'sparse dynamic', 2: 'static schedule',
                     3: 'static schedule, compiled'}[d['path']]
        del d['_']
        return d

    def jit_launch_count(self):
        return int(self.lib.ahk_jit_launch_count())

    def _batch(self):
        b = _abi.ahk_batch()
        b.B = self.B
        b.n_vary = len(self._slots)
        b.vary_slots = self._slots.ctypes.data_as(C.POINTER(C.c_int32))
        b.vary = self.vary.data_ptr() if self.vary.numel() else None
        return b

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _new(self, shape, dtype=torch.float64):
        return torch.empty(shape, dtype=dtype, device=self.device)

    _ES = {torch.float64: 8, torch.int32: 4}

    def _new_packed(self, spec, cache=True):
        """Result tensors of one call as views into ONE device buffer, so that the
        device->host read of an end-to-end step is a single copy.  spec: tuple of (key,
        shape, dtype); float64 first keeps every view 8-byte aligned.  The byte layout is
        cached per spec (this sits on the launch path of sub-100-us kernels)."""
        lay = self._layouts.get(spec) if hasattr(self, '_layouts') else None
        if lay is None:
            if not hasattr(self, '_layouts'):
                self._layouts = {}
            items, off = [], 0
            for k, shape, dtype in spec:
                n_el = int(np.prod(shape))
                nb = (n_el * self._ES[dtype] + 7) & ~7
                items.append((k, tuple(shape), dtype, off, n_el * self._ES[dtype], nb))
                off += nb
            lay = (off, items)
            if cache:
                self._layouts[spec] = lay
        total, items = lay
        base = torch.empty(total, dtype=torch.uint8, device=self.device)
        out = {}
        for k, shape, dtype, off, nbytes, _nb in items:
            out[k] = base[off:off + nbytes].view(dtype).view(shape)
        out['_packed'] = (base, items)
        return out

    def _x0(self, x0):
        if x0 is None:
            return None
        x0 = torch.as_tensor(x0, dtype=torch.float64, device=self.device)
        if x0.dim() == 1 or (x0.dim() == 2 and x0.shape[0] != self.B and
                             x0.shape[-1] == 1):
            x0 = x0.reshape(1, -1).expand(self.B, self.n)
        x0 = x0.reshape(self.B, self.n).contiguous()
        return x0

    def elem_index(self, part_id):
        for i, e in enumerate(self.circ):
            if e.part_id.lower() == part_id.lower():
                return i
        raise ValueError("element %s not found" % part_id)

    # -- analyses -----------------------------------------------------------------
    def op(self, x0=None, use_guess=True, want_resid=True):
        x0 = self._x0(x0)
        spec = (('x', (self.B, self.n), torch.float64),)
        if want_resid:
            spec += (('resid', (self.B, self.n), torch.float64),)
        spec += (('iters', (self.B,), torch.int32), ('status', (self.B,), torch.int32))
        pk = self._new_packed(spec)
        x, resid, iters, status = pk['x'], pk.get('resid'), pk['iters'], pk['status']
        b = self._batch()
        with torch.cuda.device(self.device):
            rc = self.lib.ahk_op(self._h, C.byref(b), _ptr(x0),
                                 int(bool(use_guess)), _ptr(x), _ptr(resid),
                                 _ptr(iters), _ptr(status), self._stream())
        if rc != 0:
            raise EngineError(self._err())
        return dict(x=x, resid=resid, iters=iters, status=status, _packed=pk['_packed'])

    def op_host(self, vary_host=None, use_guess=True, want_resid=True, nchunks=4):
        """Operating point from HOST parameters to HOST results (ahk_op_host): the batch
        flows in chunks through copy-in / kernel / copy-out streams inside the library.
        vary_host: pinned [n_vary][B] float64 tensor (default: the batch the engine was
        built with).  Returns pinned host tensors; asynchronous on the current stream —
        synchronise before reading them.  The output buffers are cached per batch size and
        reused by the next call (copy what must survive it)."""
        if vary_host is None:
            if self._vary_host is None:
                self._vary_host = torch.empty(self.vary.shape, dtype=torch.float64, pin_memory=True)
                self._vary_host.copy_(self.vary)
                torch.cuda.synchronize(self.device)
            vary_host = self._vary_host
        B = vary_host.shape[1] if vary_host.numel() else self.B
        key = ('op_host', B, bool(want_resid))
        if not hasattr(self, '_pinned'):
            self._pinned = {}
        out = self._pinned.get(key)
        if out is None:
            out = dict(x=torch.empty((B, self.n), dtype=torch.float64, pin_memory=True),
                       iters=torch.empty((B,), dtype=torch.int32, pin_memory=True),
                       status=torch.empty((B,), dtype=torch.int32, pin_memory=True))
            if want_resid:
                out['resid'] = torch.empty((B, self.n), dtype=torch.float64, pin_memory=True)
            self._pinned[key] = out
        with torch.cuda.device(self.device):
            rc = self.lib.ahk_op_host(
                self._h, B, len(self._slots), self._slots.ctypes.data_as(C.POINTER(C.c_int32)),
                _ptr(vary_host), int(bool(use_guess)), _ptr(out['x']), _ptr(out.get('resid')),
                _ptr(out['iters']), _ptr(out['status']), int(nchunks), self._stream())
        if rc != 0:
            raise EngineError(self._err())
        return dict(out)

    def dc_sweep(self, source, sweep_values, x0=None, use_guess=True):
        src = source if isinstance(source, int) else self.elem_index(source)
        sweep = np.ascontiguousarray(sweep_values, np.float64)
        P = len(sweep)
        x0 = self._x0(x0)
        pk = self._new_packed((('x', (self.B, P, self.n), torch.float64),
                               ('iters', (self.B, P), torch.int32),
                               ('status', (self.B, P), torch.int32)))
        x, iters, status = pk['x'], pk['iters'], pk['status']
        b = self._batch()
        with torch.cuda.device(self.device):
            rc = self.lib.ahk_dc_sweep(
                self._h, C.byref(b), src,
                sweep.ctypes.data_as(C.POINTER(C.c_double)), P, _ptr(x0),
                int(bool(use_guess)), _ptr(x), _ptr(iters), _ptr(status),
                self._stream())
        if rc != 0:
            raise EngineError(self._err())
        return dict(x=x, iters=iters, status=status, sweep=sweep, _packed=pk['_packed'])

    def tran(self, x0, tstart, tstep, tstop, method='TRAP',
             use_step_control=True, max_rows=None):
        method = method.upper()
        if method not in _abi.METHODS:
            raise ValueError("unknown method %s" % method)
        if max_rows is None:
            max_rows = self._cap_max_rows(self.default_max_rows(tstart, tstep, tstop,
                                                                use_step_control))
        x0 = self._x0(x0)
        pk = self._new_packed((('t', (self.B, max_rows), torch.float64),
                               ('x', (self.B, max_rows, self.n), torch.float64),
                               ('rows', (self.B,), torch.int32),
                               ('counters', (self.B, 4), torch.int32),
                               ('status', (self.B,), torch.int32)))
        t, x, rows, counters, status = pk['t'], pk['x'], pk['rows'], pk['counters'], pk['status']
        b = self._batch()
        with torch.cuda.device(self.device):
            rc = self.lib.ahk_tran(
                self._h, C.byref(b), _ptr(x0), float(tstart), float(tstep),
                float(tstop), _abi.METHODS[method], int(bool(use_step_control)),
                int(max_rows), _ptr(t), _ptr(x), _ptr(rows), _ptr(counters),
                _ptr(status), self._stream())
        if rc != 0:
            raise EngineError(self._err())
        return dict(t=t, x=x, rows=rows, counters=counters, status=status, _packed=pk['_packed'])

    def pack_tran(self, raw):
        """Ragged transient results -> contiguous (ahk_pack_rows): returns a dict with
        t [sum rows], x [sum rows][n], offsets [B + 1] (int64, instance b owns rows
        offsets[b]:offsets[b+1]) plus rows / counters / status.  Reads the row total back
        (one small synchronising copy); the device->host read of the result then moves only
        the rows that exist instead of B x max_rows.  Everything but `offsets` (= cumsum of
        rows) lives in ONE device buffer (`_packed`): one D2H copy, one message of the
        multi-GPU result gather."""
        rows = raw['rows']
        B, max_rows = raw['t'].shape
        offs = torch.zeros(B + 1, dtype=torch.int64, device=self.device)
        torch.cumsum(rows, 0, out=offs[1:])
        total = int(offs[-1].item())
        spec = (('t', (max(total, 1),), torch.float64), ('x', (max(total, 1), self.n), torch.float64),
                ('rows', (B,), torch.int32), ('counters', (B, 4), torch.int32),
                ('status', (B,), torch.int32))
        if 'op_iters' in raw:
            spec += (('op_iters', (B,), torch.int32),)
        pk = self._new_packed(spec, cache=False)
        t, x = pk['t'], pk['x']
        with torch.cuda.device(self.device):
            rc = self.lib.ahk_pack_rows(_ptr(raw['t']), _ptr(raw['x']), _ptr(rows), _ptr(offs), B,
                                        max_rows, self.n, _ptr(t), _ptr(x), self._stream())
        if rc != 0:
            raise EngineError(self._err())
        for k in ('rows', 'counters', 'status', 'op_iters'):
            if k in pk:
                pk[k].copy_(raw[k])
        out = dict(t=t[:total], x=x[:total], offsets=offs, rows=pk['rows'], counters=pk['counters'],
                   status=pk['status'], _packed=pk['_packed'])
        if 'op_iters' in pk:
            out['op_iters'] = pk['op_iters']
        return out

    @staticmethod
    def default_max_rows(tstart, tstep, tstop, use_step_control):
        nominal = int(np.ceil((tstop - tstart) / tstep)) + 2 if tstep > 0 else 16
        if not use_step_control:
            return nominal + 2
        # LTE control starts at (tstop-tstart)/9999 and may at most double per
        # step; 64x the nominal count (min 4096) covers stiff oscillators, the
        # ROWS_FULL status bit reports the rare overflow.
        return max(4096, 64 * nominal)

    def _cap_max_rows(self, max_rows):
        """Row buffers are [B][max_rows][n + 1] doubles: keep the default inside a quarter of
        the device memory (the ROWS_FULL status bit reports an overflow)."""
        try:
            free, total = torch.cuda.mem_get_info(self.device)
        except Exception:
            return max_rows
        per_row = 8.0 * (self.n + 1) * max(1, self.B)
        return int(max(16, min(max_rows, (0.25 * total) // per_row)))

    def ac(self, omegas, x_op=None):
        om = np.ascontiguousarray(omegas, np.float64)
        P = len(om)
        x_op = self._x0(x_op)
        pk = self._new_packed((('x', (self.B, P, self.n, 2), torch.float64),
                               ('status', (self.B,), torch.int32)))
        xo, status = pk['x'], pk['status']
        b = self._batch()
        with torch.cuda.device(self.device):
            rc = self.lib.ahk_ac(self._h, C.byref(b), _ptr(x_op),
                                 om.ctypes.data_as(C.POINTER(C.c_double)), P,
                                 _ptr(xo), _ptr(status), self._stream())
        if rc != 0:
            raise EngineError(self._err())
        return dict(x=torch.view_as_complex(xo), status=status, omegas=om, _packed=pk['_packed'])

    def ac_host(self, omegas, x_op=None, vary_host=None):
        """AC sweep from HOST parameters to HOST results without copies: the kernel reads the pinned
        parameter matrix and writes the pinned result rows itself (unified addressing), so the
        PCIe traffic overlaps the computation inside the launch — what `op_host(nchunks=0)` is
        for the operating point.  vary_host: pinned [n_vary][B] float64 tensor (default: a pinned
        copy of the batch the engine was built with).  Returns pinned host tensors, cached per
        shape and reused by the next call; asynchronous on the current stream."""
        om = np.ascontiguousarray(omegas, np.float64)
        P = len(om)
        if vary_host is None:
            if self._vary_host is None:
                self._vary_host = torch.empty(self.vary.shape, dtype=torch.float64, pin_memory=True)
                self._vary_host.copy_(self.vary)
                torch.cuda.synchronize(self.device)
            vary_host = self._vary_host
        B = vary_host.shape[1] if vary_host.numel() else self.B
        x_op = self._x0(x_op)
        if not hasattr(self, '_pinned'):
            self._pinned = {}
        key = ('ac_host', B, P)
        out = self._pinned.get(key)
        if out is None:
            out = dict(x=torch.empty((B, P, self.n, 2), dtype=torch.float64, pin_memory=True),
                       status=torch.empty((B,), dtype=torch.int32, pin_memory=True))
            self._pinned[key] = out
        b = _abi.ahk_batch()
        b.B = B
        b.n_vary = len(self._slots)
        b.vary_slots = self._slots.ctypes.data_as(C.POINTER(C.c_int32))
        b.vary = vary_host.data_ptr() if vary_host.numel() else None
        with torch.cuda.device(self.device):
            rc = self.lib.ahk_ac(self._h, C.byref(b), _ptr(x_op),
                                 om.ctypes.data_as(C.POINTER(C.c_double)), P,
                                 _ptr(out['x']), _ptr(out['status']), self._stream())
        if rc != 0:
            raise EngineError(self._err())
        return dict(x=torch.view_as_complex(out['x']), status=out['status'], omegas=om)

    def to_host(self, raw, keys=None, fresh=False):
        """Device->host read of a result dict into PINNED buffers (asynchronous on the
        current stream; synchronise before reading).  This is the D2H copy of an
        end-to-end step.

        ALIASING: by default the pinned buffers are cached per result shape and REUSED —
        the tensors returned by an earlier call change under the caller at the next call
        with the same shapes (what a benchmark loop wants).  fresh=True allocates new
        pinned buffers that belong to the caller."""
        if not hasattr(self, '_pinned'):
            self._pinned = {}
        if fresh:
            saved, self._pinned = self._pinned, {}
            try:
                return self.to_host(raw, keys)
            finally:
                self._pinned = saved
        out = {}
        if keys is None and raw.get('_packed') is not None:
            base, layout = raw['_packed']             # one copy for the whole result set
            hb = self.pinned_bytes(base.numel())
            hb.copy_(base, non_blocking=True)
            for k, shape, dtype, off, nbytes, _nb in layout:
                out[k] = hb[off:off + nbytes].view(dtype).view(shape)
            for k, v in raw.items():                  # tensors that are not part of the pack
                if torch.is_tensor(v) and k not in out and k != '_packed':
                    out[k] = v.cpu()
            return out
        for k, v in raw.items():
            if not torch.is_tensor(v) or (keys is not None and k not in keys):
                continue
            src = torch.view_as_real(v) if v.is_complex() else v
            key = (k, tuple(src.shape), src.dtype)
            buf = self._pinned.get(key)
            if buf is None:
                buf = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
                self._pinned[key] = buf
            buf.copy_(src, non_blocking=True)
            out[k] = buf
        return out

    def to_host_owned(self, raw):
        """The whole packed result set of one call in ONE device->host copy, returned as
        numpy arrays that OWN their pinned buffer: it comes from a per-engine pool and goes
        back to it when the last array / view of this result set is garbage collected, so
        results of earlier calls never change under the caller (unlike `to_host`, whose
        cached buffers suit a benchmark loop) and no memory is pinned per call.
        Synchronous: the arrays are valid on return."""
        import weakref
        base, layout = raw['_packed']
        nbytes = int(base.numel())
        pool = self.__dict__.setdefault('_host_pool', [])
        hb = None
        for i, t in enumerate(pool):            # smallest free buffer that fits
            if t.numel() >= nbytes and (hb is None or t.numel() < hb.numel()):
                hb, at = t, i
        if hb is not None:
            pool.pop(at)
        else:
            hb = torch.empty(max(4096, int(nbytes * 1.125)), dtype=torch.uint8, pin_memory=True)
        with torch.cuda.device(self.device):
            hb[:nbytes].copy_(base, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
        arr = hb.numpy()                        # every view below keeps `arr` alive as its base

        def give_back(pool=pool, hb=hb):
            if len(pool) < 8:
                pool.append(hb)
        weakref.finalize(arr, give_back)
        npd = {torch.float64: np.float64, torch.int32: np.int32}
        out = {}
        for k, shape, dtype, off, nb, _nb in layout:
            out[k] = arr[off:off + nb].view(npd[dtype]).reshape(shape)
        return out

    def pinned_bytes(self, nbytes):
        """A pinned host buffer of nbytes (uint8) out of ONE cached, grow-only allocation
        (pinning memory costs milliseconds: never per call)."""
        if not hasattr(self, '_pinned'):
            self._pinned = {}
        hb = self._pinned.get('_bytes')
        if hb is None or hb.numel() < nbytes:
            hb = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, pin_memory=True)
            self._pinned['_bytes'] = hb
        return hb[:nbytes]

    def probe_fp64(self, iters=4096, blocks=None, threads=256):
        """Launch the FP64-FMA probe kernel (diagnostics); returns its flop count."""
        blocks = blocks or 148 * 8
        if not hasattr(self, '_probe_out'):
            self._probe_out = self._new((8,))
        rc = self.lib.ahk_probe_fp64(int(blocks), int(threads), int(iters),
                                     _ptr(self._probe_out), self._stream())
        if rc != 0:
            raise EngineError(self._err())
        return float(blocks) * threads * iters * 16.0

    def launch_count(self):
        return int(self.lib.ahk_launch_count())
