"""Chunked, pipelined end-to-end execution: host parameters in, host results out.

The batch axis is cut into chunks; three CUDA streams form the classic pipeline

    copy-in  (H2D of chunk c+1's parameters)
    compute  (the analysis kernels of chunk c — always ONE stream, so the engine's
              per-circuit device workspaces are never shared by two running kernels)
    copy-out (D2H of chunk c-1's results into pinned host memory)

so that PCIe traffic in both directions overlaps the kernels.  Instances are
independent: chunking changes no result.  This is host plumbing
around the same C-ABI calls; every chunk is an ordinary sub-batch."""
import copy
import ctypes as C

import torch


class _SubEngine(object):
    """A view of an Engine on a contiguous sub-batch with its own `vary` buffer."""

    def __init__(self, eng, B, vary):
        self.__dict__.update(eng.__dict__)
        self.B = B
        self.vary = vary
        self._pinned = {}
        self._layouts = {}
        self._h_owner = eng          # keeps the parent (and the C handle) alive

    def __getattr__(self, name):     # methods / class attributes come from the parent's class
        a = getattr(type(self._h_owner), name)
        return a.__get__(self) if hasattr(a, '__get__') and callable(a) else a

    def __del__(self):
        pass


class Pipeline(object):
    def __init__(self, eng, nchunks=4):
        self.eng = eng
        self.nchunks = int(nchunks)
        dev = eng.device
        self.s_in = torch.cuda.Stream(dev)
        self.s_comp = torch.cuda.Stream(dev)
        self.s_out = torch.cuda.Stream(dev)
        self._vbuf = {}
        self._host = {}

    def _chunks(self, B):
        k = max(1, min(self.nchunks, B))
        base, rem = divmod(B, k)
        lo = 0
        for i in range(k):
            hi = lo + base + (1 if i < rem else 0)
            yield i, lo, hi
            lo = hi

    def _host_out(self, key, shape, dtype):
        k = (key, tuple(shape), dtype)
        buf = self._host.get(k)
        if buf is None:
            buf = torch.empty(shape, dtype=dtype, pin_memory=True)
            self._host[k] = buf
        return buf

    def run(self, step, vary_host, extra=None):
        """step(sub_engine, lo, hi, extra_dev) -> dict of device tensors with the
        sub-batch on axis 0.  vary_host: pinned [n_vary][B] float64.  extra: optional
        dict of pinned host tensors with the batch on axis 0 (e.g. x0), sliced and
        uploaded per chunk.  Returns a dict of pinned host tensors (full batch);
        synchronises before returning."""
        eng = self.eng
        B = vary_host.shape[1] if vary_host.numel() else eng.B
        n_vary = vary_host.shape[0]
        cur = torch.cuda.current_stream(eng.device)
        for s in (self.s_in, self.s_comp, self.s_out):
            s.wait_stream(cur)
        out = {}
        free_ev = {}
        for i, lo, hi in self._chunks(B):
            Bc = hi - lo
            slot = i % 2
            key = (slot, Bc)
            with torch.cuda.stream(self.s_in):
                vb = self._vbuf.get(key)
                if vb is None:
                    vb = torch.empty((n_vary, Bc), dtype=torch.float64, device=eng.device)
                    self._vbuf[key] = vb
                if slot in free_ev:                 # the buffer's previous chunk has computed
                    self.s_in.wait_event(free_ev[slot])
                if n_vary:
                    rc = eng.lib.ahk_h2d_columns(C.c_void_p(vb.data_ptr()),
                                                 C.c_void_p(vary_host.data_ptr()), n_vary, B, lo, hi,
                                                 C.c_void_p(self.s_in.cuda_stream))
                    if rc != 0:
                        raise RuntimeError(eng._err())
                xdev = {}
                for k_, v_ in (extra or {}).items():
                    xdev[k_] = v_[lo:hi].to(eng.device, non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(self.s_in)
            with torch.cuda.stream(self.s_comp):
                self.s_comp.wait_event(ev_in)
                sub = _SubEngine(eng, Bc, vb)
                raw = step(sub, lo, hi, xdev)
                ev_c = torch.cuda.Event()
                ev_c.record(self.s_comp)
                free_ev[slot] = ev_c
                for v_ in xdev.values():
                    v_.record_stream(self.s_comp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_c)
                for k_, v_ in raw.items():
                    if not torch.is_tensor(v_):
                        continue
                    src = torch.view_as_real(v_) if v_.is_complex() else v_
                    src.record_stream(self.s_out)
                    dst = self._host_out(k_, (B,) + tuple(src.shape[1:]), src.dtype)
                    dst[lo:hi].copy_(src, non_blocking=True)
                    out[k_] = dst
        cur.wait_stream(self.s_out)
        cur.wait_stream(self.s_comp)
        self.s_out.synchronize()
        return out
