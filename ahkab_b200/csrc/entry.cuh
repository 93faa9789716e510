// entry.cuh — what one group does for its instance, per analysis: load x0,
// run the analysis from core.cuh, store results.  Shared by the CUDA kernels
// (engine.cu) and the host-emulation test harness (tests/hostemu).
#pragma once
#include "core.cuh"

namespace ahk {

// OP = a sweep of one point with no swept source (src_elem < 0, sweep == nullptr)
struct DcArgs {
  int src_elem; const double* sweep; int npts;
  const double* x0; int use_guess;
  double* x_out; double* resid; int* iters; int* status;
};

template <class V>
AHK_DEV void init_ctx(Ctx& c) {
  c.singular = 0; c.dc_src = -1; c.dc_val = 0.0; c.C1 = 0.0; c.gadd = 0.0; c.G = V::G;
  if constexpr (V::NC == 0) {
    if (c.L->big) big_init<V::G>(c, c.Ag, (size_t)c.p->n * c.L->RS);   // large n: matrix zero, no entry flagged
  }
}

// dc_analysis.op_analysis (one point) / dc_analysis.dc_analysis for one instance: the
// full OP per sweep value, warm-started from the previous point (dc_analysis.py:563-581)
template <class V>
AHK_DEV void run_dc(Ctx& c, const DcArgs& a) {
  const int n = c.p->n, G = V::G;
  init_ctx<V>(c);
  build_values<V>(c, false);
  double* x = c.x();
  bool have = a.x0 != nullptr;
  if (have) for (int i = c.sub; i < n; i += G) x[i] = a.x0[c.inst * n + i];
  Grp<V::G>::sync(c.mask);
  c.dc_src = a.src_elem;
  for (int k = 0; k < a.npts; k++) {
    if (a.src_elem >= 0) c.dc_val = a.sweep[k];
    int it = 0;
    int st = op_analysis<V>(c, have, a.use_guess != 0, &it);
    long long o = c.inst * a.npts + k;
    const double* res = c.sm + c.L->res;
    for (int i = c.sub; i < n; i += G) {
      a.x_out[o * n + i] = x[i];
      if (a.resid) a.resid[o * n + i] = (st & AHK_ST_OK) ? res[i] : nan("");
    }
    if (c.sub == 0) { a.iters[o] = it; a.status[o] = st; }
    have = (st & AHK_ST_OK) != 0;   // a failed point returns None -> next one restarts
    Grp<V::G>::sync(c.mask);
  }
}

#ifdef __CUDACC__
// Operating point of the one-lane specialised kernel with BLOCK-COOPERATIVE result rows: a lane
// that writes its own x[0..n) row puts 8-byte pieces 8 n bytes apart — partial sectors, which is
// what made the zero-copy form of ahk_op_host (results written straight into the caller's pinned
// buffers over PCIe) four times slower than a packed copy.  x and the residual already sit in the
// lanes' shared-memory slices: after a block barrier consecutive THREADS write consecutive doubles
// of the block's [instances][n] rows (256 contiguous bytes per warp store).  Warp-uniform: every
// lane runs (`live` false for the padding lanes of the last block, which repeat the last instance
// and write nothing).  One point (OP) only; sweeps keep run_dc.
template <class V>
__device__ __forceinline__ void run_op_rows(Ctx& c, const DcArgs& a, bool live) {
  extern __shared__ __align__(16) double smem[];
  const int n = c.p->n;
  init_ctx<V>(c);
  build_values<V>(c, false);
  double* x = c.x();
  const bool have = a.x0 != nullptr;
  if (have) for (int i = 0; i < n; i++) x[i] = a.x0[c.inst * n + i];
  c.dc_src = a.src_elem;
  if (a.src_elem >= 0) c.dc_val = a.sweep[0];
  int it = 0;
  const int st = op_analysis<V>(c, have, a.use_guess != 0, &it);
  double* res = c.sm + c.L->res;
  if (a.resid && !(st & AHK_ST_OK)) for (int i = 0; i < n; i++) res[i] = nan("");
  if (live) { a.iters[c.inst] = it; a.status[c.inst] = st; }
  __syncthreads();
  const long long first = (long long)blockIdx.x * blockDim.x;
  const long long left = c.B - first;
  const int cnt = (int)(left < (long long)blockDim.x ? left : (long long)blockDim.x);   // instances of this block
  const int stride = c.L->stride, ox = c.L->x, ores = c.L->res;
  for (int j = threadIdx.x; j < cnt * n; j += blockDim.x) {
    const int il = j / n, i = j - il * n;
    a.x_out[first * n + j] = smem[il * stride + ox + i];
    if (a.resid) a.resid[first * n + j] = smem[il * stride + ores + i];
  }
}
#endif

// transient.transient_analysis for one instance
// (warp-uniform: EVERY lane of the warp calls it — `live` = false for the padding groups of a
// partial last warp, whose c.inst names some valid instance to read parameters from)
template <class V>
AHK_DEV void run_tran(Ctx& c, const double* x0, const TranArgs& a, int* status, bool live = true) {
  const int n = c.p->n, G = V::G;
  init_ctx<V>(c);
  build_values<V>(c, true);
  double* x = c.x();
  for (int i = c.sub; i < n; i += G) x[i] = x0 ? x0[c.inst * n + i] : 0.0;
  Grp<V::G>::sync(c.mask);
  int st = tran_analysis<V>(c, a, live);
  if (live && c.sub == 0) status[c.inst] = st;
}

// kernel parameters of the circuit-specialised kernels (jit_kernels.cuh; filled by
// engine.cu): no program / options / layout — those are constants of the build
struct KDcJ { const double* vary; long long B; DcArgs a; int* pivrec; };
struct KTranJ { const double* vary; long long B; const double* x0; TranArgs a; int* status; int* pivrec; };

#ifdef __CUDACC__
// the group's context: its lanes, its instance, its shared-memory slice
template <int G>
__device__ __forceinline__ bool make_ctx(Ctx& c, const Prog* p, const Opts* o, const Layout* L,
                                         int stride, const double* vary, long long B,
                                         long long units, long long& unit) {
  extern __shared__ __align__(16) double smem[];
  const int gi = threadIdx.x / G;
  const int ipb = blockDim.x / G;
  unit = (long long)blockIdx.x * ipb + gi;
  if (unit >= units) return false;
  const int lane = threadIdx.x & 31;
  c.p = p; c.o = o; c.L = L;
  c.sm = smem + (size_t)gi * stride;
  c.vary = vary; c.B = B; c.inst = unit;
  c.sub = threadIdx.x % G;
  c.G = G;
  c.mask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
  return true;
}

// Context of a WARP-UNIFORM kernel (flat transient loop): no lane leaves — the padding groups
// of the last block run along on the last instance's parameters with *live = false — and the
// group's lane mask is the literal full mask, so that every shuffle / barrier of the solve
// compiles to its unguarded form.  The launch must use whole warps (blockDim % 32 == 0).
template <int G>
__device__ __forceinline__ void make_ctx_uniform(Ctx& c, const Prog* p, const Opts* o, const Layout* L,
                                                 int stride, const double* vary, long long B, bool& live) {
  extern __shared__ __align__(16) double smem[];
  const int gi = threadIdx.x / G;
  const int ipb = blockDim.x / G;
  const long long unit = (long long)blockIdx.x * ipb + gi;
  live = unit < B;
  c.p = p; c.o = o; c.L = L;
  c.sm = smem + (size_t)gi * stride;
  c.vary = vary; c.B = B; c.inst = live ? unit : B - 1;
  c.sub = threadIdx.x % G;
  c.G = G;
  c.mask = 0xffffffffu;
}

#endif

}  // namespace ahk
