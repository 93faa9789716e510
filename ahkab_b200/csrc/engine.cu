// engine.cu — CUDA kernels of the small-circuit engine and the C-ABI
// (include/ahkab_b200.h).  sm_100a only; no CPU fallback.
//
// One kernel per analysis; each is a persistent-per-instance kernel: a group of
// G lanes carries its instance through the whole analysis (every Newton
// iteration of both OP passes / every sweep point / every time step / every
// frequency) out of shared memory, touching HBM only for the parameter reads
// and the result rows.  See core.cuh for the per-instance algorithm.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "ac.cuh"
#include "biglin.cuh"
#include "entry.cuh"
#include "program_build.hpp"

using namespace ahk;

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(const std::string& m) { g_err = m; return -1; }

#define CK(call)                                                                     \
  do {                                                                               \
    cudaError_t e_ = (call);                                                         \
    if (e_ != cudaSuccess)                                                           \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e_));               \
  } while (0)

struct KBase {
  Prog p; Opts o; Layout L;
  const double* vary; long long B;
};
struct KOp { KBase b; OpArgs a; };
struct KDc { KBase b; DcArgs a; };
struct KTran { KBase b; const double* x0; TranArgs a; int* status; };
struct KAc { Prog p; Opts o; AcLayout AL; const double* vary; long long B; AcArgs a; int nchunks; };

template <int G>
__device__ __forceinline__ bool make_ctx(Ctx& c, const Prog* p, const Opts* o, const Layout* L,
                                         int stride, const double* vary, long long B,
                                         long long units, long long& unit) {
  extern __shared__ double smem[];
  const int gi = threadIdx.x / G;
  const int ipb = blockDim.x / G;
  unit = (long long)blockIdx.x * ipb + gi;
  if (unit >= units) return false;
  const int lane = threadIdx.x & 31;
  c.p = p; c.o = o; c.L = L;
  c.sm = smem + (size_t)gi * stride;
  c.vary = vary; c.B = B; c.inst = unit;
  c.sub = threadIdx.x % G;
  c.mask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
  return true;
}

template <int G>
__global__ void __launch_bounds__(128) k_op(const __grid_constant__ KOp k) {
  Ctx c; long long u;
  if (!make_ctx<G>(c, &k.b.p, &k.b.o, &k.b.L, k.b.L.stride, k.b.vary, k.b.B, k.b.B, u)) return;
  run_op<G>(c, k.a);
}

template <int G>
__global__ void __launch_bounds__(128) k_dc(const __grid_constant__ KDc k) {
  Ctx c; long long u;
  if (!make_ctx<G>(c, &k.b.p, &k.b.o, &k.b.L, k.b.L.stride, k.b.vary, k.b.B, k.b.B, u)) return;
  run_dc<G>(c, k.a);
}

template <int G>
__global__ void __launch_bounds__(128) k_tran(const __grid_constant__ KTran k) {
  Ctx c; long long u;
  if (!make_ctx<G>(c, &k.b.p, &k.b.o, &k.b.L, k.b.L.stride, k.b.vary, k.b.B, k.b.B, u)) return;
  run_tran<G>(c, k.x0, k.a, k.status);
}

// AC: unit = (instance, frequency chunk)
template <int G>
__global__ void __launch_bounds__(128) k_ac(const __grid_constant__ KAc k) {
  Ctx c; long long u;
  if (!make_ctx<G>(c, &k.p, &k.o, &k.AL.L, k.AL.stride, k.vary, k.B, k.B * k.nchunks, u)) return;
  c.inst = u / k.nchunks;
  const int ch = (int)(u % k.nchunks);
  AcArgs a = k.a;
  const int per = (a.npts + k.nchunks - 1) / k.nchunks;
  a.f0 = ch * per;
  a.f1 = min(a.npts, a.f0 + per);
  run_ac<G>(c, k.AL, a);
}

// FP64 pipe probe: 8 independent DFMA chains per thread, no memory traffic in the
// loop.  bench.py times it to get the measured FP64 roofline denominator
// (MEASURED_PEAKS.json has no FP64 figure).
__global__ void __launch_bounds__(256) k_probe_fp64(int iters, double seed, double* out) {
  double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5,
         a6 = seed + 6, a7 = seed + 7;
  const double m = 0.999999 + 1e-9 * threadIdx.x, c = 1e-7;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 12345.678) out[0] = s;   // keeps the chains live, (almost) never writes
}

__global__ void k_fill_i32(int* p, long long n, int v) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace

struct ahk_engine {
  HostProg H;
  Opts o;
  int n = 0;
  int group = 0;           // 0 = heuristic
  int device = 0;
  int sm_count = 148;
  size_t smem_max = 227 * 1024;
  void* blob = nullptr;    // device copy of the program tables
  size_t blob_bytes = 0;
  Prog dprog;              // device-pointer view
  int* d_slot_vary = nullptr;
  std::vector<int> cur_slots;
  bool slots_valid = false;
  double* d_scratch = nullptr;   // sweep values / omegas
  size_t scratch_cap = 0;
  BigLin big;              // large-n linear path (biglin.cuh)
};

int BigLin::run(ahk_engine* e, const ahk_batch* batch, int src_elem, const double* sweep_host,
                int npts, double* x_out, int32_t* iters, int32_t* status, double* resid,
                cudaStream_t s, std::string& err, std::atomic<long long>& launches) {
#define CKB(call)                                                                    \
  do {                                                                               \
    cudaError_t e_ = (call);                                                         \
    if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); return -1; } \
  } while (0)
  const Prog& h = e->H.hdr;
  const int n = h.n;
  if (n > 60000) { err = "n too large"; return -1; }
  const int RS = (n + 1) | 1;
  BigSmem S = big_smem(n, h.npos);
  if ((size_t)S.total > e->smem_max) { err = "circuit too large for the large-n path"; return -1; }
  int per_sm = (int)std::min<size_t>(16, e->smem_max / (size_t)S.total);
  int want = (int)std::min<long long>(batch->B, (long long)e->sm_count * per_sm);
  const int capn = n * (n - 1) / 2 + n;
  if (want > nwarps || capn != cap) {
    CKB(cudaStreamSynchronize(s));
    release();
    nwarps = want; cap = capn;
    CKB(cudaMalloc(&wsA, (size_t)nwarps * n * RS * sizeof(double)));
    CKB(cudaMalloc(&wsLv, (size_t)nwarps * 2 * cap * sizeof(double)));
    CKB(cudaMalloc(&wsUv, (size_t)nwarps * 2 * cap * sizeof(double)));
    CKB(cudaMalloc(&wsLi, (size_t)nwarps * 2 * cap * sizeof(unsigned short)));
    CKB(cudaMalloc(&wsUi, (size_t)nwarps * 2 * cap * sizeof(unsigned short)));
  }
  if (npts > sweep_cap) {
    CKB(cudaStreamSynchronize(s));
    cudaFree(d_sweep);
    sweep_cap = npts * 2 + 16;
    CKB(cudaMalloc(&d_sweep, sweep_cap * sizeof(double)));
  }
  CKB(cudaMemcpyAsync(d_sweep, sweep_host, npts * sizeof(double), cudaMemcpyHostToDevice, s));
  BigK k;
  k.p = e->dprog; k.o = e->o; k.vary = batch->vary; k.B = batch->B;
  k.src_elem = src_elem; k.sweep = d_sweep; k.npts = npts; k.x0 = nullptr;
  k.x_out = x_out; k.iters = iters; k.status = status; k.resid = resid;
  k.wsA = wsA; k.wsLv = wsLv; k.wsUv = wsUv; k.wsLi = wsLi; k.wsUi = wsUi;
  k.cap = cap; k.RS = RS;
  if (S.total > 48 * 1024)
    CKB(cudaFuncSetAttribute(k_biglin, cudaFuncAttributeMaxDynamicSharedMemorySize, S.total));
  k_biglin<<<want, 32, S.total, s>>>(k, S);
  CKB(cudaGetLastError());
  launches++;
  return 0;
#undef CKB
}

namespace {

template <typename T>
size_t put(std::vector<unsigned char>& buf, const std::vector<T>& v) {
  size_t off = (buf.size() + 15) & ~size_t(15);
  buf.resize(off + v.size() * sizeof(T) + 16);
  if (!v.empty()) std::memcpy(buf.data() + off, v.data(), v.size() * sizeof(T));
  return off;
}

int upload_program(ahk_engine* e) {
  HostProg& H = e->H;
  std::vector<unsigned char> buf;
  size_t o_row = put(buf, H.row_ptr), o_pcol = put(buf, H.pcol), o_prowi = put(buf, H.prowi), o_pdiag = put(buf, H.pdiag);
  size_t o_tp = put(buf, H.term_ptr), o_ts = put(buf, H.term_slot), o_tk = put(buf, H.term_kind);
  size_t o_nlp = put(buf, H.nl_ptr), o_nls = put(buf, H.nl_src), o_nlg = put(buf, H.nl_sgn);
  size_t o_txp = put(buf, H.tx_ptr), o_txs = put(buf, H.tx_src), o_txg = put(buf, H.tx_sgn);
  size_t o_lin = put(buf, H.lin), o_src = put(buf, H.src), o_dev = put(buf, H.dev);
  size_t o_lock = put(buf, H.lock), o_gG = put(buf, H.guessG), o_gs = put(buf, H.gslot);
  size_t o_gsc = put(buf, H.gscale), o_gc = put(buf, H.gconst), o_nom = put(buf, H.nominal);
  size_t o_sv = put(buf, H.slot_vary);
  CK(cudaMalloc(&e->blob, buf.size()));
  e->blob_bytes = buf.size();
  CK(cudaMemcpy(e->blob, buf.data(), buf.size(), cudaMemcpyHostToDevice));
  unsigned char* d = (unsigned char*)e->blob;
  Prog p = H.hdr;
  p.row_ptr = (const int*)(d + o_row); p.pcol = (const unsigned short*)(d + o_pcol);
  p.prowi = (const unsigned short*)(d + o_prowi); p.pdiag = d + o_pdiag;
  p.term_ptr = (const int*)(d + o_tp); p.term_slot = (const int*)(d + o_ts);
  p.term_kind = (const signed char*)(d + o_tk);
  p.nl_ptr = (const int*)(d + o_nlp); p.nl_src = (const short*)(d + o_nls);
  p.nl_sgn = (const signed char*)(d + o_nlg);
  p.tx_ptr = (const int*)(d + o_txp); p.tx_src = (const short*)(d + o_txs);
  p.tx_sgn = (const signed char*)(d + o_txg);
  p.lin = (const LinOp*)(d + o_lin); p.src = (const SrcOp*)(d + o_src);
  p.dev = (const DevOp*)(d + o_dev); p.lock = (const int*)(d + o_lock);
  p.guessG = (const double*)(d + o_gG); p.gslot = (const int*)(d + o_gs);
  p.gscale = (const double*)(d + o_gsc); p.gconst = (const double*)(d + o_gc);
  p.nominal = (const double*)(d + o_nom);
  e->d_slot_vary = (int*)(d + o_sv);
  p.slot_vary = e->d_slot_vary;
  e->dprog = p;
  return 0;
}

int set_batch(ahk_engine* e, const ahk_batch* b, cudaStream_t s) {
  if (!b || b->B <= 0) return fail("batch: B must be > 0");
  if (b->n_vary < 0 || (b->n_vary > 0 && (!b->vary_slots || !b->vary)))
    return fail("batch: vary_slots / vary missing");
  std::vector<int> slots(b->vary_slots, b->vary_slots + b->n_vary);
  if (e->slots_valid && slots == e->cur_slots) return 0;
  try { set_vary_slots(e->H, b->n_vary, slots.data()); }
  catch (const std::exception& ex) { return fail(ex.what()); }
  CK(cudaStreamSynchronize(s));   // in-flight kernels may still read the old table
  CK(cudaMemcpyAsync(e->d_slot_vary, e->H.slot_vary.data(), e->H.slot_vary.size() * sizeof(int),
                     cudaMemcpyHostToDevice, s));
  e->cur_slots = slots;
  e->slots_valid = true;
  return 0;
}

int to_scratch(ahk_engine* e, const double* host, int cnt, cudaStream_t s) {
  if ((size_t)cnt > e->scratch_cap) {
    CK(cudaStreamSynchronize(s));
    if (e->d_scratch) cudaFree(e->d_scratch);
    e->scratch_cap = (size_t)cnt * 2 + 64;
    CK(cudaMalloc(&e->d_scratch, e->scratch_cap * sizeof(double)));
  }
  CK(cudaMemcpyAsync(e->d_scratch, host, (size_t)cnt * sizeof(double), cudaMemcpyHostToDevice, s));
  return 0;
}

int pow2floor(long long v) { int g = 1; while (2LL * g <= v) g *= 2; return g; }
int pow2ceil(int v) { int g = 1; while (g < v) g *= 2; return g; }

// lanes per instance: enough to spread the rows / devices of one instance, but
// no more than what still fills the machine with B instances
int pick_group(const ahk_engine* e, long long units, int useful_min) {
  if (e->group > 0) return e->group;
  const Prog& h = e->H.hdr;
  int useful = pow2ceil(std::max(std::max(h.ndev, (h.n + 1) / 2), useful_min));
  long long slots = (long long)e->sm_count * 1024;
  int fill = pow2floor(std::max(1LL, slots / std::max(1LL, units)));
  int g = std::min(useful, fill);
  return std::max(1, std::min(32, g));
}

struct Launch { int G, threads, blocks; size_t smem; };

int plan(const ahk_engine* e, long long units, int stride_of_G(const ahk_engine*, int, void*),
         void* ud, int G, Launch& L) {
  for (int threads = 128; threads >= 32; threads >>= 1) {
    if (threads < G) break;
    int ipb = threads / G;
    size_t smem = (size_t)ipb * stride_of_G(e, G, ud) * sizeof(double);
    if (smem <= e->smem_max) {
      L.G = G; L.threads = threads; L.smem = smem;
      L.blocks = (int)((units + ipb - 1) / ipb);
      return 0;
    }
  }
  return fail("circuit too large for the shared-memory engine (n = " + std::to_string(e->n) + ")");
}

template <typename K, typename F>
int launch_g(int G, F&& f) {
  switch (G) {
    case 1: return f(std::integral_constant<int, 1>());
    case 2: return f(std::integral_constant<int, 2>());
    case 4: return f(std::integral_constant<int, 4>());
    case 8: return f(std::integral_constant<int, 8>());
    case 16: return f(std::integral_constant<int, 16>());
    case 32: return f(std::integral_constant<int, 32>());
  }
  return fail("bad group size");
}

template <typename KernelT, typename Args>
int do_launch(KernelT kern, const Args& args, const Launch& L, cudaStream_t s) {
  if (L.smem > 48 * 1024)
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
  kern<<<L.blocks, L.threads, L.smem, s>>>(args);
  CK(cudaGetLastError());
  g_launches++;
  return 0;
}

KBase make_base(ahk_engine* e, const ahk_batch* b, const Layout& L) {
  KBase k;
  k.p = e->dprog; k.o = e->o; k.L = L; k.vary = b->vary; k.B = b->B;
  return k;
}

struct LayoutReq { int mode, method, usc; };
int stride_small(const ahk_engine* e, int G, void* ud) {
  LayoutReq* r = (LayoutReq*)ud;
  return make_layout(e->H, r->mode, r->method, r->usc, G).stride;
}
int stride_ac(const ahk_engine* e, int G, void*) { return make_ac_layout(e->H, G).stride; }

}  // namespace

extern "C" {

const char* ahk_last_error(void) { return g_err.c_str(); }
int64_t ahk_launch_count(void) { return g_launches.load(); }

int ahk_create(const ahk_circuit* circ, const ahk_options* opts, ahk_engine** out) {
  if (!circ || !opts || !out) return fail("ahk_create: null argument");
  static_assert(sizeof(Opts) == sizeof(ahk_options), "Opts must mirror ahk_options");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail("ahk_create: no CUDA device available (this engine has no CPU fallback)");
  ahk_engine* e = new (std::nothrow) ahk_engine();
  if (!e) return fail("out of memory");
  try { build_program(circ, e->H); }
  catch (const std::exception& ex) { delete e; return fail(ex.what()); }
  std::memcpy(&e->o, opts, sizeof(Opts));
  e->n = e->H.hdr.n;
  cudaGetDevice(&e->device);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, e->device) == cudaSuccess) {
    e->sm_count = prop.multiProcessorCount;
    e->smem_max = prop.sharedMemPerBlockOptin;
  }
  if (upload_program(e) != 0) { delete e; return -1; }
  *out = e;
  return 0;
}

void ahk_destroy(ahk_engine* e) {
  if (!e) return;
  if (e->blob) cudaFree(e->blob);
  if (e->d_scratch) cudaFree(e->d_scratch);
  e->big.release();
  delete e;
}

int ahk_size(const ahk_engine* e) { return e ? e->n : -1; }

int ahk_set_options(ahk_engine* e, const ahk_options* opts) {
  if (!e || !opts) return fail("null argument");
  std::memcpy(&e->o, opts, sizeof(Opts));
  return 0;
}

int ahk_set_group(ahk_engine* e, int lanes) {
  if (!e) return fail("null engine");
  if (lanes != 0 && lanes != 1 && lanes != 2 && lanes != 4 && lanes != 8 && lanes != 16 && lanes != 32)
    return fail("group must be 0,1,2,4,8,16,32");
  e->group = lanes;
  return 0;
}

int ahk_probe_fp64(int blocks, int threads, int iters, double* out, void* stream) {
  if (blocks <= 0 || threads <= 0 || threads > 256 || iters <= 0 || !out)
    return fail("ahk_probe_fp64: bad argument");
  k_probe_fp64<<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, 1.0, out);
  CK(cudaGetLastError());
  return 0;
}

int ahk_op(ahk_engine* e, const ahk_batch* batch, const double* x0, int use_guess, double* x,
           double* resid, int32_t* iters, int32_t* status, void* stream) {
  if (!e || !x || !iters || !status) return fail("ahk_op: null argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (set_batch(e, batch, s)) return -1;
  if (e->n > AHK_SMALL_NMAX) {
    if (e->H.nonlinear) return fail("non-linear circuits with n > 64 are not supported yet");
    double v = 0.0;
    return e->big.run(e, batch, -1, &v, 1, x, iters, status, resid, s, g_err, g_launches);
  }
  LayoutReq rq = {MODE_OP, 0, 0};
  Launch L;
  if (plan(e, batch->B, stride_small, &rq, pick_group(e, batch->B, 1), L)) return -1;
  KOp k;
  k.b = make_base(e, batch, make_layout(e->H, MODE_OP, 0, 0, L.G));
  k.a = OpArgs{x0, use_guess, x, resid, iters, status};
  return launch_g<KOp>(L.G, [&](auto g) { return do_launch(k_op<decltype(g)::value>, k, L, s); });
}

int ahk_dc_sweep(ahk_engine* e, const ahk_batch* batch, int src_elem, const double* sweep_values,
                 int npts, const double* x0, int use_guess, double* x_out, int32_t* iters,
                 int32_t* status, void* stream) {
  if (!e || !sweep_values || npts <= 0 || !x_out || !iters || !status)
    return fail("ahk_dc_sweep: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (set_batch(e, batch, s)) return -1;
  bool ok = false;
  for (auto& so : e->H.src) if (so.elem == src_elem && !so.tf) ok = true;
  if (!ok) return fail("ahk_dc_sweep: src_elem is not a DC V/I source");
  if (e->n > AHK_SMALL_NMAX) {
    if (e->H.nonlinear) return fail("non-linear circuits with n > 64 are not supported yet");
    return e->big.run(e, batch, src_elem, sweep_values, npts, x_out, iters, status, nullptr, s,
                      g_err, g_launches);
  }
  if (to_scratch(e, sweep_values, npts, s)) return -1;
  LayoutReq rq = {MODE_OP, 0, 0};
  Launch L;
  if (plan(e, batch->B, stride_small, &rq, pick_group(e, batch->B, 1), L)) return -1;
  KDc k;
  k.b = make_base(e, batch, make_layout(e->H, MODE_OP, 0, 0, L.G));
  k.a = DcArgs{src_elem, e->d_scratch, npts, x0, use_guess, x_out, iters, status};
  return launch_g<KDc>(L.G, [&](auto g) { return do_launch(k_dc<decltype(g)::value>, k, L, s); });
}

int ahk_tran(ahk_engine* e, const ahk_batch* batch, const double* x0, double tstart, double tstep,
             double tstop, int method, int use_step_control, int64_t max_rows, double* t_out,
             double* x_out, int32_t* rows, int32_t* counters, int32_t* status, void* stream) {
  if (!e || !t_out || !x_out || !rows || !counters || !status || max_rows <= 0)
    return fail("ahk_tran: bad argument");
  if (!(method == AHK_IMPLICIT_EULER || method == AHK_TRAP || (method >= AHK_GEAR1 && method <= AHK_GEAR6)))
    return fail("ahk_tran: unknown method");
  if (tstart > tstop) return fail("tstart > tstop");       // transient.py:153-155
  if (tstep < 0) return fail("tstep < 0");                 // transient.py:156-158
  if (e->n > AHK_SMALL_NMAX) return fail("transient needs n <= 64 in this build");
  cudaStream_t s = (cudaStream_t)stream;
  if (set_batch(e, batch, s)) return -1;
  LayoutReq rq = {MODE_TRAN, method, use_step_control};
  Launch L;
  if (plan(e, batch->B, stride_small, &rq, pick_group(e, batch->B, 4), L)) return -1;
  KTran k;
  k.b = make_base(e, batch, make_layout(e->H, MODE_TRAN, method, use_step_control, L.G));
  k.x0 = x0;
  k.a = TranArgs{tstart, tstep, tstop, method, use_step_control, max_rows, t_out, x_out, rows, counters};
  k.status = status;
  return launch_g<KTran>(L.G, [&](auto g) { return do_launch(k_tran<decltype(g)::value>, k, L, s); });
}

int ahk_ac(ahk_engine* e, const ahk_batch* batch, const double* x_op, const double* omegas,
           int npts, double* x_out, int32_t* status, void* stream) {
  if (!e || !omegas || npts <= 0 || !x_out || !status) return fail("ahk_ac: bad argument");
  if (e->H.nonlinear && !x_op) return fail("ahk_ac: non-linear circuit needs x_op");
  if (e->n > AHK_SMALL_NMAX) return fail("AC needs n <= 64 in this build");
  cudaStream_t s = (cudaStream_t)stream;
  if (set_batch(e, batch, s)) return -1;
  if (to_scratch(e, omegas, npts, s)) return -1;
  // frequency chunks: enough (instance, chunk) units to fill the machine
  int nchunks = 1;
  {
    long long want = (long long)e->sm_count * 64;
    while (nchunks < npts && batch->B * nchunks < want && npts / (nchunks * 2) >= 4) nchunks *= 2;
  }
  long long units = batch->B * nchunks;
  Launch L;
  if (plan(e, units, stride_ac, nullptr, pick_group(e, units, 8), L)) return -1;
  {
    long long nb = (batch->B + 255) / 256;
    k_fill_i32<<<(int)nb, 256, 0, s>>>(status, batch->B, AHK_ST_OK);
    CK(cudaGetLastError());
    g_launches++;
  }
  KAc k;
  k.p = e->dprog; k.o = e->o; k.AL = make_ac_layout(e->H, L.G);
  k.vary = batch->vary; k.B = batch->B; k.nchunks = nchunks;
  k.a = AcArgs{x_op, e->d_scratch, npts, 0, npts, x_out, status};
  return launch_g<KAc>(L.G, [&](auto g) { return do_launch(k_ac<decltype(g)::value>, k, L, s); });
}

}  // extern "C"
