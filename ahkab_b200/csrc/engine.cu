// engine.cu — host side of the C-ABI (include/ahkab_b200.h): program upload, kernel-variant
// selection, launches.  The kernels live in k_dc.cu / k_tran.cu / k_ac.cu (small circuits,
// core.cuh) and biglin.cuh (large linear circuits).  sm_100a only; no CPU fallback.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <sstream>
#include <string>
#include <vector>

#include "biglin3.cuh"
#include "biglin3_gen.hpp"
#include "jit.hpp"
#include "jit_gen.hpp"
#include "kernels.cuh"

namespace ahk { namespace jit {
const Source g_sources[] = {
#include "jit_sources.inc"
};
const int g_nsources = (int)(sizeof(g_sources) / sizeof(g_sources[0]));
} }

using namespace ahk;

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};
std::atomic<long long> g_jit_launches{0};   // launches of circuit-specialised kernels
std::atomic<long long> g_jit_builds{0};     // NVRTC compilations

int fail(const std::string& m) { g_err = m; return -1; }

// appends v (16-byte aligned) to a blob that is uploaded in one piece; returns its offset
template <typename T>
size_t put(std::vector<unsigned char>& buf, const std::vector<T>& v) {
  size_t off = (buf.size() + 15) & ~size_t(15);
  buf.resize(off + v.size() * sizeof(T) + 16);
  if (!v.empty()) std::memcpy(buf.data() + off, v.data(), v.size() * sizeof(T));
  return off;
}

#define CK(call)                                                                     \
  do {                                                                               \
    cudaError_t e_ = (call);                                                         \
    if (e_ != cudaSuccess)                                                           \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e_));               \
  } while (0)

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

// Ragged transient results -> contiguous: block b copies the rows[b] accepted rows of instance
// b (out of max_rows allocated) behind those of the instances before it.  Both copies are
// contiguous runs of doubles (coalesced); HBM bound.
__global__ void __launch_bounds__(256) k_pack_rows(const double* __restrict__ t_out, const double* __restrict__ x_out,
                                                   const int* __restrict__ rows, const long long* __restrict__ offsets,
                                                   long long B, long long max_rows, int n,
                                                   double* __restrict__ t_packed, double* __restrict__ x_packed) {
  for (long long b = blockIdx.x; b < B; b += gridDim.x) {
    const long long r = rows[b], off = offsets[b];
    const double* ts = t_out + b * max_rows;
    const double* xs = x_out + b * max_rows * n;
    double* td = t_packed + off;
    double* xd = x_packed + off * n;
    for (long long i = threadIdx.x; i < r; i += blockDim.x) td[i] = ts[i];
    for (long long i = threadIdx.x; i < r * n; i += blockDim.x) xd[i] = xs[i];
  }
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
  int group = 0;           // 0 = heuristic, else only variants with that many lanes
  int force_vid = -1;      // real-valued kernels: forced variant id (-1 = heuristic)
  int force_ac_vid = -1;   // AC kernels
  int device = 0;
  int sm_count = 148;
  size_t smem_max = 227 * 1024;
  void* blob = nullptr;    // device copy of the program tables
  size_t blob_bytes = 0;
  Prog dprog;              // device-pointer view
  int* d_slot_vary = nullptr;
  DevOp* d_dev = nullptr;  // device copy of the device table (its constant-block offsets follow the batch)
  std::vector<int> cur_slots;
  bool slots_valid = false;
  double* d_scratch = nullptr;   // sweep values / omegas
  double* bigA = nullptr;        // large-n engine: [B][n x RS] matrix workspace (grow-only)
  size_t bigA_cap = 0;           // doubles
  size_t scratch_cap = 0;
  BigLin big;              // large-n linear path (biglin.cuh)
  bool force_general = false;   // tests / tuning: large LINEAR circuits through the general large-n engine too
  // circuit-specialised kernels (jit_gen.hpp): 0 = never, 1 = always (failing to build one
  // is an error), 2 = batches of at least jit_min_batch instances, interpreted kernel if
  // NVRTC is not available
  int jit_mode = 2;
  long long jit_min_batch = 4096;
  int jit_R = 0, jit_G = 0;   // forced (rows per lane, lanes per instance) of the specialised kernels, 0 = heuristic
  int jit_ilp = 0;            // devices a lane evaluates interleaved (0 = heuristic)
  int jit_threads = 0, jit_minb = 0;   // forced block size / resident blocks per SM of the specialised kernels
  int last[8] = {-1, 0, 0, 0, 0, 0, 0, 0};   // ahk_last_launch: vid, NC, R, G, threads, blocks, specialised?, analysis
  // variant / layout / launch shape / kernel of recent small-circuit calls, so that a repeated
  // call (the common case: same batch size again and again) goes straight to the launch
  struct Prepared { int vid; Layout Ly; Launch L; jit::Kernel* jk; };
  std::map<std::vector<long long>, Prepared> prepared;
  long long slots_epoch = 0, opts_epoch = 0;   // bumped when the varied slots / options change
  std::map<std::vector<long long>, jit::Kernel> jit_cache;
  // static-schedule LU of the specialised one-lane / replicated kernels (core.cuh: slu_solve_thread):
  // pivot sequences learned by calibration launches, per (mode, method, step control)
  std::map<std::vector<long long>, std::vector<std::vector<int> > > slu_seq;
  int* d_pivrec = nullptr;     // [32][256] calibration records (core.cuh: slu_record)
  int slu_mode = -1;           // ahk_set_jit_slu: -1 = AHK_SLU / default (transients), 0 off, 1 transients, 2 OP / DC too
  int slu_stats[4] = {0, 0, 0, 0};   // calibration launches, schedules replaced, last: solves off the schedule / solves
  double slu_c1_hint = 0.0;    // DF coefficient of the transient being planned (nominal step), for slu_guess
  // host-buffer entry points (ahk_op_host): three internal streams + two chunk workspaces
  struct HostPipe {
    enum { NS = 8 };   // chunk slots: each has its own kernel stream, events and workspace
    cudaStream_t s_in = nullptr, s_comp[NS] = {}, s_out = nullptr;
    cudaEvent_t ev_start = nullptr, ev_in[NS] = {}, ev_comp[NS] = {}, ev_out[NS] = {}, ev_done = nullptr;
    unsigned char* ws[NS] = {};   // per slot: vary | x | resid | iters | status
    size_t ws_bytes = 0;
    bool ready = false;
    // the whole chunked sequence as ONE CUDA graph (a dozen driver calls per chunk would
    // otherwise make the call CPU bound): captured the second time the same call is seen
    cudaStream_t s_org = nullptr;
    std::vector<long long> last_key, graph_key;
    cudaGraphExec_t graph_exec = nullptr;
    long long graph_launches = 0, graph_jit_launches = 0;
  } hp;
};

void BigLin::release_c3() {
  for (int i = 0; i < 2; i++) {
    jit::Kernel* k = (jit::Kernel*)c3_mod[i];
    if (k) { if (i == 0) jit::unload(*k); delete k; }
    c3_mod[i] = nullptr;
  }
  cudaFree(c3_F); cudaFree(c3_Mv); cudaFree(c3_flagm); cudaFree(c3_scr);
  c3_F = c3_Mv = c3_scr = nullptr; c3_flagm = nullptr;
  c3_F_cap = c3_Mv_cap = c3_flag_cap = c3_scr_cap = 0; c3_state = 0;
}
void BigLin::release3() {
  release_c3();
  delete s3_S; s3_S = nullptr;
  cudaFree(s3_sched); cudaFree(s3_V); cudaFree(s3_Mv); cudaFree(s3_bs); cudaFree(s3_xs); cudaFree(s3_flag);
  s3_sched = nullptr; s3_V = s3_Mv = s3_bs = s3_xs = nullptr; s3_flag = nullptr;
  s3_V_cap = s3_Mv_cap = s3_scr_cap = s3_flag_cap = 0; s3_ok = false; s3_slots_epoch = s3_opts_epoch = -1;
}

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
  if (n >= 32768) { err = "n too large"; return -1; }
  if (npts > sweep_cap) {
    CKB(cudaStreamSynchronize(s));
    cudaFree(d_sweep);
    sweep_cap = npts * 2 + 16;
    CKB(cudaMalloc(&d_sweep, sweep_cap * sizeof(double)));
  }
  CKB(cudaMemcpyAsync(d_sweep, sweep_host, npts * sizeof(double), cudaMemcpyHostToDevice, s));

  // ---- 0. static elimination schedule (biglin3.cuh): one thread per instance ----------------
  // The schedule follows the pivot sequence of the nominal instance; instances whose pivots
  // differ are appended to the overflow list and go through the dynamic kernel below.
  bool staged = false;
  // (AHK_BIGLIN3_OFF / AHK_BIGLIN_DENSE / AHK_BIGLIN_SC: tests force the other paths)
  if (!getenv("AHK_BIGLIN3_OFF") && !getenv("AHK_BIGLIN_DENSE") && !getenv("AHK_BIGLIN_SC")) {
    if (s3_slots_epoch != e->slots_epoch || s3_opts_epoch != e->opts_epoch) {
      Big3Sched S;
      s3_ok = big3_build(e->H, e->o, S);
      s3_slots_epoch = e->slots_epoch; s3_opts_epoch = e->opts_epoch;
      // the schedule as straight-line code (biglin3_gen.hpp), under the engine's JIT policy;
      // built lazily below, when a batch that the policy wants specialised arrives
      CKB(cudaStreamSynchronize(s));
      release_c3();
      delete s3_S; s3_S = nullptr;
      if (s3_ok && big3_codegen_fits(S)) s3_S = new Big3Sched(S); else c3_state = -1;
      if (s3_ok) {
        std::vector<unsigned char> buf;
        size_t off[11];
        off[9] = put(buf, S.rslot); off[10] = put(buf, S.rindex);
        off[0] = put(buf, S.step); off[1] = put(buf, S.prow); off[2] = put(buf, S.cand_slot);
        off[3] = put(buf, S.cand_row); off[4] = put(buf, S.l_slot); off[5] = put(buf, S.l_row);
        off[6] = put(buf, S.u_slot); off[7] = put(buf, S.u_col); off[8] = put(buf, S.upd);
        CKB(cudaStreamSynchronize(s));
        cudaFree(s3_sched); s3_sched = nullptr;
        CKB(cudaMalloc(&s3_sched, buf.size() + 64));
        CKB(cudaMemcpy(s3_sched, buf.data(), buf.size(), cudaMemcpyHostToDevice));
        for (int i = 0; i < 11; i++) s3_ptr[i] = (const unsigned char*)s3_sched + off[i];
        s3_n = S.n; s3_npos = S.npos; s3_nslot = S.nslot; s3_nrec = (int)S.rslot.size();
      }
    }
    // ---- 0a. the schedule compiled for this circuit (NVRTC) ----------------------------------
    if (s3_ok && c3_state >= 0 && (e->jit_mode == 1 || (e->jit_mode == 2 && batch->B >= e->jit_min_batch))) {
      if (c3_state == 0) {
        Big3Gen G;
        std::string jerr;
        bool cached = false;
        jit::Kernel* kf = new jit::Kernel(); jit::Kernel* ks = new jit::Kernel();
        if (s3_S && big3_codegen(e->H, e->o, *s3_S, G) &&
            jit::build_text2(G.src, "ahk_big3c_factor", "ahk_big3c_solve", *kf, *ks, jerr, &cached)) {
          c3_mod[0] = kf; c3_mod[1] = ks; c3_nf = G.nf; c3_npos = s3_S->npos; c3_state = 1;
          if (!cached) g_jit_builds++;
          if (getenv("AHK_DEBUG"))
            fprintf(stderr, "[ahk] compiled elimination schedule %s: %zu statements, %d factor values per matrix\n",
                    cached ? "loaded from the disk cache" : "built", G.statements, G.nf);
        } else {
          delete kf; delete ks; c3_state = -1;
          if (e->jit_mode == 1) { err = "compiled elimination schedule: " + jerr; return -1; }
          if (getenv("AHK_DEBUG")) fprintf(stderr, "[ahk] compiled schedule not built: %s\n", jerr.c_str());
        }
      }
      if (c3_state == 1) {
        const long long B = batch->B;
        // one launch for all sweep points unless the grid would be enormous: measured on C5
        // (gpurun_out/r2i_ptc.txt) 1 / 2 / 3 / 5 / 10 points per launch = 6.9 / 3.6 / 3.0 / 1.7 /
        // 1.15 ms — the solve is a serial chain per thread, more threads in flight hide it
        int ptc = (int)std::max<long long>(1, std::min<long long>(npts, (1ll << 22) / std::max<long long>(1, B)));
        if (const char* f = getenv("AHK_BIGLIN3C_PTC")) ptc = std::max(1, std::min(npts, atoi(f)));   // tuning
        // Solve blocks: the straight-line code is ~300 KB per thread and every warp streams it once
        // (instruction fetch is the first stall, ncu r3c / r3p: no_instruction 10 cycles per issued
        // instruction); a barrier per block of steps keeps the warps of a block in the same code
        // window.  One big block per SM was tried and is not better than small blocks.
        long long sthr = 64;   // (measured on C5: 64 / 128 / 288 threads = 0.542 / 0.557 / 0.569 ms, gpurun_out/r3u_c5.txt)
        if (const char* f = getenv("AHK_BIGLIN3C_THREADS")) sthr = std::max(32, std::min(320, atoi(f) / 32 * 32));   // tuning
        const long long sgrid_max = (B * ptc + sthr - 1) / sthr;
        // scratch of the solve threads: right-hand side and iterate, [2 n][threads of a launch]
        const size_t B32 = (size_t)(B + 31) / 32 * 32;   // (tiled by 32 lanes: biglin3_gen.hpp)
        const size_t needF = (size_t)2 * c3_nf * B32, needM = (size_t)c3_npos * B32, needS = (size_t)2 * n * sgrid_max * sthr;
        if (needF > c3_F_cap || needM > c3_Mv_cap || needS > c3_scr_cap || (size_t)2 * B > c3_flag_cap || B + 1 > ovf_cap) {
          CKB(cudaStreamSynchronize(s));
          if (needF > c3_F_cap) { cudaFree(c3_F); c3_F = nullptr; CKB(cudaMalloc(&c3_F, needF * 8)); c3_F_cap = needF; }
          if (needM > c3_Mv_cap) { cudaFree(c3_Mv); c3_Mv = nullptr; CKB(cudaMalloc(&c3_Mv, needM * 8)); c3_Mv_cap = needM; }
          if (needS > c3_scr_cap) { cudaFree(c3_scr); c3_scr = nullptr; CKB(cudaMalloc(&c3_scr, needS * 8)); c3_scr_cap = needS; }
          if ((size_t)2 * B > c3_flag_cap) { cudaFree(c3_flagm); c3_flagm = nullptr; CKB(cudaMalloc(&c3_flagm, (size_t)2 * B * 4)); c3_flag_cap = 2 * B; }
          if (B + 1 > ovf_cap) {
            cudaFree(d_ovf); d_ovf = nullptr; ovf_cap = B + 1;
            CKB(cudaMalloc(&d_ovf, (size_t)ovf_cap * sizeof(int)));
          }
        }
        CKB(cudaMemsetAsync(d_ovf, 0, sizeof(int), s));
        Big3CK ck;
        ck.vary = batch->vary; ck.B = B; ck.src_elem = src_elem; ck.sweep = d_sweep; ck.npts = npts;
        ck.pt0 = 0; ck.ptc = npts;
        ck.x_out = x_out; ck.iters = iters; ck.status = status; ck.resid = resid;
        ck.F = c3_F; ck.Mv = c3_Mv; ck.flagm = c3_flagm; ck.ovf = d_ovf;
        ck.scr = c3_scr; ck.scr_stride = sgrid_max * sthr;
        std::string lerr;
        const int T = 64;
        if (!jit::launch(*(jit::Kernel*)c3_mod[0], (int)((2 * B + T - 1) / T), T, 0, s, &ck, lerr)) { err = lerr; return -1; }
        launches++; g_jit_launches++;
        for (int pt0 = 0; pt0 < npts; pt0 += ptc) {
          ck.pt0 = pt0; ck.ptc = std::min(ptc, npts - pt0);
          const long long TT = B * ck.ptc;
          if (!jit::launch(*(jit::Kernel*)c3_mod[1], (int)((TT + sthr - 1) / sthr), (int)sthr, 0, s, &ck, lerr)) { err = lerr; return -1; }
          launches++; g_jit_launches++;
        }
        staged = true; last_path = 3;
      }
    }
    if (s3_ok && !staged) {
      const long long B = batch->B;
      // scratch of the solve kernel: [n][B * ptc] doubles twice, at most ~1 GB per buffer
      int ptc = (int)std::max<long long>(1, std::min<long long>(npts, (1ll << 27) / ((long long)n * B)));
      // (the reciprocal table shares the allocation of the matrix values: Mv | G)
      const size_t needV = (size_t)2 * s3_nslot * B, needM = (size_t)(s3_npos + s3_nrec) * B, needS = (size_t)n * B * ptc;
      if (needV > s3_V_cap || needM > s3_Mv_cap || needS > s3_scr_cap || (size_t)B > s3_flag_cap ||
          batch->B + 1 > ovf_cap) {
        CKB(cudaStreamSynchronize(s));
        if (needV > s3_V_cap) { cudaFree(s3_V); s3_V = nullptr; CKB(cudaMalloc(&s3_V, needV * 8)); s3_V_cap = needV; }
        if (needM > s3_Mv_cap) { cudaFree(s3_Mv); s3_Mv = nullptr; CKB(cudaMalloc(&s3_Mv, needM * 8)); s3_Mv_cap = needM; }
        if (needS > s3_scr_cap) {
          cudaFree(s3_bs); cudaFree(s3_xs); s3_bs = s3_xs = nullptr;
          CKB(cudaMalloc(&s3_bs, needS * 8)); CKB(cudaMalloc(&s3_xs, needS * 8)); s3_scr_cap = needS;
        }
        if ((size_t)B > s3_flag_cap) { cudaFree(s3_flag); s3_flag = nullptr; CKB(cudaMalloc(&s3_flag, (size_t)B * 4)); s3_flag_cap = B; }
        if (batch->B + 1 > ovf_cap) {
          cudaFree(d_ovf); d_ovf = nullptr; ovf_cap = batch->B + 1;
          CKB(cudaMalloc(&d_ovf, (size_t)ovf_cap * sizeof(int)));
        }
      }
      CKB(cudaMemsetAsync(d_ovf, 0, sizeof(int), s));
      Big3K k3;
      k3.p = e->dprog; k3.o = e->o; k3.vary = batch->vary; k3.B = B;
      k3.src_elem = src_elem; k3.sweep = d_sweep; k3.npts = npts; k3.pt0 = 0; k3.ptc = ptc;
      k3.b_smem = 0;
      k3.idx32 = (needV < (1ull << 31) && needM < (1ull << 31) && needS < (1ull << 31)) ? 1 : 0;
      k3.x_out = x_out; k3.iters = iters; k3.status = status; k3.resid = resid;
      k3.n = s3_n; k3.npos = s3_npos; k3.nslot = s3_nslot;
      k3.step = (const Big3Step*)s3_ptr[0]; k3.prow = (const int*)s3_ptr[1];
      k3.cand_slot = (const int*)s3_ptr[2]; k3.cand_row = (const int*)s3_ptr[3];
      k3.l_slot = (const int*)s3_ptr[4]; k3.l_row = (const int*)s3_ptr[5];
      k3.u_slot = (const int*)s3_ptr[6]; k3.u_col = (const int*)s3_ptr[7]; k3.upd = (const int*)s3_ptr[8];
      k3.rslot = (const int*)s3_ptr[9]; k3.rindex = (const int*)s3_ptr[10]; k3.nrec = s3_nrec;
      k3.G = s3_Mv + (size_t)s3_npos * B;
      k3.V = s3_V; k3.Mv = s3_Mv; k3.flag = s3_flag; k3.bs = s3_bs; k3.xs = s3_xs; k3.ovf = d_ovf;
      // one thread per instance: small blocks spread the few warps over all SMs
      int fthreads = 32;
      if (const char* f = getenv("AHK_BIGLIN3_FT")) fthreads = std::max(32, std::min(128, atoi(f)));   // tuning
      k_big3_factor<<<(int)((B + fthreads - 1) / fthreads), fthreads, 0, s>>>(k3);
      CKB(cudaGetLastError());
      launches++;
      // right-hand sides: global scratch [n][thread].  Shared memory (AHK_BIGLIN3_SMEM_RHS=1,
      // blocks of 64 threads x 132 KB for n = 257) was measured and is WORSE — C5 8.6 vs 3.6 ms:
      // two warps per SM cannot hide the latency of the factor / schedule loads
      int sthreads = 128; size_t ssmem = 0;
      if (const char* f = getenv("AHK_BIGLIN3_ST")) sthreads = std::max(32, std::min(128, atoi(f)));   // tuning
      k3.b_smem = 0;
      if (getenv("AHK_BIGLIN3_SMEM_RHS")) {
        for (int th = 64; th >= 32; th >>= 1)
          if ((size_t)n * th * 8 <= e->smem_max) { sthreads = th; ssmem = (size_t)n * th * 8; k3.b_smem = 1; break; }
      }
      if (ssmem > 48 * 1024)
        CKB(cudaFuncSetAttribute(k_big3_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssmem));
      for (int pt0 = 0; pt0 < npts; pt0 += ptc) {
        k3.pt0 = pt0; k3.ptc = std::min(ptc, npts - pt0);
        const long long T = B * k3.ptc;
        k_big3_solve<<<(int)((T + sthreads - 1) / sthreads), sthreads, ssmem, s>>>(k3);
        CKB(cudaGetLastError());
        launches++;
      }
      if (getenv("AHK_DEBUG"))
        fprintf(stderr, "[ahk] biglin3 n=%d slots %d, %lld instances, %d point(s) per solve launch\n", n, s3_nslot, B, ptc);
      staged = true; last_path = 2;
    }
  }

  // ---- 1. sparse shared-memory kernel (n <= 1024) ----------------------------------------
  const int W = (n + 31) / 32;
  bool sparse = !staged && W <= 32 && !getenv("AHK_BIGLIN_DENSE");
  if (!staged) last_path = sparse ? 1 : 0;
  int SC = 0, PC = 0;
  Big2Smem S2;
  if (sparse) {
    // slots per row and column / points per solve chunk: as large as still lets 7 warps share an SM
    const size_t budget = (228 * 1024) / 7 - 1024;
    int maxrow = 0;   // largest number of entries in a row or in a column of the pattern
    for (int r = 0; r < n; r++) maxrow = std::max(maxrow, e->H.row_ptr[r + 1] - e->H.row_ptr[r]);
    {
      std::vector<int> colcnt(n, 0);
      for (int q = 0; q < h.npos; q++) maxrow = std::max(maxrow, ++colcnt[e->H.pcol[q]]);
    }
    PC = std::min(npts, 16);
    SC = std::min(12, std::max(6, maxrow + 3));
    while (PC > 1 && (size_t)big2_smem(n, h.npos, SC, PC, W).total > budget) PC--;
    while (SC > 6 && (size_t)big2_smem(n, h.npos, SC, PC, W).total > budget) SC--;
    while (SC < 32 && SC < maxrow + 4 && (size_t)big2_smem(n, h.npos, SC + 1, PC, W).total <= budget) SC++;
    if (const char* f = getenv("AHK_BIGLIN_SC")) SC = std::max(maxrow, atoi(f));   // tests: force overflows
    if (const char* f = getenv("AHK_BIGLIN_PC")) PC = std::max(1, std::min(npts, atoi(f)));   // tuning
    S2 = big2_smem(n, h.npos, SC, PC, W);
    if ((size_t)S2.total > e->smem_max || maxrow > SC) sparse = false;
  }
  if (sparse) {
    int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (228 * 1024) / ((size_t)S2.total + 1024)));
    int want = (int)std::min<long long>(batch->B, (long long)e->sm_count * per_sm);
    const int cap2n = n * (SC + 2);
    if (want > nwarps2 || cap2n != cap2 || batch->B + 1 > ovf_cap) {
      CKB(cudaStreamSynchronize(s));
      cudaFree(s2Lv); cudaFree(s2Uv); cudaFree(s2Li); cudaFree(s2Ui); cudaFree(d_ovf);
      nwarps2 = want; cap2 = cap2n; ovf_cap = batch->B + 1;
      CKB(cudaMalloc(&s2Lv, (size_t)nwarps2 * 2 * cap2 * sizeof(double)));
      CKB(cudaMalloc(&s2Uv, (size_t)nwarps2 * 2 * cap2 * sizeof(double)));
      CKB(cudaMalloc(&s2Li, (size_t)nwarps2 * 2 * cap2 * sizeof(unsigned short)));
      CKB(cudaMalloc(&s2Ui, (size_t)nwarps2 * 2 * cap2 * sizeof(unsigned short)));
      CKB(cudaMalloc(&d_ovf, (size_t)ovf_cap * sizeof(int)));
    }
    CKB(cudaMemsetAsync(d_ovf, 0, sizeof(int), s));
    Big2K k2;
    k2.p = e->dprog; k2.o = e->o; k2.vary = batch->vary; k2.B = batch->B;
    k2.src_elem = src_elem; k2.sweep = d_sweep; k2.npts = npts; k2.x0 = nullptr;
    k2.x_out = x_out; k2.iters = iters; k2.status = status; k2.resid = resid;
    k2.Lv = s2Lv; k2.Uv = s2Uv; k2.Li = s2Li; k2.Ui = s2Ui;
    k2.cap2 = cap2; k2.SC = SC; k2.PC = PC; k2.W = W; k2.ovf = d_ovf;
    if (S2.total > 48 * 1024)
      CKB(cudaFuncSetAttribute(k_biglin2, cudaFuncAttributeMaxDynamicSharedMemorySize, S2.total));
    if (getenv("AHK_DEBUG"))
      fprintf(stderr, "[ahk] biglin2 n=%d SC=%d PC=%d smem %d B/warp, %d warps (%d per SM)\n", n, SC, PC,
              S2.total, want, per_sm);
    k_biglin2<<<want, 32, S2.total, s>>>(k2, S2);
    CKB(cudaGetLastError());
    launches++;
  }

  // ---- 2. dense-workspace kernel: everything (no sparse pass) or the overflow list ------------
  const int RS = (n + 1) | 1;
  BigSmem S = big_smem(n, h.npos);
  if ((size_t)S.total > e->smem_max) { err = "circuit too large for the large-n path"; return -1; }
  int per_sm = (int)std::min<size_t>(16, e->smem_max / (size_t)S.total);
  int want = (int)std::min<long long>(batch->B, (long long)e->sm_count * ((sparse || staged) ? 1 : per_sm));
  const int capn = n * (n - 1) / 2 + n;
  if (want > nwarps || capn != cap) {
    CKB(cudaStreamSynchronize(s));
    cudaFree(wsA); cudaFree(wsLv); cudaFree(wsUv); cudaFree(wsLi); cudaFree(wsUi);
    nwarps = want; cap = capn;
    CKB(cudaMalloc(&wsA, (size_t)nwarps * n * RS * sizeof(double)));
    CKB(cudaMalloc(&wsLv, (size_t)nwarps * 2 * cap * sizeof(double)));
    CKB(cudaMalloc(&wsUv, (size_t)nwarps * 2 * cap * sizeof(double)));
    CKB(cudaMalloc(&wsLi, (size_t)nwarps * 2 * cap * sizeof(unsigned short)));
    CKB(cudaMalloc(&wsUi, (size_t)nwarps * 2 * cap * sizeof(unsigned short)));
  }
  BigK k;
  k.p = e->dprog; k.o = e->o; k.vary = batch->vary; k.B = batch->B;
  k.src_elem = src_elem; k.sweep = d_sweep; k.npts = npts; k.x0 = nullptr;
  k.x_out = x_out; k.iters = iters; k.status = status; k.resid = resid;
  k.wsA = wsA; k.wsLv = wsLv; k.wsUv = wsUv; k.wsLi = wsLi; k.wsUi = wsUi;
  k.cap = cap; k.RS = RS; k.list = (sparse || staged) ? d_ovf : nullptr;
  if (S.total > 48 * 1024)
    CKB(cudaFuncSetAttribute(k_biglin, cudaFuncAttributeMaxDynamicSharedMemorySize, S.total));
  k_biglin<<<want, 32, S.total, s>>>(k, S);
  CKB(cudaGetLastError());
  launches++;
  return 0;
#undef CKB
}

namespace {

void drop_stale(ahk_engine* e);

int upload_program(ahk_engine* e) {
  HostProg& H = e->H;
  std::vector<unsigned char> buf;
  size_t o_pcd = put(buf, H.pcd), o_pst = put(buf, H.pstage), o_gps = put(buf, H.gpslot);
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
  p.pcd = (const unsigned short*)(d + o_pcd);
  p.pstage = (const int*)(d + o_pst); p.gpslot = (const int*)(d + o_gps);
  p.term_ptr = (const int*)(d + o_tp); p.term_slot = (const int*)(d + o_ts);
  p.term_kind = (const signed char*)(d + o_tk);
  p.nl_ptr = (const int*)(d + o_nlp); p.nl_src = (const short*)(d + o_nls);
  p.nl_sgn = (const signed char*)(d + o_nlg);
  p.tx_ptr = (const int*)(d + o_txp); p.tx_src = (const short*)(d + o_txs);
  p.tx_sgn = (const signed char*)(d + o_txg);
  p.lin = (const LinOp*)(d + o_lin); p.src = (const SrcOp*)(d + o_src);
  p.dev = (const DevOp*)(d + o_dev); p.lock = (const int*)(d + o_lock);
  e->d_dev = (DevOp*)(d + o_dev);
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
  if (!e->H.dev.empty())
    CK(cudaMemcpyAsync(e->d_dev, e->H.dev.data(), e->H.dev.size() * sizeof(DevOp), cudaMemcpyHostToDevice, s));
  e->cur_slots = slots;
  e->slots_valid = true;
  e->slots_epoch++;
  drop_stale(e);
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

// ---- kernel-variant selection -------------------------------------------------------------
bool fits(const VarInfo& v, int n) {
  if (v.NC == 0) return n <= (v.G == 32 ? AHK_BIG_NMAX : AHK_SMALL_NMAX);   // warp per instance: also the large-n engine
  return v.NC - 1 >= n && v.R * v.G >= n;
}

int plan(const ahk_engine* e, long long units, int vid, const VarInfo& v, int stride, Launch& L,
         int reg_blocks, long long* resident);

// Smallest column count that fits (least padding).  Among those: the FEWEST lanes per
// instance (no idle lanes, no cross-lane traffic) whose resident instances per SM
// (bounded by the shared-memory slice and the register cap) hold the SM's share of the
// batch in one wave while still giving the SM >= 128 busy lanes; failing that, the
// variant with the most busy lanes per SM (small batches, or slices too big for one
// wave).  stride_of(v) = doubles per instance of the layout.
template <typename StrideOf>
int pick_variant(const ahk_engine* e, const VarInfo* tab, int ntab, int forced, long long units,
                 int reg_blocks_of(int), StrideOf stride_of, int* vid, bool exact_nc = false) {
  const int n = e->n;
  if (forced >= 0) {
    if (forced >= ntab || !fits(tab[forced], n)) return fail("forced kernel variant does not fit this circuit");
    if (tab[forced].NC == 0 && n > AHK_SMALL_NMAX) return fail("circuit too large");
    *vid = forced;
    return 0;
  }
  // exact_nc: the choice is made for a specialised kernel, which is compiled with exactly
  // n + 1 columns whatever the table says — every fitting (R, G) shape competes
  auto eff = [&](int i) { VarInfo q = tab[i]; if (exact_nc && q.NC > 0) q.NC = n + 1; return q; };
  const long long per_sm = (units + e->sm_count - 1) / e->sm_count;   // instances one SM must hold
  int minNC = 1 << 30;
  for (int i = 2; i < ntab; i++)
    if (fits(tab[i], n) && (e->group <= 0 || tab[i].G == e->group)) minNC = std::min(minNC, eff(i).NC);
  int best = -1, best_one_wave = -1;
  long long best_lanes = -1;
  for (int i = 2; i < ntab; i++) {
    const VarInfo q = eff(i);
    if (!fits(tab[i], n) || q.NC != minNC) continue;
    if (e->group > 0 && q.G != e->group) continue;
    long long res = 0;
    Launch tmp;
    if (plan(e, units, i, q, stride_of(q), tmp, reg_blocks_of(q.NC), &res))
      continue;   // does not fit in shared memory
    const long long lanes = std::min(res, per_sm) * q.G;
    // (specialised kernels, one lane per instance: measured best at EVERY batch size on C2 —
    // 14.3 us at 2 K-16 K instances where 4 / 16 lanes per instance take 35 / 31 us,
    // profiles/r2s_c2_shapes.txt: the straight-line one-lane LU has the shortest serial chain,
    // and a small batch is bound by that chain, not by idle lanes)
    const bool one_lane_jit = exact_nc && q.G == 1 && res >= per_sm;
    if (res >= per_sm && (per_sm * q.G >= 128 || one_lane_jit) &&
        (best_one_wave < 0 || q.G < tab[best_one_wave].G))
      best_one_wave = i;
    if (best < 0 || lanes > best_lanes || (lanes == best_lanes && q.G < tab[best].G)) {
      best = i; best_lanes = lanes;
    }
  }
  if (best_one_wave >= 0) best = best_one_wave;
  if (best < 0) {   // generic fallback
    if (n > AHK_SMALL_NMAX) return fail("circuit too large for the shared-memory engine");
    if (e->group > 0 && e->group != 8 && e->group != 32)
      return fail("no kernel variant with that many lanes fits this circuit");
    best = e->group == 8 ? 0 : (e->group == 32 ? 1 : (units >= 32768 ? 0 : 1));
  }
  *vid = best;
  return 0;
}

// Block size: the per-instance shared-memory slice bounds how many instances are
// resident per SM (228 KB per SM, 1 KB reserved per block); smaller blocks waste
// less of it.  Registers (<= 128 per thread for the common variants) allow 512
// threads per SM.  Pick the block size with the most resident instances.
int plan(const ahk_engine* e, long long units, int vid, const VarInfo& v, int stride, Launch& L,
         int reg_blocks, long long* resident) {
  const size_t sm_total = 228 * 1024;
  long long best_res = -1;
  for (int threads = 128; threads >= 32; threads >>= 1) {
    if (threads < v.G) break;
    const int ipb = threads / v.G;
    const size_t smem = (size_t)ipb * stride * sizeof(double);
    if (smem > e->smem_max) continue;
    long long blocks_sm = (long long)(sm_total / (smem + 1024));
    if (blocks_sm > 32) blocks_sm = 32;
    long long thr = blocks_sm * threads;
    if (thr > 128 * reg_blocks) thr = 128 * reg_blocks;
    const long long res = thr / v.G;
    if (res > best_res) {
      best_res = res;
      L.vid = vid; L.G = v.G; L.threads = threads; L.smem = smem;
      L.blocks = (int)((units + ipb - 1) / ipb);
    }
  }
  if (best_res < 0)
    return fail("circuit too large for the shared-memory engine (n = " + std::to_string(e->n) + ")");
  if (resident) { *resident = best_res; return 0; }
  if (getenv("AHK_DEBUG"))
    fprintf(stderr, "[ahk] n=%d variant %d (NC=%d R=%d G=%d) stride %d doubles, %d thr/block, %d blocks, "
            "%zu B smem/block, %lld instances resident/SM\n", e->n, vid, v.NC, v.R, v.G, stride, L.threads,
            L.blocks, L.smem, best_res);
  return 0;
}

int launched(int rc) {
  if (rc != 0) return fail(std::string("kernel launch: ") + cudaGetErrorString((cudaError_t)rc));
  g_launches++;
  return 0;
}

KBase make_base(ahk_engine* e, const ahk_batch* b, const Layout& L) {
  KBase k;
  k.p = e->dprog; k.o = e->o; k.L = L; k.vary = b->vary; k.B = b->B;
  k.Ag = L.big ? e->bigA : nullptr;
  k.Voff = L.big ? (long long)e->n * L.RS : 0;
  k.Astride = L.big ? k.Voff + L.vlen + (long long)e->n * 8 : 0;   // matrix | value block | panel multipliers
  return k;
}

// Devices one lane evaluates interleaved in a specialised kernel (ekv_eval_v): as many as the
// lane owns, at most 3 (register budget) — the interleaved chains are what keeps the FP64 pipe
// busy at one or two warps per scheduler.  AHK_JIT_ILP / ahk_set_jit_tuning force it.
int jit_ilp_for(const ahk_engine* e, const VarInfo& v, int mode) {
  if (mode == JIT_MODE_AC) return 1;
  int ilp = e->jit_ilp;
  static const int env = getenv("AHK_JIT_ILP") ? atoi(getenv("AHK_JIT_ILP")) : 0;   // tuning
  if (env > 0) ilp = env;
  const int per_lane = ((int)e->H.dev.size() + v.G - 1) / std::max(1, v.G);
  if (ilp <= 0) ilp = std::min(3, per_lane);
  return std::max(1, std::min(std::min(ilp, 4), std::max(1, per_lane)));
}

// AHK_SLU = 0: pivoting LU everywhere; tran (default): static schedule in the transient kernels;
// all: in the OP / DC kernels too (their homotopy stages tend to change the pivot sequence).
std::vector<std::string> split_semicolon(const char* f) {
  std::vector<std::string> out; std::string cur;
  for (const char* q = f; *q; q++) { if (*q == ';') { out.push_back(cur); cur.clear(); } else cur += *q; }
  out.push_back(cur);
  return out;
}

bool slu_wanted(const ahk_engine* e, int mode) {
  int m = e ? e->slu_mode : -1;
  if (m < 0) {
    static const char* env = getenv("AHK_SLU");
    m = !env ? 1 : (env[0] == '0' ? 0 : (env[0] == 'a' ? 2 : 1));
  }
  return m == 2 || (m == 1 && (mode == MODE_TRAN || mode == JIT_MODE_AC));
}

// ... and the calibration has not given up on this (analysis, method, step control)
bool slu_active(const ahk_engine* e, int mode, int method, int usc) {
  if (!slu_wanted(e, mode)) return false;
  auto f = e->slu_seq.find(std::vector<long long>{mode, method, usc});
  return f == e->slu_seq.end() || !f->second.empty();
}

// The specialised kernel for (analysis, variant, launch shape, options, varied slots) —
// compiled on first use, cached in the engine.  *out = nullptr: use the interpreted kernel.
int get_jit(ahk_engine* e, long long B, int mode, int vid, const VarInfo& v, const Layout& Ly,
            const Launch& L, int reg_blocks, int method, int usc, jit::Kernel** out,
            const JitAc* ac = nullptr) {
  *out = nullptr;
  if (e->jit_mode == 0 || v.NC == 0) return 0;   // the generic shared-memory fallback is not specialised
  if (e->jit_mode == 2 && B < e->jit_min_batch) return 0;
  const bool must = e->jit_mode == 1;
  const int ilp = jit_ilp_for(e, v, mode);
  // static elimination schedule: one lane per instance or replicated rows (jit_gen.hpp: slu_codegen)
  std::vector<std::vector<int> > seqs;
  const bool all_rows = v.R >= e->n;
  if (slu_active(e, mode, method, usc) && v.NC > 0 && all_rows && e->n <= 31) {
    auto f = e->slu_seq.find(std::vector<long long>{mode, method, usc});
    if (f != e->slu_seq.end()) seqs = f->second;
    else seqs.push_back(slu_guess(e->H, e->o.gmin, mode == MODE_TRAN ? e->slu_c1_hint : 0.0));
    if (slu_codegen(e->H, seqs, mode == JIT_MODE_AC).empty()) seqs.clear();
  }
  std::vector<long long> key = {mode, vid, L.threads, reg_blocks, method, usc, ilp};
  for (auto& q : seqs) { key.push_back(-1); for (int t : q) key.push_back(t); }
  key.push_back(e->slots_epoch); key.push_back(e->opts_epoch);
  auto it = e->jit_cache.find(key);
  if (it == e->jit_cache.end()) {
    jit::Kernel k;
    // reg_blocks > 0: resident 128-thread blocks the registers are capped for; < 0: -min_blocks of this block size
    const int minb = reg_blocks < 0 ? -reg_blocks : std::max(1, std::min(32, reg_blocks * 128 / L.threads));
    JitCfg cfg = {v.NC, v.R, v.G, L.threads, minb, mode, method, usc, ilp};
    cfg.slu = seqs; k.slu = !seqs.empty();
    std::string err;
    const std::string spec = jit_spec_source(e->H, e->o, Ly, cfg, ac);
    bool cached = false;
    const char* entry = mode == JIT_MODE_AC ? "ahk_jit_ac" : (mode == MODE_TRAN ? "ahk_jit_tran" : "ahk_jit_dc");
    if (!jit::build(spec, entry, L.smem, k, err, &cached)) {
      k.failed = true; k.why = err;
      if (getenv("AHK_DEBUG")) fprintf(stderr, "[ahk] specialised kernel not built: %s\n", err.c_str());
    } else {
      if (!cached) g_jit_builds++;
      if (getenv("AHK_DEBUG"))
        fprintf(stderr, "[ahk] specialised kernel %s (mode %d, variant %d)\n",
                cached ? "loaded from the disk cache" : "built", mode, vid);
    }
    it = e->jit_cache.emplace(key, k).first;
  }
  if (it->second.failed) {
    if (must) return fail("specialised kernel: " + it->second.why);
    return 0;
  }
  *out = &it->second;
  return 0;
}

bool wants_jit(const ahk_engine* e, long long B) {
  return e->jit_mode == 1 || (e->jit_mode == 2 && B >= e->jit_min_batch);
}

// Variant, layout and launch shape of a small-circuit OP/DC (mode MODE_OP) or transient
// call, and its specialised kernel if the policy asks for one (*jk = nullptr: launch the
// interpreted kernel).  A specialised kernel has its own, smaller layout; if it cannot be
// built under the automatic policy the call is planned again for the interpreted kernel.
int prepare_small_uncached(ahk_engine* e, long long B, int mode, int method, int usc, int* vid_out, Layout* Ly,
                           Launch* L, jit::Kernel** jk);

int prepare_small(ahk_engine* e, long long B, int mode, int method, int usc, int* vid_out, Layout* Ly,
                  Launch* L, jit::Kernel** jk) {
  const std::vector<long long> key = {B, mode, method, usc, e->slots_epoch, e->opts_epoch, e->jit_mode,
                                      e->jit_min_batch, e->force_vid, e->group, e->jit_R, e->jit_G, e->jit_ilp, e->jit_threads, e->jit_minb, e->slu_mode};
  auto it = e->prepared.find(key);
  if (it != e->prepared.end()) {
    *vid_out = it->second.vid; *Ly = it->second.Ly; *L = it->second.L; *jk = it->second.jk;
    return 0;
  }
  if (prepare_small_uncached(e, B, mode, method, usc, vid_out, Ly, L, jk)) return -1;
  if (e->prepared.size() > 64) e->prepared.clear();
  e->prepared[key] = ahk_engine::Prepared{*vid_out, *Ly, *L, *jk};
  return 0;
}

int prepare_small_uncached(ahk_engine* e, long long B, int mode, int method, int usc, int* vid_out, Layout* Ly,
                           Launch* L, jit::Kernel** jk) {
  *jk = nullptr;
  if (e->n > AHK_SMALL_NMAX) {
    // General large-n engine: the generic variant with one warp per instance, vectors in shared
    // memory, the dense matrix in the global workspace (core.cuh: lu_solve_glob).  Interpreted
    // kernels only: the specialised ones hold the matrix in registers.
    if (e->n > AHK_BIG_NMAX) return fail("circuit too large: n = " + std::to_string(e->n) + " > " + std::to_string(AHK_BIG_NMAX));
    if (e->jit_mode == 1) return fail("specialised kernels need n <= 64");
    const int vid = 1;
    const VarInfo v = real_variants()[vid];
    *vid_out = vid;
    *Ly = make_layout(e->H, mode, method, usc, v.G, 0, 1, false, true);
    if (plan(e, B, vid, v, Ly->stride, *L, real_min_blocks(0), nullptr)) return -1;
    return 0;
  }
  static const bool full_layout = getenv("AHK_JIT_FULL_LAYOUT") != nullptr;   // tuning experiments
  for (int jit = wants_jit(e, B) ? 1 : 0; jit >= 0; jit--) {
    int vid;
    const bool slim = jit != 0 && !full_layout;
    VarInfo v;
    bool flat_auto = false;
    if (jit && e->jit_G > 0) {
      // a specialised kernel is not bound to the compiled variant table: any (R, G) shape
      vid = 1000 + e->jit_R * 64 + e->jit_G;
      v = VarInfo{e->n + 1, e->jit_R, e->jit_G};
    } else if (jit && mode == MODE_TRAN && e->force_vid < 0 && e->group <= 0 && e->n + 1 <= 32) {
      // Specialised TRANSIENT kernel (flat warp-uniform loop, core.cuh: tran_analysis), shape
      // measured on C4 / C3 (profiles/r2*_sweep_*.txt):
      //  * small matrices (n <= 6) behind devices: REPLICATED — the lanes of a group split the
      //    device evaluations, at most three per lane (evaluated interleaved), and every lane
      //    solves the whole system redundantly on compile-time indices;
      //  * larger matrices: rows distributed, about five per lane.
      const int nd = (int)e->H.dev.size();
      int G = 1, R;
      //  * small batches leave the machine empty and every instance is a serial chain: more
      //    lanes per instance shorten it, as long as the batch still runs in one wave (C4 at
      //    2048 / 4096 instances: 8 lanes 37 / 43 ms, 2 lanes 65 / 66 ms; C3 at 2048 / 4096: 4
      //    lanes 4.7 / 4.9 ms, 2 lanes 6.3 / 6.7 ms — profiles/r2s_c4_small.txt, r2s_c3_small.txt;
      //    at the full 16 384 the two-lane shapes win: 68 vs 87 ms, 6.5 vs 7.1 ms)
      if (e->n <= 6) {
        while (G < 8 && (nd + G - 1) / G > 3) G *= 2;
        // (C4: 8 lanes — one device per lane — win up to 8 K instances: 55.6 vs 63.0 ms with 4
        // lanes and 69 ms with 2; from 16 K on the batch fills the machine and 2 lanes win)
        int gdev = 1;                       // lanes that would each own a device
        while (gdev < 8 && gdev < nd) gdev *= 2;
        if (gdev > G && B * gdev / 32 <= 2368) G = gdev;
        R = e->n;
      } else if (slu_active(e, MODE_TRAN, method, usc) && e->n <= 12) {
        //  * static-schedule LU (core.cuh: slu_solve_thread): the solve is a few dozen multiply-adds
        //    on literal indices, cheaper done redundantly by every lane than distributed — two
        //    replicated lanes per instance split the vector work of the step (C3: 2.67 ms at
        //    16 384 instances against 5.1 with one lane, 4.2 with four and 6.5 for the distributed
        //    Gauss-Jordan; 2.2 ms from 2 K to 8 K instances, profiles/r3f_slu_c3.txt)
        G = 2;
        while (G < 8 && (nd + G - 1) / G > 3) G *= 2;
        if (B * G / 32 > 2368) G = std::max(1, G / 2);
        R = e->n;
      } else {
        while (G < 32 && (e->n + G - 1) / G > 5) G *= 2;
        // (C3 at 8 K instances: 4 lanes 5.3 ms, 2 lanes 6.8 ms; at 4 K: 4 lanes 4.9, 8 lanes 5.5)
        if (B * (2 * G) / 32 <= 1184) G *= 2;
        while (G < 16 && B * (2 * G) / 32 <= 600) G *= 2;
        R = (e->n + G - 1) / G;
      }
      vid = 1000 + R * 64 + G;
      v = VarInfo{e->n + 1, R, G};
      flat_auto = true;
    } else {
      if (pick_variant(e, real_variants(), AHK_N_REAL_VARIANTS, e->force_vid, B, real_min_blocks,
                       [&](const VarInfo& q) { return make_layout(e->H, mode, method, usc, q.G, q.NC, q.R, slim).stride; },
                       &vid, jit != 0))
        return -1;
      v = real_variants()[vid];
    }
    // a specialised kernel is compiled for this circuit: exactly n + 1 matrix columns, no
    // padding up to the column count of the variant table
    if (jit && v.NC > 0) v.NC = e->n + 1;
    *vid_out = vid;
    *Ly = make_layout(e->H, mode, method, usc, v.G, v.NC, v.R, slim && v.NC > 0);
    static const int minb_env = getenv("AHK_JIT_MINB") ? atoi(getenv("AHK_JIT_MINB")) : 0;   // tuning
    int rb = (jit && minb_env > 0) ? minb_env : real_min_blocks(v.NC);
    if (flat_auto && e->jit_threads <= 0 && e->jit_minb <= 0) {
      // ONE WAVE: an instance lives for the whole analysis, so every SM must hold its share of
      // the batch at once — blocks of two warps, and the register cap that lets
      // ceil(share / instances per block) of them be resident (255 registers when the batch
      // leaves the SM mostly empty: the interleaved device evaluations use them)
      const long long per_sm = (B + e->sm_count - 1) / e->sm_count;
      int threads = 64;
      if (threads < v.G) threads = v.G;
      const int ipb = threads / v.G;
      int minb = (int)((per_sm + ipb - 1) / ipb);
      const size_t smem = (size_t)ipb * Ly->stride * sizeof(double);
      const int smem_blocks = (int)((size_t)(228 * 1024) / (smem + 1024));
      if (smem <= e->smem_max && smem_blocks >= 1) {
        minb = std::max(1, std::min(minb, std::min(smem_blocks, 2048 / threads)));
        if (65536 / (minb * threads) < 64) minb = std::max(1, 65536 / (64 * threads));   // never below 64 registers
        L->vid = vid; L->G = v.G; L->threads = threads; L->smem = smem;
        L->blocks = (int)((B + ipb - 1) / ipb);
        if (getenv("AHK_DEBUG"))
          fprintf(stderr, "[ahk] flat transient kernel: n=%d R=%d G=%d, %d thr/block, %d blocks, %zu B smem/block, "
                  "%d resident blocks/SM\n", e->n, v.R, v.G, threads, L->blocks, smem, minb);
        if (get_jit(e, B, mode, vid, v, *Ly, *L, -minb, method, usc, jk)) return -1;
        if (*jk) return 0;
      }
      continue;
    }
    if (jit && (e->jit_threads > 0 || e->jit_minb > 0)) {
      // forced launch shape (ahk_set_jit_tuning): `threads` per block, registers capped for
      // `min_blocks` resident blocks; rb carries min_blocks * threads / 128 to get_jit
      const int threads = e->jit_threads > 0 ? e->jit_threads : 128;
      if (threads < v.G) return fail("ahk_set_jit_tuning: block smaller than one instance group");
      const int ipb = threads / v.G;
      L->vid = vid; L->G = v.G; L->threads = threads;
      L->smem = (size_t)ipb * Ly->stride * sizeof(double);
      L->blocks = (int)((B + ipb - 1) / ipb);
      if (L->smem > e->smem_max) return fail("forced block size needs too much shared memory");
      const int minb = e->jit_minb > 0 ? e->jit_minb : std::max(1, real_min_blocks(v.NC) * 128 / threads);
      if (get_jit(e, B, mode, vid, v, *Ly, *L, -minb, method, usc, jk)) return -1;
      if (*jk) return 0;
      continue;
    }
    if (plan(e, B, vid, v, Ly->stride, *L, rb, nullptr)) return -1;
    if (!jit || v.NC == 0) return 0;
    if (get_jit(e, B, mode, vid, v, *Ly, *L, rb, method, usc, jk)) return -1;
    if (*jk) return 0;
  }
  return 0;
}

void note_launch(ahk_engine* e, int vid, const Launch& L, bool jit, int analysis, bool ac = false) {
  VarInfo v{0, 0, L.G};
  if (vid >= 1000) v = VarInfo{e->n + 1, (vid - 1000) / 64, (vid - 1000) % 64};
  else if (vid >= 0) { v = (ac ? ac_variants() : real_variants())[vid]; if (jit && v.NC > 0) v.NC = e->n + 1; }
  e->last[0] = vid; e->last[1] = v.NC; e->last[2] = v.R; e->last[3] = v.G; e->last[4] = L.threads;
  e->last[5] = L.blocks; e->last[6] = jit ? 1 : 0; e->last[7] = analysis;
}

int jit_launched(bool ok, const std::string& err) {
  if (!ok) return fail(err);
  g_launches++;
  g_jit_launches++;
  return 0;
}

// workspace of the large-n engine: one n x RS matrix per instance (L2 / HBM resident)
int ensure_big_workspace(ahk_engine* e, long long B, const Layout& Ly, cudaStream_t s) {
  if (!Ly.big) return 0;
  const size_t need = (size_t)B * ((size_t)e->n * Ly.RS + Ly.vlen + (size_t)e->n * 8);
  if (need <= e->bigA_cap) return 0;
  if (need * sizeof(double) > ((size_t)96 << 30))
    return fail("large-n workspace of " + std::to_string(need * sizeof(double) >> 20) + " MiB: split the batch");
  CK(cudaStreamSynchronize(s));
  if (e->bigA) { cudaFree(e->bigA); e->bigA = nullptr; e->bigA_cap = 0; }
  CK(cudaMalloc(&e->bigA, need * sizeof(double)));
  e->bigA_cap = need;
  return 0;
}

// Calibration of a static-schedule kernel (first launch of a freshly built one): ONE block of
// the real call runs with the record buffer on, its first 32 instances count their solves and those whose
// pivots left the schedule (they went through the pivoting LU: results are right either way).
// If most of them did, the sequence it recorded becomes the schedule and the kernel is built
// again — at most twice.  `launch1(jk, L, pivrec)` launches one block; `replan()` plans the
// call again (new kernel).  Calibration launches are counted in slu_stats, not in the launch
// counters of the call.  Returns 0, or -1 on a CUDA error.
template <typename Launch1, typename Replan>
int slu_calibrate(ahk_engine* e, int mode, int method, int usc, jit::Kernel*& jk, cudaStream_t s,
                  Launch1 launch1, Replan replan) {
  for (int round = 0; round < 3 && jk && jk->slu && !jk->calibrated; round++) {
    if (!e->d_pivrec) CK(cudaMalloc(&e->d_pivrec, 32 * 256 * sizeof(int)));
    CK(cudaMemsetAsync(e->d_pivrec, 0, 32 * 256 * sizeof(int), s));
    if (launch1(jk, e->d_pivrec)) return -1;
    std::vector<int> all(32 * 256);
    CK(cudaMemcpyAsync(all.data(), e->d_pivrec, all.size() * sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    // 32 records (transients / OP: the first 32 instances; AC: the frequency chunks of instance 0) -> merged
    int rec[256] = {0};
    for (int l = 0; l < 32; l++) {
      const int* r = all.data() + 256 * l;
      rec[0] += r[0]; rec[1] += r[1];
      for (int i = 0; i < r[2] && i < 4; i++) {
        int at = -1;
        for (int j = 0; j < rec[2]; j++) if (std::equal(r + 16 + 32 * i, r + 16 + 32 * i + e->n, rec + 16 + 32 * j)) at = j;
        if (at < 0 && rec[2] < 4) { at = rec[2]++; std::copy(r + 16 + 32 * i, r + 16 + 32 * i + e->n, rec + 16 + 32 * at); }
        if (at >= 0) rec[8 + at] += r[8 + i];
      }
    }
    jk->calibrated = true;
    const int off = rec[0], solves = rec[1];
    e->slu_stats[0]++; e->slu_stats[2] = off; e->slu_stats[3] = solves;
    if (getenv("AHK_DEBUG"))
      fprintf(stderr, "[ahk] static-schedule calibration: %d of %d solves of the calibration sample left the schedule (%d other sequences)\n", off, solves, rec[2]);
    // more than 2 % off the schedule: the sequences they took (each >= 2 % of the solves) join
    // it — or replace it, if nothing stayed on it — and the kernel is built again.  A circuit
    // whose pivoting will not settle (a quarter of the solves still off after the last round, or
    // nothing left to add) loses the schedule: its kernels are built with the pivoting LU.
    if (off * 50 <= solves) break;
    std::vector<std::vector<int> >& set = e->slu_seq[std::vector<long long>{mode, method, usc}];
    bool added = false;
    if (round < 2) {
      if (off == solves) set.clear();
      else if (set.empty()) set.push_back(slu_guess(e->H, e->o.gmin, mode == MODE_TRAN ? e->slu_c1_hint : 0.0));
      for (int i = 0; i < rec[2] && i < 4; i++) {
        if (rec[8 + i] * 50 <= solves || set.size() >= 4) continue;
        set.push_back(std::vector<int>(rec + 16 + 32 * i, rec + 16 + 32 * i + e->n));
        added = true;
      }
    }
    if (!added) {
      if (off * 4 <= solves) break;            // good enough: the rest goes through the fallback
      set.clear();                             // (present but empty: get_jit builds without a schedule)
    }
    e->prepared.clear();
    e->slu_stats[1]++;
    if (replan()) return -1;
  }
  return 0;
}

int run_dc_small(ahk_engine* e, const ahk_batch* batch, int src_elem, const double* d_sweep, int npts,
                 const double* x0, int use_guess, double* x, double* resid, int32_t* iters,
                 int32_t* status, cudaStream_t s) {
  int vid; Layout Ly; Launch L;
  jit::Kernel* jk = nullptr;
  if (prepare_small(e, batch->B, MODE_OP, 0, 0, &vid, &Ly, &L, &jk)) return -1;
  note_launch(e, vid, L, jk != nullptr, 0);
  if (jk) {
    KDcJ kj = {batch->vary, batch->B, DcArgs{src_elem, d_sweep, npts, x0, use_guess, x, resid, iters, status}, nullptr};
    std::string err;
    if (jk->slu && !jk->calibrated) {
      if (slu_calibrate(e, MODE_OP, 0, 0, jk, s,
            [&](jit::Kernel* q, int* rec) { kj.pivrec = rec; const bool ok = jit::launch(*q, 1, L.threads, L.smem, s, &kj, err); kj.pivrec = nullptr; return ok ? 0 : fail(err); },
            [&]() { return prepare_small(e, batch->B, MODE_OP, 0, 0, &vid, &Ly, &L, &jk); })) return -1;
      if (!jk) return fail("static-schedule calibration lost the specialised kernel");
    }
    return jit_launched(jit::launch(*jk, L.blocks, L.threads, L.smem, s, &kj, err), err);
  }
  if (ensure_big_workspace(e, batch->B, Ly, s)) return -1;
  KDc k;
  k.b = make_base(e, batch, Ly);
  k.a = DcArgs{src_elem, d_sweep, npts, x0, use_guess, x, resid, iters, status};
  return launched(launch_dc(k, L, s));
}

// The varied slots or the options changed: plans and specialised kernels compiled for an older
// epoch can never be used again — unload their modules (device memory) and forget them.
// key layout of jit_cache: {..., slots_epoch, opts_epoch} are the last two entries.
void drop_stale(ahk_engine* e) {
  e->prepared.clear();
  if (e->hp.graph_exec) {      // the captured ahk_op_host sequence launches kernels of the old modules
    cudaDeviceSynchronize();
    cudaGraphExecDestroy(e->hp.graph_exec);
    e->hp.graph_exec = nullptr;
    e->hp.graph_key.clear(); e->hp.last_key.clear();
  }
  for (auto it = e->jit_cache.begin(); it != e->jit_cache.end();) {
    const std::vector<long long>& k = it->first;
    const bool stale = k.size() < 2 || k[k.size() - 2] != e->slots_epoch || k[k.size() - 1] != e->opts_epoch;
    if (stale) {
      cudaDeviceSynchronize();   // a launch of the old module may still be in flight
      jit::unload(it->second);
      it = e->jit_cache.erase(it);
    } else ++it;
  }
}

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
  if (e->d_pivrec) cudaFree(e->d_pivrec);
  if (e->bigA) cudaFree(e->bigA);
  for (auto& kv : e->jit_cache) jit::unload(kv.second);
  if (e->hp.ready) {
    if (e->hp.graph_exec) cudaGraphExecDestroy(e->hp.graph_exec);
    cudaStreamDestroy(e->hp.s_org);
    cudaStreamDestroy(e->hp.s_in); cudaStreamDestroy(e->hp.s_out);
    cudaEventDestroy(e->hp.ev_start); cudaEventDestroy(e->hp.ev_done);
    for (int i = 0; i < ahk_engine::HostPipe::NS; i++) {
      cudaStreamDestroy(e->hp.s_comp[i]);
      cudaEventDestroy(e->hp.ev_in[i]); cudaEventDestroy(e->hp.ev_comp[i]); cudaEventDestroy(e->hp.ev_out[i]);
      if (e->hp.ws[i]) cudaFree(e->hp.ws[i]);
    }
  }
  e->big.release();
  delete e;
}

int ahk_size(const ahk_engine* e) { return e ? e->n : -1; }

int ahk_set_options(ahk_engine* e, const ahk_options* opts) {
  if (!e || !opts) return fail("null argument");
  // unchanged options keep every cached plan, specialised kernel, captured graph and
  // elimination schedule valid: callers (analyses.get_engine) set them before every run
  if (std::memcmp(&e->o, opts, sizeof(Opts)) == 0) return 0;
  std::memcpy(&e->o, opts, sizeof(Opts));
  e->opts_epoch++;
  drop_stale(e);
  return 0;
}

int ahk_set_jit(ahk_engine* e, int mode, int64_t min_batch) {
  if (!e) return fail("null engine");
  if (mode < 0 || mode > 2) return fail("ahk_set_jit: mode must be 0 (off), 1 (on) or 2 (auto)");
  e->jit_mode = mode;
  if (min_batch > 0) e->jit_min_batch = min_batch;
  return 0;
}

int ahk_set_jit_shape(ahk_engine* e, int rows_per_lane, int lanes) {
  if (!e) return fail("null engine");
  if (rows_per_lane == 0 && lanes == 0) { e->jit_R = e->jit_G = 0; return 0; }
  if (lanes != 1 && lanes != 2 && lanes != 4 && lanes != 8 && lanes != 16 && lanes != 32)
    return fail("ahk_set_jit_shape: lanes must be 1,2,4,8,16,32");
  if (rows_per_lane < 1 || rows_per_lane * lanes < e->n || e->n + 1 > 32 || rows_per_lane > 16)
    return fail("ahk_set_jit_shape: rows_per_lane * lanes must cover the n <= 31 unknowns");
  e->jit_R = rows_per_lane; e->jit_G = lanes;
  return 0;
}

int ahk_set_jit_tuning(ahk_engine* e, int ilp, int threads, int min_blocks) {
  if (!e) return fail("null engine");
  if (ilp < 0 || ilp > 4) return fail("ahk_set_jit_tuning: ilp must be 0..4");
  if (threads < 0 || threads > 1024 || threads % 32) return fail("ahk_set_jit_tuning: threads must be a multiple of 32");
  if (min_blocks < 0 || min_blocks > 32) return fail("ahk_set_jit_tuning: min_blocks must be 0..32");
  e->jit_ilp = ilp; e->jit_threads = threads; e->jit_minb = min_blocks;
  return 0;
}

int ahk_set_jit_slu(ahk_engine* e, int mode) {
  if (!e) return fail("null engine");
  if (mode < -1 || mode > 2) return fail("ahk_set_jit_slu: mode must be -1..2");
  e->slu_mode = mode;
  return 0;
}

int ahk_jit_slu_stats(const ahk_engine* e, int32_t* out4) {
  if (!e || !out4) return fail("null argument");
  for (int i = 0; i < 4; i++) out4[i] = e->slu_stats[i];
  return 0;
}

int ahk_last_launch(const ahk_engine* e, int32_t* out8) {
  if (!e || !out8) return fail("null argument");
  for (int i = 0; i < 8; i++) out8[i] = e->last[i];
  return 0;
}

int ahk_big_stats(const ahk_engine* e, int64_t* out8) {
  if (!e || !out8) return fail("null argument");
  for (int i = 0; i < 8; i++) out8[i] = 0;
  const BigLin& b = e->big;
  out8[0] = b.last_path;
  if (const Big3Sched* S = b.s3_S) {
    // flop the schedule executes: per pivot step nl divisions-as-multiplies + 1 reciprocal,
    // nl * nu multiply-subtracts (2 flop); substitutions: 2 flop per stored multiplier / row entry
    long long ff = 0, fs = 0, nf[2] = {0, 0};
    for (int m = 0; m < 2; m++)
      for (int kk = 0; kk < S->n; kk++) {
        const Big3Step& st = S->step[m * S->n + kk];
        ff += 1 + st.nl + 2ll * st.nl * st.nu;
        fs += 2ll * st.nl + 2ll * st.nu + 1;
        nf[m] += 1 + st.nl + st.nu;
      }
    out8[1] = ff;                                   // factorisation flop per instance (both matrices)
    out8[2] = fs + 2ll * e->H.hdr.npos + 2ll * S->n;   // flop per (instance, point): substitutions + residual mat-vec + update
    out8[3] = nf[0] > nf[1] ? nf[0] : nf[1];        // factor values stored per matrix
    out8[4] = S->npos;
    out8[5] = S->nslot;
    out8[6] = S->n;
  }
  return 0;
}

int64_t ahk_jit_launch_count(void) { return g_jit_launches.load(); }
int64_t ahk_jit_build_count(void) { return g_jit_builds.load(); }

int ahk_jit_compile(const ahk_circuit* circ, const ahk_options* opts, int n_vary, const int32_t* vary_slots,
                    int kind, int method, int use_step_control, int variant_id, int64_t* cubin_bytes) {
  // variant_id >= 1000 (OP / transient only): free shape, 1000 + rows_per_lane * 64 + lanes, compiled
  // with exactly n + 1 columns like the engine does for ahk_set_jit_shape
  const bool free_shape = variant_id >= 1000;   // (AC: lane-per-chunk static-schedule kernel)
  if (!circ || !opts || kind < 0 || kind > 2 || variant_id < 2 ||
      (!free_shape && variant_id >= (kind == 2 ? AHK_N_AC_VARIANTS : AHK_N_REAL_VARIANTS)))
    return fail("ahk_jit_compile: bad argument");
  try {
    HostProg H;
    build_program(circ, H);
    set_vary_slots(H, n_vary, vary_slots);
    VarInfo v;
    if (free_shape) {
      v = VarInfo{H.hdr.n + 1, (variant_id - 1000) / 64, (variant_id - 1000) % 64};
      if (v.G < 1 || (v.G & (v.G - 1)) || v.G > 32 || v.R < 1 || v.R * v.G < H.hdr.n || H.hdr.n + 1 > 32)
        return fail("ahk_jit_compile: bad free shape");
    } else v = (kind == 2 ? ac_variants() : real_variants())[variant_id];
    if (!fits(v, H.hdr.n)) return fail("variant does not fit this circuit");
    Opts o; std::memcpy(&o, opts, sizeof(Opts));
    std::string spec;
    if (kind == 2) {
      AcLayout AL = make_ac_layout(H, v.G, v.NC, false, free_shape);
      const JitAc ja = {AL.RS, AL.Ac, AL.xc, AL.dxc, AL.Nac, AL.pvc, AL.stride};
      JitCfg cfg = {v.NC, v.R, v.G, free_shape ? 64 : 128, free_shape ? 1 : ac_min_blocks(v.NC), JIT_MODE_AC, 0, 0};
      if (free_shape) {   // the static-schedule kernel (AHK_SLU_SEQ="8 9 9 ...;..." or the nominal guess)
        if (const char* f = getenv("AHK_SLU_SEQ")) {
          for (std::string part : split_semicolon(f)) {
            std::istringstream is(part); int q; std::vector<int> sq;
            while (is >> q) sq.push_back(q);
            if (!sq.empty()) cfg.slu.push_back(sq);
          }
        } else cfg.slu.push_back(slu_guess(H, o.gmin));
      }
      spec = jit_spec_source(H, o, AL.L, cfg, &ja);
    } else {
      const int mode = kind ? MODE_TRAN : MODE_OP;
      Layout Ly = make_layout(H, mode, method, use_step_control, v.G, v.NC, v.R, true);
      JitCfg cfg = {v.NC, v.R, v.G, 128, real_min_blocks(v.NC), mode, method, use_step_control ? 1 : 0};
      if (const char* f = getenv("AHK_JIT_ILP")) cfg.ilp = std::max(1, std::min(4, atoi(f)));        // tuning
      if (const char* f = getenv("AHK_JIT_THREADS")) cfg.threads = std::max(32, atoi(f));            // tuning
      if (const char* f = getenv("AHK_JIT_MINB")) cfg.min_blocks = std::max(1, atoi(f));             // tuning
      // AHK_SLU_SEQ="5 6 2 ..." (tools/jit_sass.py): the static-schedule build along that pivot sequence
      if (const char* f = getenv("AHK_SLU_SEQ")) {
        for (std::string part : split_semicolon(f)) {
          std::istringstream is(part); int q; std::vector<int> sq;
          while (is >> q) sq.push_back(q);
          if (!sq.empty()) cfg.slu.push_back(sq);
        }
      } else if (slu_wanted(nullptr, mode)) cfg.slu.push_back(slu_guess(H, o.gmin));
      spec = jit_spec_source(H, o, Ly, cfg);
    }
    std::string err;
    bool cached = false;
    std::vector<char> cubin = jit::compile(spec, err, &cached);
    if (cubin.empty()) return fail(err);
    if (!cached) g_jit_builds++;
    if (cubin_bytes) *cubin_bytes = (int64_t)cubin.size();
    return 0;
  } catch (const std::exception& ex) { return fail(ex.what()); }
}

int ahk_set_group(ahk_engine* e, int lanes) {
  if (!e) return fail("null engine");
  if (lanes != 0 && lanes != 1 && lanes != 2 && lanes != 4 && lanes != 8 && lanes != 16 && lanes != 32)
    return fail("group must be 0,1,2,4,8,16,32");
  e->group = lanes;
  return 0;
}

int ahk_variant_count(int kind) { return kind == 1 ? AHK_N_AC_VARIANTS : AHK_N_REAL_VARIANTS; }

int ahk_variant_info(int kind, int id, int32_t* nc_r_g) {
  if (!nc_r_g || id < 0 || id >= ahk_variant_count(kind)) return fail("bad variant id");
  const VarInfo& v = (kind == 1 ? ac_variants() : real_variants())[id];
  nc_r_g[0] = v.NC; nc_r_g[1] = v.R; nc_r_g[2] = v.G;
  return 0;
}

int ahk_set_variant(ahk_engine* e, int real_id, int ac_id) {
  if (!e) return fail("null engine");
  if (real_id >= AHK_N_REAL_VARIANTS || ac_id >= AHK_N_AC_VARIANTS) return fail("bad variant id");
  if (real_id >= 0 && !fits(real_variants()[real_id], e->n)) return fail("variant does not fit this circuit");
  if (ac_id >= 0 && !fits(ac_variants()[ac_id], e->n)) return fail("variant does not fit this circuit");
  e->force_vid = real_id < 0 ? -1 : real_id;
  e->force_ac_vid = ac_id < 0 ? -1 : ac_id;
  // the generic warp-per-instance variant forced on a large LINEAR circuit: the general large-n
  // engine instead of the factor-once linear path (tests, comparisons)
  e->force_general = real_id == 1 && e->n > AHK_SMALL_NMAX;
  return 0;
}

int ahk_h2d_columns(double* dst, const double* src_host, int n_rows, int64_t row_len, int64_t lo,
                    int64_t hi, void* stream) {
  if (!dst || !src_host || n_rows < 0 || lo < 0 || hi < lo || hi > row_len) return fail("ahk_h2d_columns: bad argument");
  if (n_rows == 0 || hi == lo) return 0;
  CK(cudaMemcpy2DAsync(dst, (size_t)(hi - lo) * sizeof(double), src_host + lo, (size_t)row_len * sizeof(double),
                       (size_t)(hi - lo) * sizeof(double), (size_t)n_rows, cudaMemcpyHostToDevice,
                       (cudaStream_t)stream));
  return 0;
}

int ahk_pack_rows(const double* t_out, const double* x_out, const int32_t* rows, const int64_t* offsets,
                  int64_t B, int64_t max_rows, int n, double* t_packed, double* x_packed, void* stream) {
  if (!t_out || !x_out || !rows || !offsets || !t_packed || !x_packed || B <= 0 || max_rows <= 0 || n <= 0)
    return fail("ahk_pack_rows: bad argument");
  static_assert(sizeof(long long) == sizeof(int64_t), "offsets are 64-bit");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = (int)std::min<int64_t>(B, (int64_t)sms * 8);   // 8 resident blocks of 256 threads per SM
  k_pack_rows<<<blocks, 256, 0, (cudaStream_t)stream>>>(t_out, x_out, rows, (const long long*)offsets, B, max_rows, n,
                                                        t_packed, x_packed);
  CK(cudaGetLastError());
  g_launches++;
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
  if (e->n > AHK_SMALL_NMAX && !e->H.nonlinear && !e->force_general) {
    // large LINEAR circuits: factor once per instance, static sparse schedule (biglin*.cuh)
    double v = 0.0;
    return e->big.run(e, batch, -1, &v, 1, x, iters, status, resid, s, g_err, g_launches);
  }
  // (n > 64, non-linear: the general large-n engine inside run_dc_small)
  return run_dc_small(e, batch, -1, nullptr, 1, x0, use_guess, x, resid, iters, status, s);
}

// Operating point with HOST buffers: the batch is cut into chunks; one copy-in stream, one
// kernel stream PER CHUNK SLOT and one copy-out stream, so that the PCIe traffic of both
// directions overlaps the kernels — and the kernels overlap each other: these kernels are
// latency bound (a quarter of the batch takes as long as all of it), so the chunks must
// run side by side, not back to back.  Asynchronous: `stream` waits for the
// last copy; the caller synchronises `stream` before reading the results.
int ahk_op_host(ahk_engine* e, int64_t B, int n_vary, const int32_t* vary_slots, const double* vary_host,
                int use_guess, double* x_host, double* resid_host, int32_t* iters_host,
                int32_t* status_host, int nchunks, void* stream) {
  if (!e || B <= 0 || n_vary < 0 || (n_vary > 0 && (!vary_slots || !vary_host)) || !x_host || !iters_host ||
      !status_host)
    return fail("ahk_op_host: bad argument");
  if (e->n > AHK_SMALL_NMAX) return fail("ahk_op_host: small circuits only (n <= 64)");
  cudaStream_t user = (cudaStream_t)stream;
  auto& hp = e->hp;
  const int NS = ahk_engine::HostPipe::NS;
  if (!hp.ready) {
    CK(cudaStreamCreateWithFlags(&hp.s_org, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&hp.s_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&hp.s_out, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&hp.ev_start, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&hp.ev_done, cudaEventDisableTiming));
    for (int i = 0; i < NS; i++) {
      CK(cudaStreamCreateWithFlags(&hp.s_comp[i], cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&hp.ev_in[i], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&hp.ev_comp[i], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&hp.ev_out[i], cudaEventDisableTiming));
    }
    hp.ready = true;
  }
  const int n = e->n;
  if (nchunks == 0) {
    // zero-copy: the kernel reads the parameters from and writes the results to the
    // caller's PINNED host buffers directly (unified addressing maps them into the device's
    // address space): no DMA copies, no staging — the PCIe traffic of both directions
    // overlaps the computation inside the one kernel
    ahk_batch b; b.B = B; b.n_vary = n_vary; b.vary_slots = vary_slots; b.vary = vary_host;
    if (set_batch(e, &b, user)) return -1;
    return run_dc_small(e, &b, -1, nullptr, 1, nullptr, use_guess, x_host, resid_host, iters_host, status_host, user);
  }
  if (nchunks < 1) nchunks = 1;
  if ((int64_t)nchunks > B) nchunks = (int)B;
  const int64_t Bc_max = (B + nchunks - 1) / nchunks;
  // workspace of one slot: vary[n_vary][Bc] | x[Bc][n] | resid[Bc][n] | iters[Bc] | status[Bc]
  const size_t o_x = (size_t)n_vary * Bc_max * 8, o_r = o_x + (size_t)Bc_max * n * 8,
               o_i = o_r + (size_t)Bc_max * n * 8, o_s = o_i + (((size_t)Bc_max * 4 + 7) & ~size_t(7)),
               need = o_s + (size_t)Bc_max * 4 + 64;
  if (need > hp.ws_bytes) {
    CK(cudaDeviceSynchronize());
    if (hp.graph_exec) { cudaGraphExecDestroy(hp.graph_exec); hp.graph_exec = nullptr; }   // it points into ws
    hp.graph_key.clear(); hp.last_key.clear();
    for (int i = 0; i < NS; i++) { if (hp.ws[i]) cudaFree(hp.ws[i]); hp.ws[i] = nullptr; }
    for (int i = 0; i < NS; i++) CK(cudaMalloc(&hp.ws[i], need));
    hp.ws_bytes = need;
  }
  // the varied slots must be installed before any chunk launches (may synchronise once)
  {
    ahk_batch b0; b0.B = Bc_max; b0.n_vary = n_vary; b0.vary_slots = vary_slots; b0.vary = (const double*)hp.ws[0];
    const long long epoch = e->slots_epoch;
    if (set_batch(e, &b0, hp.s_comp[0])) return -1;
    if (e->slots_epoch != epoch) CK(cudaStreamSynchronize(hp.s_comp[0]));   // table upload: before ANY chunk
  }
  // ---- graph replay / capture / eager ------------------------------------------------------
  const std::vector<long long> key = {B, n_vary, (long long)(size_t)vary_host, (long long)(size_t)x_host,
                                      (long long)(size_t)resid_host, (long long)(size_t)iters_host,
                                      (long long)(size_t)status_host, nchunks, use_guess, e->slots_epoch,
                                      e->opts_epoch, e->jit_mode, e->jit_min_batch, e->force_vid, e->group};
  static const bool no_graph = getenv("AHK_NO_GRAPH") != nullptr;
  if (hp.graph_exec && key == hp.graph_key) {
    CK(cudaGraphLaunch(hp.graph_exec, user));
    g_launches += hp.graph_launches;
    g_jit_launches += hp.graph_jit_launches;
    return 0;
  }
  const bool capture = !no_graph && key == hp.last_key;   // second identical call: kernels are built
  hp.last_key = key;
  const long long l0 = g_launches.load(), j0 = g_jit_launches.load();
  cudaStream_t org = capture ? hp.s_org : user;
  if (capture) CK(cudaStreamBeginCapture(hp.s_org, cudaStreamCaptureModeRelaxed));
  auto issue = [&]() -> int {
  CK(cudaEventRecord(hp.ev_start, org));
  CK(cudaStreamWaitEvent(hp.s_in, hp.ev_start, 0));
  CK(cudaStreamWaitEvent(hp.s_out, hp.ev_start, 0));
  int64_t lo = 0;
  for (int c = 0; c < nchunks; c++) {
    const int64_t hi = std::min<int64_t>(B, lo + Bc_max), Bc = hi - lo;
    if (Bc <= 0) break;
    const int slot = c % NS;
    cudaStream_t sc = hp.s_comp[slot];
    unsigned char* w = hp.ws[slot];
    double* d_vary = (double*)w; double* d_x = (double*)(w + o_x); double* d_r = (double*)(w + o_r);
    int32_t* d_i = (int32_t*)(w + o_i); int32_t* d_s = (int32_t*)(w + o_s);
    // copy-in: the slot's previous chunk must have been computed AND copied out
    if (c >= NS) { CK(cudaStreamWaitEvent(hp.s_in, hp.ev_out[slot], 0)); }
    if (n_vary)
      CK(cudaMemcpy2DAsync(d_vary, (size_t)Bc * 8, vary_host + lo, (size_t)B * 8, (size_t)Bc * 8, (size_t)n_vary,
                           cudaMemcpyHostToDevice, hp.s_in));
    CK(cudaEventRecord(hp.ev_in[slot], hp.s_in));
    // kernels
    CK(cudaStreamWaitEvent(sc, hp.ev_in[slot], 0));   // (ev_in was recorded after ev_start / ev_out[slot])
    ahk_batch b; b.B = Bc; b.n_vary = n_vary; b.vary_slots = vary_slots; b.vary = d_vary;
    if (run_dc_small(e, &b, -1, nullptr, 1, nullptr, use_guess, d_x, resid_host ? d_r : nullptr, d_i, d_s, sc))
      return -1;
    CK(cudaEventRecord(hp.ev_comp[slot], sc));
    // copy-out
    CK(cudaStreamWaitEvent(hp.s_out, hp.ev_comp[slot], 0));
    CK(cudaMemcpyAsync(x_host + lo * n, d_x, (size_t)Bc * n * 8, cudaMemcpyDeviceToHost, hp.s_out));
    if (resid_host)
      CK(cudaMemcpyAsync(resid_host + lo * n, d_r, (size_t)Bc * n * 8, cudaMemcpyDeviceToHost, hp.s_out));
    CK(cudaMemcpyAsync(iters_host + lo, d_i, (size_t)Bc * 4, cudaMemcpyDeviceToHost, hp.s_out));
    CK(cudaMemcpyAsync(status_host + lo, d_s, (size_t)Bc * 4, cudaMemcpyDeviceToHost, hp.s_out));
    CK(cudaEventRecord(hp.ev_out[slot], hp.s_out));
    lo = hi;
  }
  CK(cudaEventRecord(hp.ev_done, hp.s_out));
  CK(cudaStreamWaitEvent(org, hp.ev_done, 0));
  return 0;
  };
  const int rc = issue();
  if (!capture) return rc;
  cudaGraph_t graph = nullptr;
  cudaError_t ce = cudaStreamEndCapture(hp.s_org, &graph);
  if (rc != 0) { if (graph) cudaGraphDestroy(graph); return rc; }
  if (ce != cudaSuccess) return fail(std::string("ahk_op_host: graph capture: ") + cudaGetErrorString(ce));
  if (hp.graph_exec) { cudaGraphExecDestroy(hp.graph_exec); hp.graph_exec = nullptr; }
  ce = cudaGraphInstantiate(&hp.graph_exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ce != cudaSuccess) { hp.graph_exec = nullptr; return fail(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce)); }
  hp.graph_key = key;
  hp.graph_launches = g_launches.load() - l0;       // kernels of one replay (counted at capture)
  hp.graph_jit_launches = g_jit_launches.load() - j0;
  CK(cudaGraphLaunch(hp.graph_exec, user));
  return 0;
}

int ahk_dc_sweep(ahk_engine* e, const ahk_batch* batch, int src_elem, const double* sweep_values,
                 int npts, const double* x0, int use_guess, double* x_out, int32_t* iters,
                 int32_t* status, void* stream) {
  if (!e || !sweep_values || npts <= 0 || !x_out || !iters || !status)
    return fail("ahk_dc_sweep: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (set_batch(e, batch, s)) return -1;
  bool ok = false;
  // (a source with a waveform sweeps too: dc_analysis sets its dc_value, which is what V(time=None)
  // returns — VSource.py:92-95)
  for (auto& so : e->H.src) if (so.elem == src_elem) ok = true;
  if (!ok) return fail("ahk_dc_sweep: src_elem is not a V/I source");
  if (e->n > AHK_SMALL_NMAX && !e->H.nonlinear && !e->force_general)
    return e->big.run(e, batch, src_elem, sweep_values, npts, x_out, iters, status, nullptr, s,
                      g_err, g_launches);
  if (to_scratch(e, sweep_values, npts, s)) return -1;
  return run_dc_small(e, batch, src_elem, e->d_scratch, npts, x0, use_guess, x_out, nullptr, iters,
                      status, s);
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
  cudaStream_t s = (cudaStream_t)stream;
  if (set_batch(e, batch, s)) return -1;
  int vid; Layout Ly; Launch L;
  jit::Kernel* jk = nullptr;
  e->slu_c1_hint = tstep > 0 ? (method == AHK_TRAP ? 2.0 : 1.0) / tstep : 0.0;
  if (prepare_small(e, batch->B, MODE_TRAN, method, use_step_control ? 1 : 0, &vid, &Ly, &L, &jk)) return -1;
  note_launch(e, vid, L, jk != nullptr, 1);
  if (jk) {
    KTranJ kj = {batch->vary, batch->B, x0,
                 TranArgs{tstart, tstep, tstop, method, use_step_control, max_rows, t_out, x_out, rows, counters},
                 status, nullptr};
    std::string err;
    if (jk->slu && !jk->calibrated) {
      const int usc = use_step_control ? 1 : 0;
      if (slu_calibrate(e, MODE_TRAN, method, usc, jk, s,
            [&](jit::Kernel* q, int* rec) { kj.pivrec = rec; const bool ok = jit::launch(*q, 1, L.threads, L.smem, s, &kj, err); kj.pivrec = nullptr; return ok ? 0 : fail(err); },
            [&]() { return prepare_small(e, batch->B, MODE_TRAN, method, usc, &vid, &Ly, &L, &jk); })) return -1;
      if (!jk) return fail("static-schedule calibration lost the specialised kernel");
    }
    return jit_launched(jit::launch(*jk, L.blocks, L.threads, L.smem, s, &kj, err), err);
  }
  if (ensure_big_workspace(e, batch->B, Ly, s)) return -1;
  KTran k;
  k.b = make_base(e, batch, Ly);
  k.x0 = x0;
  k.a = TranArgs{tstart, tstep, tstop, method, use_step_control, max_rows, t_out, x_out, rows, counters};
  k.status = status;
  return launched(launch_tran(k, L, s));
}

int ahk_ac(ahk_engine* e, const ahk_batch* batch, const double* x_op, const double* omegas,
           int npts, double* x_out, int32_t* status, void* stream) {
  if (!e || !omegas || npts <= 0 || !x_out || !status) return fail("ahk_ac: bad argument");
  if (e->H.nonlinear && !x_op) return fail("ahk_ac: non-linear circuit needs x_op");
  if (e->n > AHK_BIG_NMAX) return fail("circuit too large: n = " + std::to_string(e->n) + " > " + std::to_string(AHK_BIG_NMAX));
  const bool big = e->n > AHK_SMALL_NMAX;    // general large-n engine: complex matrix in the global workspace
  cudaStream_t s = (cudaStream_t)stream;
  if (set_batch(e, batch, s)) return -1;
  if (to_scratch(e, omegas, npts, s)) return -1;
  // frequency chunks: enough (instance, chunk) units to fill the machine
  int nchunks = 1;
  {
    long long want = (long long)e->sm_count * 64;
    while (nchunks < npts && batch->B * nchunks < want && npts / (nchunks * 2) >= 4) nchunks *= 2;
    if (const char* f = getenv("AHK_AC_CHUNKS")) nchunks = std::max(1, std::min(npts, atoi(f)));   // tuning
  }
  long long units = batch->B * nchunks;
  int vid = 1;                                 // big: the generic warp-per-instance variant
  if (!big && pick_variant(e, ac_variants(), AHK_N_AC_VARIANTS, e->force_ac_vid, units, ac_min_blocks,
                           [&](const VarInfo& q) { return make_ac_layout(e->H, q.G, q.NC).stride; }, &vid))
    return -1;
  const VarInfo& v = ac_variants()[vid];
  AcLayout AL = make_ac_layout(e->H, v.G, v.NC, big);
  if (big) {
    if (e->jit_mode == 1) return fail("specialised kernels need n <= 64");
    Layout Lw = AL.L; Lw.RS = 2 * AL.RS;       // (complex elements: twice the doubles per row)
    if (ensure_big_workspace(e, units, Lw, s)) return -1;
  }
  Launch L;
  if (plan(e, units, vid, v, AL.stride, L, ac_min_blocks(v.NC), nullptr)) return -1;
  {
    long long nb = (batch->B + 255) / 256;
    k_fill_i32<<<(int)nb, 256, 0, s>>>(status, batch->B, AHK_ST_OK);
    CK(cudaGetLastError());
    g_launches++;
  }
  // One lane per (instance, frequency chunk) with zgesv's LU along a static schedule
  // (ac.cuh: zslu_solve_thread): the solve of a sparse circuit is a few hundred flop, and 16 lanes
  // sharing a dense complex Gauss-Jordan spend theirs on zeros (C1: 0.84 ms -> see DESIGN.md).
  if (wants_jit(e, batch->B) && !big && e->n <= 31 && e->force_ac_vid < 0 && slu_active(e, JIT_MODE_AC, 0, 0)) {
    // lanes per instance = frequency chunks: every lane is a serial chain, so as many as keep
    // >= 3 frequencies per lane (the group shares its instance's set-up; C1, 100 frequencies:
    // 4 / 8 / 16 / 32 lanes = 0.57 / 0.45 / 0.38 / 0.34 ms, gpurun_out/r3i_c1.txt)
    int Gc = 1;
    while (Gc < 32 && npts / (Gc * 2) >= 3) Gc *= 2;
    if (const char* f = getenv("AHK_AC_CHUNKS")) { Gc = 1; while (Gc * 2 <= atoi(f) && Gc < 32) Gc *= 2; }   // tuning
    const VarInfo vs = {e->n + 1, e->n, Gc};
    const int vids = 1000 + e->n * 64 + Gc;
    const AcLayout ALs = make_ac_layout(e->H, Gc, vs.NC, false, true);
    int thr = 64;
    if (const char* f = getenv("AHK_AC_THREADS")) thr = std::max(32, std::min(256, atoi(f) / 32 * 32));   // tuning
    Launch Ls;
    Ls.vid = vids; Ls.G = Gc; Ls.threads = std::max(thr, Gc);
    Ls.smem = (size_t)(Ls.threads / Gc) * ALs.stride * sizeof(double);
    while (Ls.smem > e->smem_max && Ls.threads > 32 && Ls.threads / 2 >= Gc) { Ls.threads /= 2; Ls.smem = (size_t)(Ls.threads / Gc) * ALs.stride * sizeof(double); }
    if (Ls.smem <= e->smem_max) {
      const int ipb = Ls.threads / Gc;
      Ls.blocks = (int)((batch->B + ipb - 1) / ipb);
      const JitAc ja = {ALs.RS, ALs.Ac, ALs.xc, ALs.dxc, ALs.Nac, ALs.pvc, ALs.stride};
      jit::Kernel* jk = nullptr;
      int acminb = 1;
      if (const char* f = getenv("AHK_AC_MINB")) acminb = std::max(1, std::min(16, atoi(f)));   // tuning
      auto get = [&]() { return get_jit(e, batch->B, JIT_MODE_AC, vids, vs, ALs.L, Ls, -acminb, 0, 0, &jk, &ja); };
      if (get()) return -1;
      if (jk && jk->slu) {
        KAcJ kj = {batch->vary, batch->B, AcArgs{x_op, e->d_scratch, npts, 0, npts, x_out, status}, Gc, nullptr};
        std::string err;
        if (!jk->calibrated) {
          if (slu_calibrate(e, JIT_MODE_AC, 0, 0, jk, s,
                [&](jit::Kernel* q, int* rec) {
                  kj.pivrec = rec;
                  const bool ok = jit::launch(*q, 1, Ls.threads, Ls.smem, s, &kj, err);
                  kj.pivrec = nullptr;
                  return ok ? 0 : fail(err);
                },
                [&]() { return get(); })) return -1;
          // (the calibration block may have flagged a singular instance: status is rewritten)
          long long nb = (batch->B + 255) / 256;
          k_fill_i32<<<(int)nb, 256, 0, s>>>(status, batch->B, AHK_ST_OK);
          CK(cudaGetLastError());
        }
        if (jk && jk->slu) {
          note_launch(e, vids, Ls, true, 2, true);
          return jit_launched(jit::launch(*jk, Ls.blocks, Ls.threads, Ls.smem, s, &kj, err), err);
        }
      }
    }
  }
  if (wants_jit(e, batch->B) && v.NC > 0) {   // the kernel compiled for this circuit (jit_gen.hpp)
    VarInfo vj = v;
    vj.NC = e->n + 1;                         // exactly n + 1 columns, no padding
    const AcLayout ALj = make_ac_layout(e->H, vj.G, vj.NC);
    Launch Lj;
    if (plan(e, units, vid, vj, ALj.stride, Lj, ac_min_blocks(v.NC), nullptr)) return -1;
    const JitAc ja = {ALj.RS, ALj.Ac, ALj.xc, ALj.dxc, ALj.Nac, ALj.pvc, ALj.stride};
    jit::Kernel* jk = nullptr;
    if (get_jit(e, batch->B, JIT_MODE_AC, vid, vj, ALj.L, Lj, ac_min_blocks(v.NC), 0, 0, &jk, &ja)) return -1;
    if (jk) {
      note_launch(e, vid, Lj, true, 2, true);
      KAcJ kj = {batch->vary, batch->B, AcArgs{x_op, e->d_scratch, npts, 0, npts, x_out, status}, nchunks, nullptr};
      std::string err;
      return jit_launched(jit::launch(*jk, Lj.blocks, Lj.threads, Lj.smem, s, &kj, err), err);
    }
  }
  note_launch(e, vid, L, false, 2, true);
  KAc k;
  k.p = e->dprog; k.o = e->o; k.AL = AL;
  k.vary = batch->vary; k.B = batch->B; k.nchunks = nchunks;
  k.Ag = big ? e->bigA : nullptr;
  k.Voff = big ? (long long)2 * e->n * AL.RS : 0;
  k.Astride = big ? k.Voff + AL.L.vlen : 0;
  k.a = AcArgs{x_op, e->d_scratch, npts, 0, npts, x_out, status};
  return launched(launch_ac(k, L, s));
}

}  // extern "C"
