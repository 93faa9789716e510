// biglin.cuh — large linear circuits (n > 64): OP and DC sweep, one WARP per
// instance, e.g. the 256-node R-2R ladder of tests/r2r (n = 257).
//
// For a linear circuit op_analysis (dc_analysis.py:598-687) is exactly
//   pass 1: x1 = x0 + solve(mna + Gmin, -((mna + Gmin) x0 + N))     (1 iteration)
//   pass 2: x2 = x1 + solve(mna,        -( mna         x1 + N))     (1 iteration)
// and dc_analysis (dc_analysis.py:461-595) repeats it per sweep value with
// x0 = previous x2.  Only N changes along the sweep, so each instance factors its
// two matrices ONCE (the reference refactors both at every point, 2 dgesv per
// point) and every sweep point costs two pairs of triangular solves.
//
// LU = right-looking Gaussian elimination with partial pivoting, the same pivot
// rule as LAPACK dgesv; the matrix is stored dense (row-major, in a per-warp
// global workspace that stays L2-resident) but the elimination only touches rows
// whose multiplier is non-zero and pivot-row entries that are non-zero, which is
// bit-identical to the dense algorithm (a - 0*b == a) and costs O(nnz) instead of
// O(n^3) on circuit matrices.  L and U are kept as compact per-step lists.
#pragma once
#include <atomic>
#include <string>

#include "core.cuh"
#include "program_build.hpp"

namespace ahk {

struct BigK {
  Prog p; Opts o;
  const double* vary; long long B;
  int src_elem; const double* sweep; int npts;
  const double* x0;
  double* x_out; int* iters; int* status; double* resid;
  double* wsA;                 // [nwarps][n*RS]
  double* wsLv; double* wsUv;  // [nwarps][2][cap]
  unsigned short* wsLi; unsigned short* wsUi;
  int cap, RS;
  const int* list;             // null: all B instances; else list[0] = count, list[1..] = instances
};

struct BigSmem {   // offsets in bytes inside one warp's dynamic shared memory
  int Mv, x, x1, b, dx, N, upiv, pend_l, pv, Lptr, Uptr, prow, stepof, pend_r, pc, total;
};

inline BigSmem big_smem(int n, int npos) {
  BigSmem s; int off = 0;
  auto take = [&](int bytes) { int o = off; off += (bytes + 15) & ~15; return o; };
  s.Mv = take(8 * npos); s.x = take(8 * n); s.x1 = take(8 * n); s.b = take(8 * n);
  s.dx = take(8 * n); s.N = take(8 * n); s.upiv = take(16 * n); s.pend_l = take(8 * n);
  s.pv = take(8 * n); s.Lptr = take(8 * (n + 1)); s.Uptr = take(8 * (n + 1));
  s.prow = take(4 * n); s.stepof = take(2 * n); s.pend_r = take(2 * n); s.pc = take(2 * n);
  s.total = off;
  return s;
}

#ifdef __CUDACC__
#define BIG_NONE 0xFFFFu

__global__ void __launch_bounds__(32) k_biglin(const __grid_constant__ BigK k, const BigSmem S) {
  extern __shared__ unsigned char bsm[];
  const Prog& p = k.p;
  const int n = p.n, RS = k.RS, lane = threadIdx.x;
  const unsigned FULL = 0xffffffffu, lt = (1u << lane) - 1u;
  double* Mv = (double*)(bsm + S.Mv); double* x = (double*)(bsm + S.x);
  double* x1 = (double*)(bsm + S.x1); double* b = (double*)(bsm + S.b);
  double* dx = (double*)(bsm + S.dx); double* N = (double*)(bsm + S.N);
  double* upiv = (double*)(bsm + S.upiv); double* pend_l = (double*)(bsm + S.pend_l);
  double* pv = (double*)(bsm + S.pv);
  int* Lptr = (int*)(bsm + S.Lptr); int* Uptr = (int*)(bsm + S.Uptr);
  unsigned short* prow = (unsigned short*)(bsm + S.prow);
  unsigned short* stepof = (unsigned short*)(bsm + S.stepof);
  unsigned short* pend_r = (unsigned short*)(bsm + S.pend_r);
  unsigned short* pc = (unsigned short*)(bsm + S.pc);
  double* A = k.wsA + (size_t)blockIdx.x * n * RS;
  double* Lv = k.wsLv + (size_t)blockIdx.x * 2 * k.cap;
  double* Uv = k.wsUv + (size_t)blockIdx.x * 2 * k.cap;
  unsigned short* Li = k.wsLi + (size_t)blockIdx.x * 2 * k.cap;
  unsigned short* Ui = k.wsUi + (size_t)blockIdx.x * 2 * k.cap;

  const long long count = k.list ? (long long)k.list[0] : k.B;
  for (long long idx = blockIdx.x; idx < count; idx += gridDim.x) {
    const long long inst = k.list ? (long long)k.list[1 + idx] : idx;
    auto P = [&](int slot) -> double {
      int v = p.slot_vary[slot];
      return v < 0 ? p.nominal[slot] : k.vary[(long long)v * k.B + inst];
    };
    // per-instance values of the MNA stamps (deterministic gather)
    for (int q = lane; q < p.npos; q += 32) {
      double v = 0.0;
      for (int t = p.term_ptr[q]; t < p.term_ptr[q + 1]; t++) {
        int kd = p.term_kind[t];
        double sg = kd < 0 ? -1.0 : 1.0;
        int ak = kd < 0 ? -kd : kd;
        if (ak == 1) v += sg * (1. / P(p.term_slot[t]));
        else if (ak == 2) v += sg * P(p.term_slot[t]);
        else v += sg;
      }
      Mv[q] = v;
    }
    __syncwarp();
    int singular = 0;
    // ---- factor mna + Gmin (m = 0) and mna (m = 1) --------------------------------
    for (int m = 0; m < 2; m++) {
      for (int i = lane; i < n * RS; i += 32) A[i] = 0.0;
      for (int r = lane; r < n; r += 32) stepof[r] = BIG_NONE;
      __syncwarp();
      for (int q = lane; q < p.npos; q += 32)
        A[(int)p.prowi[q] * RS + p.pcol[q]] = Mv[q] + ((m == 0 && p.pdiag[q]) ? k.o.gmin : 0.0);
      __syncwarp();
      int Lbase = m * k.cap, Ubase = m * k.cap;
      int* Lp = Lptr + m * (n + 1);
      int* Up = Uptr + m * (n + 1);
      if (lane == 0) { Lp[0] = Lbase; Up[0] = Ubase; }
      for (int kk = 0; kk < n; kk++) {
        double best = -1.0; int bi = -1;
        for (int r = lane; r < n; r += 32) {
          if (stepof[r] != BIG_NONE) continue;
          double v = fabs(A[r * RS + kk]);
          if (bi < 0 || v > best) { best = v; bi = r; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          double ov = __shfl_xor_sync(FULL, best, o);
          int oi = __shfl_xor_sync(FULL, bi, o);
          bool take = (oi >= 0) && (bi < 0 || ov > best || (ov == best && oi < bi));
          if (take) { best = ov; bi = oi; }
        }
        if (best == 0.0) { singular = 1; break; }
        const double piv = A[bi * RS + kk];
        const double inv = 1.0 / piv;
        if (lane == 0) { prow[m * n + kk] = (unsigned short)bi; stepof[bi] = (unsigned short)kk; upiv[m * n + kk] = piv; }
        __syncwarp();
        // rows with a non-zero multiplier (ascending row order, ballot compaction)
        int nL = 0;
        for (int r0 = 0; r0 < n; r0 += 32) {
          int r = r0 + lane;
          double a = 0.0;
          bool f = false;
          if (r < n && stepof[r] == BIG_NONE) { a = A[r * RS + kk]; f = a != 0.0; }
          unsigned bal = __ballot_sync(FULL, f);
          if (f) { int i = nL + __popc(bal & lt); pend_r[i] = (unsigned short)r; pend_l[i] = a * inv; }
          nL += __popc(bal);
        }
        // non-zero entries of the pivot row right of the pivot (ascending column order)
        int nU = 0;
        for (int j0 = kk + 1; j0 < n; j0 += 32) {
          int j = j0 + lane;
          double u = 0.0;
          bool f = false;
          if (j < n) { u = A[bi * RS + j]; f = u != 0.0; }
          unsigned bal = __ballot_sync(FULL, f);
          if (f) { int i = nU + __popc(bal & lt); pc[i] = (unsigned short)j; pv[i] = u; }
          nU += __popc(bal);
        }
        __syncwarp();
        if (Lbase + nL > (m + 1) * k.cap || Ubase + nU > (m + 1) * k.cap) { singular = 2; break; }
        for (int i = lane; i < nL; i += 32) { Li[Lbase + i] = pend_r[i]; Lv[Lbase + i] = pend_l[i]; }
        for (int i = lane; i < nU; i += 32) { Ui[Ubase + i] = pc[i]; Uv[Ubase + i] = pv[i]; }
        for (int t = lane; t < nL * nU; t += 32) {
          int i = t / nU, j = t - i * nU;
          A[(int)pend_r[i] * RS + pc[j]] -= pend_l[i] * pv[j];
        }
        Lbase += nL; Ubase += nU;
        if (lane == 0) { Lp[kk + 1] = Lbase; Up[kk + 1] = Ubase; }
        __syncwarp();
      }
      if (singular) break;
    }
    // ---- sweep -----------------------------------------------------------------------
    for (int i = lane; i < n; i += 32) x[i] = k.x0 ? k.x0[inst * n + i] : 0.0;
    __syncwarp();
    for (int kp = 0; kp < k.npts; kp++) {
      const long long orow = inst * k.npts + kp;
      if (singular) {
        for (int i = lane; i < n; i += 32) k.x_out[orow * n + i] = nan("");
        if (lane == 0) { k.iters[orow] = 0; k.status[orow] = AHK_ST_SINGULAR; }
        continue;
      }
      for (int i = lane; i < n; i += 32) N[i] = 0.0;
      __syncwarp();
      if (lane == 0) {
        for (int s = 0; s < p.nsrc; s++) {
          const SrcOp& so = p.src[s];
          // (a source with a waveform enters OP / DC with its dc_value: time is None, VSource.py:92-95)
          double v = (so.elem == k.src_elem) ? k.sweep[kp] : P(so.pbase);
          if (so.is_v) N[so.row0] = -1.0 * v;
          else { if (so.row0 >= 0) N[so.row0] += v; if (so.row1 >= 0) N[so.row1] -= v; }
        }
      }
      __syncwarp();
      for (int m = 0; m < 2; m++) {
        const double* xin = m == 0 ? x : x1;
        for (int r = lane; r < n; r += 32) {   // residuo = M.x + N (dc_analysis.py:816)
          double s = 0.0;
          for (int q = p.row_ptr[r]; q < p.row_ptr[r + 1]; q++)
            s += (Mv[q] + ((m == 0 && p.pdiag[q]) ? k.o.gmin : 0.0)) * xin[p.pcol[q]];
          double rr = s + N[r];
          b[r] = -rr;
          if (m == 1 && k.resid) k.resid[orow * n + r] = rr;
        }
        __syncwarp();
        const int* Lp = Lptr + m * (n + 1);
        const int* Up = Uptr + m * (n + 1);
        for (int kk = 0; kk < n; kk++) {       // forward: b[r] -= l * b[pivot row]
          int e0 = Lp[kk], e1 = Lp[kk + 1];
          if (e0 == e1) continue;
          double bp = b[prow[m * n + kk]];
          if (bp != 0.0)
            for (int e = e0 + lane; e < e1; e += 32) b[Li[e]] -= Lv[e] * bp;
          __syncwarp();
        }
        for (int kk = n - 1; kk >= 0; kk--) {  // backward
          int e0 = Up[kk], e1 = Up[kk + 1];
          double s = 0.0;
          for (int e = e0 + lane; e < e1; e += 32) s += Uv[e] * dx[Ui[e]];
          if (e1 - e0 > 1) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
          } else {
            s = __shfl_sync(FULL, s, 0);
          }
          if (lane == 0) dx[kk] = (b[prow[m * n + kk]] - s) / upiv[m * n + kk];
          __syncwarp();
        }
        double* xo = m == 0 ? x1 : x;
        for (int i = lane; i < n; i += 32) xo[i] = xin[i] + dx[i];   // td = 1: no locked nodes
        __syncwarp();
      }
      int bad = 0;
      for (int i = lane; i < n; i += 32) {     // gmin_check, results.py:493-523
        bool isv = i < p.nv1;
        double er = isv ? k.o.ver : k.o.ier, ea = isv ? k.o.vea : k.o.iea;
        if (fabs(x[i] - x1[i]) > er * fmax(fabs(x1[i]), fabs(x[i])) + ea) bad = 1;
        k.x_out[orow * n + i] = x[i];
      }
      bad = __any_sync(FULL, bad);
      if (lane == 0) { k.iters[orow] = 2; k.status[orow] = AHK_ST_OK | (bad ? AHK_ST_GMIN_CHECK : 0); }
      __syncwarp();
    }
  }
}
#endif  // __CUDACC__

}  // namespace ahk

#ifdef __CUDACC__
struct ahk_engine;
// host driver state of the large-n path (owned by ahk_engine)
namespace ahk { struct Big3Sched; }   // biglin3.cuh
struct BigLin {
  // dense-workspace kernel (biglin.cuh): fallback for instances whose rows outgrow the
  // shared-memory slots of the sparse kernel, and for n > 1024
  double* wsA = nullptr; double* wsLv = nullptr; double* wsUv = nullptr;
  unsigned short* wsLi = nullptr; unsigned short* wsUi = nullptr;
  double* d_sweep = nullptr; int sweep_cap = 0;
  int nwarps = 0, cap = 0;
  // sparse shared-memory kernel (biglin2.cuh)
  double* s2Lv = nullptr; double* s2Uv = nullptr; unsigned short* s2Li = nullptr; unsigned short* s2Ui = nullptr;
  int* d_ovf = nullptr; long long ovf_cap = 0;
  int nwarps2 = 0, cap2 = 0;
  // static-schedule kernels (biglin3.cuh): schedule blob + structure-of-arrays buffers
  void* s3_sched = nullptr;      // device copy of the schedule arrays
  const void* s3_ptr[11] = {};   // step, prow, cand_slot, cand_row, l_slot, l_row, u_slot, u_col, upd, rslot, rindex
  int s3_n = 0, s3_npos = 0, s3_nslot = 0, s3_nrec = 0;
  long long s3_slots_epoch = -1, s3_opts_epoch = -1;
  bool s3_ok = false;
  double* s3_V = nullptr; double* s3_Mv = nullptr; double* s3_bs = nullptr; double* s3_xs = nullptr;
  int* s3_flag = nullptr;
  size_t s3_V_cap = 0, s3_Mv_cap = 0, s3_scr_cap = 0, s3_flag_cap = 0;
  // the schedule COMPILED (biglin3_gen.hpp): module, factor-value count, buffers
  void* c3_mod[2] = {nullptr, nullptr};   // jit::Kernel objects (factor, solve), owned here as opaque pointers
  int c3_state = 0;                       // 0 not tried for this epoch, 1 ready, -1 unavailable
  int last_path = -1;                     // ahk_big_stats: 0 dense workspace, 1 sparse dynamic, 2 static schedule, 3 compiled schedule
  int c3_nf = 0, c3_npos = 0;
  double* c3_F = nullptr; double* c3_Mv = nullptr; int* c3_flagm = nullptr; double* c3_scr = nullptr;
  size_t c3_F_cap = 0, c3_Mv_cap = 0, c3_flag_cap = 0, c3_scr_cap = 0;
  void release_c3();
  ahk::Big3Sched* s3_S = nullptr;         // host copy of the schedule (source of the generated code), owned
  void release3();
  void release() {
    release3();
    cudaFree(wsA); cudaFree(wsLv); cudaFree(wsUv); cudaFree(wsLi); cudaFree(wsUi); cudaFree(d_sweep);
    cudaFree(s2Lv); cudaFree(s2Uv); cudaFree(s2Li); cudaFree(s2Ui); cudaFree(d_ovf);
    wsA = wsLv = wsUv = nullptr; wsLi = wsUi = nullptr; d_sweep = nullptr; nwarps = 0;
    s2Lv = s2Uv = nullptr; s2Li = s2Ui = nullptr; d_ovf = nullptr; nwarps2 = 0; ovf_cap = 0;
  }
  int run(ahk_engine* e, const ahk_batch* batch, int src_elem, const double* sweep_host, int npts,
          double* x_out, int32_t* iters, int32_t* status, double* resid, cudaStream_t s,
          std::string& err, std::atomic<long long>& launches);
};
#endif
