// ac.cuh — batched small-signal AC: per instance and per angular frequency one
// complex128 solve of (MNA + Gmin + J(x_op) + j*omega*AC) x = -Nac.
//
// Replaces ac.ac_analysis and its helpers (ahkab/ac.py:146-443).  The reference
// routes every frequency through dc_solve/mdn_solver with a dummy *linear*
// circuit (ac.py:293-305), i.e. exactly one iteration x = x_prev + solve(A,
// -(A x_prev + Nac)); the same update is done here, frequency after frequency,
// by the group that owns the instance (x_prev chains inside a frequency chunk).
// Complex solve: partial pivoting on |re|+|im| as LAPACK zgesv (izamax) does;
// register-resident Gauss-Jordan for Var<NC,R,G> (see core.cuh), LU in shared
// memory for the generic fallback Var<0,0,G>.
#pragma once
#include "core.cuh"
#ifndef AHK_JIT   // (specialised build: AcLayout is a set of constants from jit_spec.h)
#include "program_build.hpp"
#endif

namespace ahk {

#ifdef __CUDACC__
struct __align__(16) cd { double re, im; };
#else
struct alignas(16) cd { double re, im; };
#endif
AHK_DEV cd cmul(cd a, cd b) { return cd{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
AHK_DEV cd csub(cd a, cd b) { return cd{a.re - b.re, a.im - b.im}; }
AHK_DEV cd cadd(cd a, cd b) { return cd{a.re + b.re, a.im + b.im}; }
// reciprocal through the squared modulus: one division (pivots of circuit matrices are
// far from the over/underflow range of |b|^2)
AHK_DEV cd crecip(cd b) {
  // (branch-free reciprocal: an IEEE division ends in a conditional call to its slow path that
  // the compiler schedules nothing across — 14 of them per solve in the static-schedule kernel)
  const double s = fm::rcp(b.re * b.re + b.im * b.im);
  return cd{b.re * s, -b.im * s};
}
AHK_DEV cd cdivv(cd a, cd b) {   // Smith's algorithm (what C / numpy use)
  if (fabs(b.re) >= fabs(b.im)) {
    double r = b.im / b.re, den = b.re + b.im * r;
    return cd{(a.re + a.im * r) / den, (a.im - a.re * r) / den};
  } else {
    double r = b.re / b.im, den = b.re * r + b.im;
    return cd{(a.re * r + a.im) / den, (a.im * r - a.re) / den};
  }
}

#ifndef AHK_JIT
struct AcLayout {
  Layout L;          // real arrays shared with core.cuh (Mv, Dv, Bv, dout, dst, x=xr, bytes, Pv)
  int RS;            // complex row stride (elements) of the staging / fallback matrix
  int Ac, xc, dxc, Nac, pvc;   // offsets in doubles of the complex arrays
  int stride;
};

// big: n beyond AHK_SMALL_NMAX — the complex matrix lives in the global workspace (Ctx::Ag,
// n x RS complex elements), its right-hand side column in shared memory (L.brhs)
// nostage: the lane-per-chunk static-schedule kernel — rows are built from Bv / Dv on literal
// indices (no staging rows), every lane of the group has its own x / dx (its own frequencies)
inline AcLayout make_ac_layout(const HostProg& H, int G, int NC = 0, bool big = false, bool nostage = false) {
  AcLayout A; std::memset(&A, 0, sizeof A);
  const Prog& h = H.hdr;
  const int n = h.n;
  A.RS = NC > 0 ? NC : (big ? ((n + 1) & ~1) : n + 1);   // big: rows start on 32-byte sectors
  int off = 0;
  auto take = [&](int cnt) { int o = off; off += cnt > 0 ? cnt : 0; return o; };
  auto take2 = [&](int cnt) { off += off & 1; return take(cnt); };   // 16-byte aligned
  A.Ac = take2((big || nostage) ? 0 : 2 * n * A.RS);   // fallback: complex matrix; register variants: (Bv, Dv) staging rows
  A.L.big = big ? 1 : 0;
  A.L.brhs = take2(big ? 2 * n : 0);
  A.L.nzm = take(big ? (n * ((n + 31) / 32) * 4 + 7) / 8 : 0);
  A.xc = take2(2 * n * (nostage ? G : 1)); A.dxc = take2(2 * n * (nostage ? G : 1)); A.Nac = take2(2 * n);
  A.pvc = take2(NC > 0 && G > 1 ? 4 * NC : 0);
  A.L.RS = A.RS;
  // (many matrix positions — dense patterns — do not fit next to the vectors: global workspace)
  A.L.vglob = big && h.npos > AHK_VGLOB_NPOS ? 1 : 0;
  if (A.L.vglob) { A.L.Bv = 0; A.L.Mv = h.npos; A.L.Dv = 2 * h.npos; A.L.vlen = (3 * h.npos + 3) & ~3; }
  else { A.L.Bv = take(h.npos); A.L.Mv = take(h.npos); A.L.Dv = take(h.npos); }
  A.L.x = take(n);
  A.L.dout = take(h.ndout); A.L.dst = take(H.nst);
  A.L.bytes = take(big ? (6 * n + 7) / 8 + 1 : 16);
  A.L.Pv = take(h.npersist); A.L.dcon = take(H.ncon);
  A.L.buflen = 1;
  // complex elements are 16-byte aligned: even stride, 2 mod 4 to spread the groups
  while (off % 4 != 2) off++;
  A.stride = off;
  A.L.stride = off;
  return A;
}

#endif  // AHK_JIT

struct AcArgs {
  const double* x_op;      // [B][n] or null
  const double* omegas;    // device [npts]
  int npts;
  int f0, f1;              // host emulation: the chunk; kernels compute their own
  double* x_out;           // [B][npts][n] complex
  int* status;
};

// kernel parameters of the circuit-specialised AC kernel (jit_kernels.cuh; filled by engine.cu)
struct KAcJ { const double* vary; long long B; AcArgs a; int nchunks; int* pivrec; };

// generic fallback: complex LU + back-substitution on the shared-memory matrix
template <int G>
AHK_DEV int zlu_solve_smem(Ctx& c, const AcLayout& AL) {
  const int n = c.p->n, RS = AL.RS;
  cd* A = (cd*)(c.sm + AL.Ac);
  cd* dx = (cd*)(c.sm + AL.dxc);
  unsigned char* prow = c.prow();
  unsigned char* stepof = c.stepof();
  unsigned long long used = 0ull;
  for (int k = 0; k < n; k++) {
    double best = -1.0;
    int bi = -1;
    for (int r = c.sub; r < n; r += G) {
      if ((used >> r) & 1ull) continue;
      cd a = A[r * RS + k];
      double v = fabs(a.re) + fabs(a.im);
      if (bi < 0 || v > best) { best = v; bi = r; }
    }
    Grp<G>::argmax(best, bi, c.mask);
    if (best == 0.0) return 0;
    used |= 1ull << bi;
    if (c.sub == 0) { prow[k] = (unsigned char)bi; stepof[bi] = (unsigned char)k; }
    const cd* Ap = A + bi * RS;
    cd one = {1.0, 0.0};
    cd inv = cdivv(one, Ap[k]);
    for (int r = c.sub; r < n; r += G) {
      if ((used >> r) & 1ull) continue;
      cd* Ar = A + r * RS;
      cd l = cmul(Ar[k], inv);
      if (l.re == 0.0 && l.im == 0.0) continue;
      for (int j = k + 1; j <= n; j++) Ar[j] = csub(Ar[j], cmul(l, Ap[j]));
    }
    Grp<G>::sync(c.mask);
  }
  for (int k = n - 1; k >= 0; k--) {
    int pr = prow[k];
    cd xk = cdivv(A[pr * RS + n], A[pr * RS + k]);
    if (c.sub == 0) dx[k] = xk;
    for (int r = c.sub; r < n; r += G)
      if ((int)stepof[r] < k) A[r * RS + n] = csub(A[r * RS + n], cmul(A[r * RS + k], xk));
    Grp<G>::sync(c.mask);
  }
  return 1;
}

// large n: the complex matrix in global memory — the scheme of lu_solve_glob (core.cuh): per-row
// bitmaps of the non-zero columns; column k is read only where its bit is set, a row update
// touches only the columns of the pivot row's bitmap
template <int G>
AHK_DEV int zlu_solve_glob(Ctx& c, const AcLayout& AL) {
  const int n = c.p->n, RS = AL.RS, W = (n + 31) >> 5;
  cd* A = (cd*)c.Ag;
  cd* dx = (cd*)(c.sm + AL.dxc);        // column k / multipliers while eliminating, then the solution
  cd* b = (cd*)(c.sm + AL.L.brhs);
  unsigned* nzm = (unsigned*)(c.sm + AL.L.nzm);
  unsigned short* prow = (unsigned short*)(c.sm + AL.L.bytes);
  unsigned short* stepof = prow + n;
  unsigned short* list = stepof + n;
  for (int r = c.sub; r < n; r += G) stepof[r] = 0xffff;
  Grp<G>::sync(c.mask);
  for (int k = 0; k < n; k++) {
    const int kw = k >> 5;
    const unsigned kb = 1u << (k & 31);
    double best = -1.0;
    int bi = -1;
    for (int r = c.sub; r < n; r += G) {
      if (stepof[r] != 0xffff) continue;
      const cd a = (nzm[r * W + kw] & kb) ? A[(size_t)r * RS + k] : cd{0.0, 0.0};
      dx[r] = a;
      const double v = fabs(a.re) + fabs(a.im);
      if (bi < 0 || v > best) { best = v; bi = r; }
    }
    Grp<G>::argmax(best, bi, c.mask);
    if (best == 0.0) return 0;
    if (c.sub == 0) { prow[k] = (unsigned short)bi; stepof[bi] = (unsigned short)k; }
    Grp<G>::sync(c.mask);
    const cd one = {1.0, 0.0};
    const cd inv = cdivv(one, dx[bi]);
    int cnt = 0;
    for (int r0 = 0; r0 < n; r0 += G) {
      const int r = r0 + c.sub;
      bool nz = false;
      if (r < n && stepof[r] == 0xffff) {
        const cd l = cmul(dx[r], inv);
        nz = !(l.re == 0.0 && l.im == 0.0);
        if (nz) dx[r] = l;
      }
#ifdef __CUDACC__
      if (G == 32) {
        const unsigned bal = __ballot_sync(0xffffffffu, nz);
        if (nz) list[cnt + __popc(bal & ((1u << c.sub) - 1u))] = (unsigned short)r;
        cnt += __popc(bal);
      } else
#endif
      { if (nz) list[cnt++] = (unsigned short)r; }
    }
    Grp<G>::sync(c.mask);
    const cd* Ap = A + (size_t)bi * RS;
    const unsigned* nzp = nzm + bi * W;
    const cd bp = b[bi];
    const int w0 = (k + 1) >> 5, Wrem = W - w0;
    const unsigned m0 = ~0u << ((k + 1) & 31);
    int nnzp = 0;
    for (int w = w0; w < W; w++) nnzp += big_popc(nzp[w] & (w == w0 ? m0 : ~0u));
    if (k + 1 < n && cnt > 0) {
      if (nnzp * 4 > n - k - 1) {
        for (int t = 0; t < cnt; t++) {
          const int r = list[t];
          const cd l = dx[r];
          cd* Ar = A + (size_t)r * RS;
#pragma unroll 2
          for (int j = k + 1 + c.sub; j < n; j += G) Ar[j] = csub(Ar[j], cmul(l, Ap[j]));
          for (int w = w0 + c.sub; w < W; w += G) nzm[r * W + w] |= nzp[w] & (w == w0 ? m0 : ~0u);
        }
      } else {
        for (int item = c.sub; item < cnt * Wrem; item += G) {
          const int t = item / Wrem, w = w0 + item % Wrem;
          unsigned bits = nzp[w] & (w == w0 ? m0 : ~0u);
          if (!bits) continue;
          const int r = list[t];
          const cd l = dx[r];
          cd* Ar = A + (size_t)r * RS;
          nzm[r * W + w] |= bits;
          while (bits) {
            const int j = (w << 5) + big_ctz(bits);
            bits &= bits - 1;
            Ar[j] = csub(Ar[j], cmul(l, Ap[j]));
          }
        }
      }
    }
    for (int t = c.sub; t < cnt; t += G) { const int r = list[t]; b[r] = csub(b[r], cmul(dx[r], bp)); }
    Grp<G>::sync(c.mask);
  }
  for (int k = n - 1; k >= 0; k--) {
    const int pr = prow[k];
    const cd* Ap = A + (size_t)pr * RS;
    const unsigned* nzp = nzm + pr * W;
    const int w0 = (k + 1) >> 5;
    const unsigned m0 = ~0u << ((k + 1) & 31);
    cd s = {0.0, 0.0};
    if (k + 1 < n) {
      for (int w = w0 + c.sub; w < W; w += G) {
        unsigned bits = nzp[w] & (w == w0 ? m0 : ~0u);
        while (bits) {
          const int j = (w << 5) + big_ctz(bits);
          bits &= bits - 1;
          s = cadd(s, cmul(Ap[j], dx[j]));
        }
      }
    }
    s.re = Grp<G>::sum_all(s.re, c.mask);
    s.im = Grp<G>::sum_all(s.im, c.mask);
    if (c.sub == 0) dx[k] = cdivv(csub(b[pr], s), Ap[k]);
    Grp<G>::sync(c.mask);
  }
  return 1;
}

// register-resident complex Gauss-Jordan (same scheme as gj_solve in core.cuh)
template <class V>
AHK_DEV int zgj_solve(Ctx& c, const AcLayout& AL, cd (&a)[V::R > 0 ? V::R : 1][V::NC > 0 ? V::NC : 1]) {
  const int NC = V::NC, R = V::R, G = V::G;
  const int n = c.p->n;
  cd* dx = (cd*)(c.sm + AL.dxc);
  cd* pvb = (cd*)(c.sm + AL.pvc);   // [2][NC]
  unsigned usedm = 0;
  cd pd[R > 0 ? R : 1];
  int ps[R > 0 ? R : 1];
#pragma unroll
  for (int s = 0; s < R; s++) { pd[s] = cd{1.0, 0.0}; ps[s] = 0; }   // pd: RECIPROCAL of the row's pivot
#pragma unroll
  for (int k = 0; k < NC - 1; k++) {
    if (k < n) {
      unsigned long long key = 0ull;
#pragma unroll
      for (int s = 0; s < R; s++) {
        const int r = c.sub + G * s;
        const unsigned long long kb = pivot_key(fabs(a[s][k].re) + fabs(a[s][k].im), r);
        const bool cand = r < n && !((usedm >> s) & 1u);
        if (cand && kb > key) key = kb;
      }
      key = Grp<G>::kmax(key, c.mask);
      if ((key >> 5) == 0ull) return 0;
      const int bi = 31 - (int)(key & 31ull);
      const int own = bi & (G - 1), slot = bi / G;
      const bool mine = c.sub == own;
#ifdef AHK_BCAST_SHFL
      if (G > 1 && R == 1) {
        // one row per lane: every element of the pivot row is fetched from the owner's
        // registers right where it is used (4 shuffles per complex element) — no pivot-row
        // copy in registers, no shared-memory round trip, no barrier
        const cd pk = cd{Grp<G>::bcast(a[0][k].re, own, c.mask), Grp<G>::bcast(a[0][k].im, own, c.mask)};
        const cd inv = crecip(pk);
        if (mine) { usedm |= 1u; pd[0] = inv; ps[0] = k; }
        cd l = cmul(a[0][k], inv);
        if (mine) l = cd{0.0, 0.0};
#pragma unroll
        for (int j = k + 1; j < NC; j++) {
          const double pre = Grp<G>::bcast(a[0][j].re, own, c.mask);
          const double pim = Grp<G>::bcast(a[0][j].im, own, c.mask);
          a[0][j].re -= l.re * pre - l.im * pim;
          a[0][j].im -= l.re * pim + l.im * pre;
        }
#else
      if (G > 1 && R == 1) {
        // one row per lane: the owner publishes its row in one of two shared-memory buffers
        // (16-byte stores), one group barrier, then every lane reads each element right
        // where it is used (one 16-byte broadcast load per complex element instead of four
        // shuffles; no pivot-row copy in registers).  Two buffers: a lane can only reach the
        // barrier of step k+1 after its reads of step k, so buffer k & 1 is free again at k+2.
        cd* buf = pvb + (k & 1) * NC;
        if (mine) {
#pragma unroll
          for (int j = k; j < NC; j++) buf[j] = a[0][j];
        }
        Grp<G>::sync(c.mask);
        const cd pk = buf[k];
        const cd inv = crecip(pk);
        if (mine) { usedm |= 1u; pd[0] = inv; ps[0] = k; }
        cd l = cmul(a[0][k], inv);
        if (mine) l = cd{0.0, 0.0};
#pragma unroll
        for (int j = k + 1; j < NC; j++) {
          const cd pv = buf[j];
          a[0][j].re -= l.re * pv.re - l.im * pv.im;
          a[0][j].im -= l.re * pv.im + l.im * pv.re;
        }
#endif
      } else {
        cd pv[NC > 0 ? NC : 1];
        if (G > 1) {
          cd* buf = pvb + (k & 1) * NC;
          if (mine) {
#pragma unroll
            for (int s = 0; s < R; s++)
              if (slot == s) {
#pragma unroll
                for (int j = k; j < NC; j++) buf[j] = a[s][j];
              }
          }
          Grp<G>::sync(c.mask);
#pragma unroll
          for (int j = k; j < NC; j++) pv[j] = buf[j];
        } else {
#pragma unroll
          for (int j = k; j < NC; j++) {
            cd t = a[0][j];
#pragma unroll
            for (int s = 1; s < R; s++) if (slot == s) t = a[s][j];
            pv[j] = t;
          }
        }
        const cd inv = crecip(pv[k]);
        if (mine) {
          usedm |= 1u << slot;
#pragma unroll
          for (int s = 0; s < R; s++) if (slot == s) { pd[s] = inv; ps[s] = k; }
        }
#pragma unroll
        for (int s = 0; s < R; s++) {
          cd l = cmul(a[s][k], inv);
          if (mine && slot == s) l = cd{0.0, 0.0};
#pragma unroll
          for (int j = k + 1; j < NC; j++) {
            a[s][j].re -= l.re * pv[j].re - l.im * pv[j].im;
            a[s][j].im -= l.re * pv[j].im + l.im * pv[j].re;
          }
        }
      }
    }
  }
#pragma unroll
  for (int s = 0; s < R; s++) {
    const int r = c.sub + G * s;
    if (r < n) dx[ps[s]] = cmul(a[s][NC - 1], pd[s]);   // (stored pivot reciprocal: no division)
  }
  Grp<G>::sync(c.mask);
  return 1;
}

#ifdef AHK_JIT_SLU
#ifdef __CUDACC__
// (an empty asm statement orders nothing for ptxas; a real shared-memory store that the row loads
// behind it might alias does: dx[0] is rewritten by the back-substitution at the end of the body)
#define AHK_SLU_FENCE { ((volatile double*)dx)[0] = 0.0; }
#else
#define AHK_SLU_FENCE
#endif
// ---- one lane per (instance, frequency chunk): zgesv's LU along a static schedule ------------
// The G lanes of a group share their instance's set-up (parameters, Bv, Dv, Nac in the group's
// shared-memory slice) and then each walks its own chunk of the sweep with its own x / dx.
// The complex form of core.cuh: slu_solve_thread — AHK_JIT_SLU_BODY is generated by
// jit_gen.hpp (slu_codegen, cx) on the entries m[r][c]: elimination and back-substitution on
// the structural non-zeros, literal indices, every pivot verified against zgetf2's rule
// (first maximum of |re| + |im| among the rows not yet used).
template <class V>
AHK_DEV int zslu_solve_thread(const double* Bv, const double* Dv, const cd* Nac, const double om, const cd* x,
                              cd* dx) {
  // the rows are assembled inside the body (jit_gen.hpp: SluEmit::need_row), each right before the
  // first pivot step that reads it; AHK_SLU_FENCE keeps the compiler from hoisting all of those
  // loads to the top again (the matrix with its fill-in does not fit in 255 registers)
  cd m[V::R > 0 ? V::R : 1][V::NC > 0 ? V::NC : 1];
#pragma unroll
  for (int s = 0; s < V::R; s++) {
#pragma unroll
    for (int j = 0; j < V::NC; j++) m[s][j] = cd{0.0, 0.0};
  }
  bool same = true;
  AHK_JIT_SLU_BODY
  return same ? 1 : 0;
}

// A solve that left the schedule (or is singular): zgetf2 / zgetrs as LAPACK runs them — pivot
// search, row swaps — on a dense local matrix, with the arithmetic of the scheduled form
// (reciprocal pivot, multiply).  Not inlined; rec receives the pivot positions.
template <class V>
#ifdef __CUDACC__
__device__ __noinline__
#else
static
#endif
int zslu_fallback(Ctx& c, const AcLayout& AL, double om, const cd* x, cd* dx, int* rec) {
  const int NC = V::NC, n = c.p->n;
  const Prog& p = *c.p;
  const double* Bv = c.vb() + AL.L.Bv;
  const double* Dv = c.vb() + AL.L.Dv;
  const cd* Nac = (const cd*)(c.sm + AL.Nac);
  cd m[V::R > 0 ? V::R : 1][NC > 0 ? NC : 1];
  cd inv[V::R > 0 ? V::R : 1];
  for (int r = 0; r < n; r++) {
    for (int j = 0; j < NC; j++) m[r][j] = cd{0.0, 0.0};
    cd acc = Nac[r];
    for (int k = p.row_ptr[r]; k < p.row_ptr[r + 1]; k++) {
      const cd v = {Bv[k], om * Dv[k]};
      m[r][p.pcol[k]] = v;
      acc = cadd(acc, cmul(v, x[p.pcol[k]]));
    }
    m[r][NC - 1] = cd{-acc.re, -acc.im};
  }
  for (int k = 0; k < n; k++) {
    int pv = k; double best = fabs(m[k][k].re) + fabs(m[k][k].im);
    for (int r = k + 1; r < n; r++) {
      const double v = fabs(m[r][k].re) + fabs(m[r][k].im);
      if (v > best) { best = v; pv = r; }
    }
    if (best == 0.0) return 0;
    rec[k] = pv;
    if (pv != k) for (int j = k; j < NC; j++) { const cd t = m[k][j]; m[k][j] = m[pv][j]; m[pv][j] = t; }
    inv[k] = crecip(m[k][k]);
    for (int r = k + 1; r < n; r++) {
      if (m[r][k].re == 0.0 && m[r][k].im == 0.0) continue;   // (exact zero: nothing to eliminate)
      const cd l = cmul(m[r][k], inv[k]);
      for (int j = k + 1; j < NC; j++) m[r][j] = csub(m[r][j], cmul(l, m[k][j]));
    }
  }
  for (int k = n - 1; k >= 0; k--) {
    cd s = m[k][NC - 1];
    for (int j = k + 1; j < n; j++) s = csub(s, cmul(m[k][j], dx[j]));
    dx[k] = cmul(s, inv[k]);
  }
  return 1;
}
#endif

template <class V>
AHK_DEV void run_ac(Ctx& c, const AcLayout& AL, const AcArgs& a) {
  const int G = V::G;
  const Prog& p = *c.p;
  const Opts& o = *c.o;
  const int n = p.n, RS = AL.RS;
  c.L = &AL.L;
  c.singular = 0; c.dc_src = -1; c.dc_val = 0.0; c.C1 = 0.0; c.gadd = 0.0; c.G = G;
  if constexpr (V::NC == 0) {
    if (AL.L.big) big_init<G>(c, c.Ag, (size_t)2 * n * RS);   // large n: complex matrix zero, no entry flagged
  }
  build_values<Var<0, 0, V::G> >(c, true);   // parameters, Mv, Dv, device constants
  double* Bv = c.vb() + AL.L.Bv;
  const double* Mv = c.vb() + AL.L.Mv;
  const double* Dv = c.vb() + AL.L.Dv;
  double* xr = c.sm + AL.L.x;
  cd* x = (cd*)(c.sm + AL.xc);
  cd* dx = (cd*)(c.sm + AL.dxc);
  cd* Nac = (cd*)(c.sm + AL.Nac);
  const bool have_op = a.x_op != nullptr;
  for (int i = c.sub; i < n; i += G) {
    xr[i] = have_op ? a.x_op[c.inst * n + i] : 0.0;
    x[i] = cd{(p.ndev && have_op) ? xr[i] : 0.0, 0.0};   // x = x0 (ac.py:290)
    Nac[i] = cd{0.0, 0.0};
  }
  Grp<G>::sync(c.mask);
  if (p.ndev) { eval_devices(c, xr); Grp<G>::sync(c.mask); }   // J at the linearisation point (ac.py:417-443)
  const double* dout = c.sm + AL.L.dout;
  for (int k = c.sub; k < p.npos; k += G) {
    double v = Mv[k];
    for (int q = p.nl_ptr[k]; q < p.nl_ptr[k + 1]; q++) v += (double)p.nl_sgn[q] * dout[p.nl_src[q]];
    if (p.pdiag[k]) v += o.gmin;
    Bv[k] = v;
  }
  if (c.sub == 0) {   // _generate_Nac, ac.py:380-414
    for (int k = 0; k < p.nsrc; k++) {
      const SrcOp& s = p.src[k];
      if (!s.has_ac) continue;
      double mag = c.P(s.pbase + 1), ph = c.P(s.pbase + 2);   // from HBM, once per instance
      cd v = {mag * cos(ph), mag * sin(ph)};
      if (s.is_v) Nac[s.row0] = cd{-1.0 * v.re, -1.0 * v.im};
      else {
        if (s.row0 >= 0) Nac[s.row0] = cadd(Nac[s.row0], v);
        if (s.row1 >= 0) Nac[s.row1] = csub(Nac[s.row1], v);
      }
    }
  }
  Grp<G>::sync(c.mask);
  // register variants: dense staging rows of (Bv, Dv) pairs, filled once per instance;
  // every frequency then builds its complex row with NC loads and no scatter
  double* BD = c.sm + AL.Ac;
#ifdef AHK_JIT_SLU
  constexpr bool SLU1 = V::NC > 0;   // lane per chunk; rows built from Bv / Dv on literal indices: no staging
#else
  constexpr bool SLU1 = false;
#endif
  if constexpr (V::NC > 0 && !SLU1) {
    for (int i = c.sub; i < 2 * n * V::NC; i += G) BD[i] = 0.0;
    Grp<G>::sync(c.mask);
    for (int k = c.sub; k < p.npos; k += G) {
      double* e = BD + 2 * ((int)p.prowi[k] * V::NC + p.pcol[k]);
      e[0] = Bv[k]; e[1] = Dv[k];
    }
    Grp<G>::sync(c.mask);
  }
  int st = AHK_ST_OK;
  int f0 = a.f0, f1 = a.f1;
  if constexpr (SLU1) {       // this lane's chunk of the sweep, its own iterate
    const int per = (a.npts + G - 1) / G;
    f0 = c.sub * per; f1 = f0 + per < a.npts ? f0 + per : a.npts;
    Grp<G>::sync(c.mask);
    cd* xs = x;
    x += c.sub * n; dx += c.sub * n;
    if (c.sub) for (int i = 0; i < n; i++) x[i] = xs[i];
  }
#if defined(AHK_JIT_SLU) && defined(__CUDACC__)
  if constexpr (SLU1) {
    // Every lane of the group runs the same number of rounds (one frequency of its chunk per
    // round), and the result rows of a round leave through the WHOLE group: the 2 n doubles of a
    // row by 2 n consecutive lanes — 224 contiguous bytes for C1 instead of 28 eight-byte pieces
    // 224 bytes apart from lane to lane (what a zero-copy write into pinned host memory needs,
    // and kinder to L2 for device buffers).  A lane whose matrix is singular fills its row with
    // NaN and keeps writing it for the rest of its chunk, as the one-lane form did.
    const int per = (a.npts + G - 1) / G;
    const cd* xs0 = (const cd*)(c.sm + AL.xc);          // lane l's iterate: xs0 + l * n
    bool dead = false;
    for (int j = 0; j < per; j++) {
      const int f = f0 + j;
      if (f < f1 && !dead) {
        const int R = V::R;
        const double om = a.omegas[f];
        int ok = zslu_solve_thread<V>(Bv, Dv, Nac, om, x, dx);
        // calibration launch: every lane of instance 0 keeps its own record (engine.cu merges them)
        int* prec = (c.pivrec != nullptr && c.inst == 0) ? c.pivrec + 256 * c.sub : nullptr;
        if (prec) prec[1]++;
        if (!ok) {
          int rec[R > 0 ? R : 1];
          ok = zslu_fallback<V>(c, AL, om, x, dx, rec);
          if (prec && ok) slu_record(prec, rec, n);
        }
        if (ok) {
          for (int i = 0; i < n; i++) x[i] = cadd(x[i], dx[i]);
        } else {
          dead = true; st = AHK_ST_SINGULAR;
          const double qn = nan("");
          for (int i = 0; i < n; i++) x[i] = cd{qn, qn};
        }
      }
      Grp<G>::sync(c.mask);
      for (int l = 0; l < G; l++) {
        const int fl = l * per + j;                       // lane l's frequency of this round
        if (fl >= a.npts) break;
        const double* src = (const double*)(xs0 + l * n);
        double* dst = a.x_out + ((c.inst * a.npts + fl) * n) * 2;
        for (int i = c.sub; i < 2 * n; i += G) dst[i] = src[i];
      }
      Grp<G>::sync(c.mask);
    }
    // status[] is pre-filled with AHK_ST_OK by the caller; lanes only downgrade it
    if (a.status && st != AHK_ST_OK) a.status[c.inst] = st;
    return;
  }
#endif
  for (int f = f0; f < f1; f++) {
    const double om = a.omegas[f];
    int ok;
#if defined(AHK_JIT_SLU) && !defined(__CUDACC__)
    if constexpr (SLU1) {
      const int R = V::R;
      ok = zslu_solve_thread<V>(Bv, Dv, Nac, om, x, dx);
      // calibration launch: every lane of instance 0 keeps its own record (engine.cu merges them)
      int* prec = (c.pivrec != nullptr && c.inst == 0) ? c.pivrec + 256 * c.sub : nullptr;
      if (prec) prec[1]++;
      if (!ok) {
        int rec[R > 0 ? R : 1];
        ok = zslu_fallback<V>(c, AL, om, x, dx, rec);
        if (prec && ok) slu_record(prec, rec, n);
      }
    } else
#endif
    if constexpr (V::NC > 0) {
      const int NC = V::NC, R = V::R;
      cd m[R > 0 ? R : 1][NC > 0 ? NC : 1];
#pragma unroll
      for (int s = 0; s < R; s++) {
        const int r = c.sub + G * s;
        if (r < n) {
          const double* row = BD + 2 * r * NC;
#pragma unroll
          for (int j = 0; j < NC - 1; j++) m[s][j] = cd{row[2 * j], om * row[2 * j + 1]};
          cd acc = Nac[r];
          for (int k = p.row_ptr[r]; k < p.row_ptr[r + 1]; k++) {
            const cd v = {Bv[k], om * Dv[k]};
            acc = cadd(acc, cmul(v, x[p.pcol[k]]));
          }
          m[s][NC - 1] = cd{-acc.re, -acc.im};
        } else {
#pragma unroll
          for (int j = 0; j < NC; j++) m[s][j] = cd{0.0, 0.0};
        }
      }
      ok = zgj_solve<V>(c, AL, m);
    } else if (AL.L.big) {     // large n: matrix in the global workspace
      cd* A = (cd*)c.Ag;
      cd* b = (cd*)(c.sm + AL.L.brhs);
      const int W = (n + 31) >> 5;
      unsigned* nzm = (unsigned*)(c.sm + AL.L.nzm);
      for (int item = c.sub; item < n * W; item += G) {     // clear the flagged entries of the last factorisation
        unsigned bits = nzm[item];
        if (!bits) continue;
        nzm[item] = 0u;
        cd* Ar = A + (size_t)(item / W) * RS + ((item % W) << 5);
        while (bits) { Ar[big_ctz(bits)] = cd{0.0, 0.0}; bits &= bits - 1; }
      }
      Grp<G>::sync(c.mask);
      for (int r = c.sub; r < n; r += G) {
        cd* Ar = A + (size_t)r * RS;
        cd s = Nac[r];
        for (int k = p.row_ptr[r]; k < p.row_ptr[r + 1]; k++) {
          int col = p.pcol[k];
          cd v = {Bv[k], om * Dv[k]};
          Ar[col] = v;
          nzm[r * W + (col >> 5)] |= 1u << (col & 31);
          s = cadd(s, cmul(v, x[col]));
        }
        b[r] = cd{-s.re, -s.im};
      }
      Grp<G>::sync(c.mask);
      ok = zlu_solve_glob<G>(c, AL);
    } else {
      cd* A = (cd*)(c.sm + AL.Ac);
      for (int r = c.sub; r < n; r += G) {
        cd* Ar = A + r * RS;
        for (int j = 0; j < n; j++) Ar[j] = cd{0.0, 0.0};
        cd s = Nac[r];
        for (int k = p.row_ptr[r]; k < p.row_ptr[r + 1]; k++) {
          int col = p.pcol[k];
          cd v = {Bv[k], om * Dv[k]};
          Ar[col] = v;
          s = cadd(s, cmul(v, x[col]));
        }
        Ar[n] = cd{-s.re, -s.im};
      }
      Grp<G>::sync(c.mask);
      ok = zlu_solve_smem<G>(c, AL);
    }
    double* xo = a.x_out + ((c.inst * a.npts + f) * n) * 2;
    const int i0 = SLU1 ? 0 : c.sub, di = SLU1 ? 1 : G;     // lane per chunk: the whole vector is this lane's
    if (!ok) {
      st = AHK_ST_SINGULAR;
      const double qn = nan("");
      for (int g = f; g < f1; g++) {
        double* xg = a.x_out + ((c.inst * a.npts + g) * n) * 2;
        for (int i = i0; i < n; i += di) { xg[2 * i] = qn; xg[2 * i + 1] = qn; }
      }
      break;
    }
    for (int i = i0; i < n; i += di) {
      cd v = cadd(x[i], dx[i]);
      x[i] = v;
      xo[2 * i] = v.re; xo[2 * i + 1] = v.im;
    }
    if constexpr (!SLU1) Grp<G>::sync(c.mask);
  }
  // status[] is pre-filled with AHK_ST_OK by the caller; chunks only downgrade it
  if ((SLU1 || c.sub == 0) && a.status && st != AHK_ST_OK) a.status[c.inst] = st;
}

}  // namespace ahk
