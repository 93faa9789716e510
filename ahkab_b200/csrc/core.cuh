// core.cuh — per-instance Newton/MNA machinery of the small-circuit engine.
//
// Execution model: a GROUP of G lanes (G = 1,2,4,8,16,32, a sub-warp) owns one
// circuit instance for the whole analysis.  The instance's dense augmented
// matrix [n][RS], its iterate, history ring and device state live in shared
// memory; matrix rows are owned cyclically by the lanes of the group
// (row r -> lane r % G).  Devices are evaluated one per lane, the Jacobian is
// assembled by a conflict-free gather, the LU uses implicit partial pivoting
// (no row swaps) with a group arg-max.  Control scalars (time, step, counters)
// are kept redundantly in registers and are group-uniform by construction.
//
// What it replaces in the reference (all per instance, pure Python there):
//   dc_analysis.mdn_solver / _update_J_and_Tx / get_td   dc_analysis.py:690-946
//   utilities.convergence_check                          utilities.py:392-549
//   dc_analysis.dc_solve (solve-method state machine)    dc_analysis.py:128-320
//   dc_analysis.op_analysis / dc_analysis                dc_analysis.py:461-687
//   transient.transient_analysis + IE/TRAP/GEAR get_df   transient.py:121-470,
//       implicit_euler.py:125-205, trap.py:98-192, gear.py:122-208
#pragma once
#include "devices.cuh"
#include "program.h"

namespace ahk {

// ---- group primitives ---------------------------------------------------------
#ifdef __CUDACC__
template <int G>
struct Grp {
  static AHK_DEV void sync(unsigned m) { if (G > 1) __syncwarp(m); }
  static AHK_DEV double fmin_all(double v, unsigned m) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(m, v, o));
    return v;
  }
  static AHK_DEV int all(int pred, unsigned m) { return G > 1 ? __all_sync(m, pred) : pred; }
  static AHK_DEV int any(int pred, unsigned m) { return G > 1 ? __any_sync(m, pred) : pred; }
  // arg-max of (val, idx): larger val wins, ties -> smaller idx; idx < 0 = no candidate
  static AHK_DEV void argmax(double& v, int& i, unsigned m) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      double ov = __shfl_xor_sync(m, v, o);
      int oi = __shfl_xor_sync(m, i, o);
      bool take = (oi >= 0) && (i < 0 || ov > v || (ov == v && oi < i));
      if (take) { v = ov; i = oi; }
    }
  }
};
#else
template <int G>
struct Grp {  // host emulation (tests/hostemu): G == 1 only
  static void sync(unsigned) {}
  static double fmin_all(double v, unsigned) { return v; }
  static int all(int p, unsigned) { return p; }
  static int any(int p, unsigned) { return p; }
  static void argmax(double&, int&, unsigned) {}
};
#endif

struct Ctx {
  const Prog* p;
  const Opts* o;
  const Layout* L;
  double* sm;            // this instance's shared-memory slice
  const double* vary;    // [n_vary][B]
  long long B, inst;
  int sub;               // lane within the group
  unsigned mask;         // lanes of the group
  int singular;
  int dc_src;            // DC sweep: element whose dc value is overridden
  double dc_val;
  double C1;             // DF coefficient of x_{n+1} (0 outside transients)

  AHK_MEM double P(int slot) const {
    int v = p->slot_vary[slot];
    return v < 0 ? p->nominal[slot] : vary[(long long)v * B + inst];
  }
  AHK_MEM double* A() const { return sm + L->A; }
  AHK_MEM double* x() const { return sm + L->x; }
  AHK_MEM unsigned char* prow() const { return (unsigned char*)(sm + L->bytes); }
  AHK_MEM unsigned char* stepof() const { return (unsigned char*)(sm + L->bytes) + 64; }
};

AHK_DEV double V_of(const double* x, int node) { return node >= 0 ? x[node] : 0.0; }

// ---- prologue: per-instance values of the linear stamps -------------------------
// dc_analysis.generate_mna_and_N (dc_analysis.py:949-1069) + transient.generate_D
// (transient.py:472-561) evaluated into the compact value arrays Mv/Dv that
// parallel the merged position list.  Serial (lane 0): runs once per instance.
AHK_DEV void build_values(Ctx& c, bool with_D) {
  const Prog& p = *c.p;
  double* Mv = c.sm + c.L->Mv;
  double* Dv = c.sm + c.L->Dv;
  if (c.sub == 0) {
    for (int k = 0; k < p.npos; k++) { Mv[k] = 0.0; if (with_D) Dv[k] = 0.0; }
    for (int k = 0; k < p.nlin; k++) {
      const LinOp& op = p.lin[k];
      const int* q = op.pos;
#define ADDP(arr, i, v) do { if (q[i] >= 0) arr[q[i]] += (v); } while (0)
#define SETP(arr, i, v) do { if (q[i] >= 0) arr[q[i]] = (v); } while (0)
      switch (op.kind) {
        case K_R: { double g = 1. / c.P(op.pbase); ADDP(Mv, 0, g); ADDP(Mv, 1, -g); ADDP(Mv, 2, -g); ADDP(Mv, 3, g); } break;
        case K_G: { double a = c.P(op.pbase); ADDP(Mv, 0, a); ADDP(Mv, 1, -a); ADDP(Mv, 2, -a); ADDP(Mv, 3, a); } break;
        case K_VDE: SETP(Mv, 0, 1.0); SETP(Mv, 1, -1.0); SETP(Mv, 2, 1.0); SETP(Mv, 3, -1.0); break;
        case K_E: { double a = c.P(op.pbase); SETP(Mv, 0, 1.0); SETP(Mv, 1, -1.0); SETP(Mv, 2, 1.0); SETP(Mv, 3, -1.0); SETP(Mv, 4, -1.0 * a); SETP(Mv, 5, a); } break;
        case K_H: { double a = c.P(op.pbase); SETP(Mv, 0, 1.0); SETP(Mv, 1, -1.0); SETP(Mv, 2, 1.0); SETP(Mv, 3, -1.0); SETP(Mv, 4, a); } break;
        case K_F: { double a = c.P(op.pbase); ADDP(Mv, 0, a); ADDP(Mv, 1, -a); } break;
        case K_C: if (with_D) { double v = c.P(op.pbase); ADDP(Dv, 0, v); ADDP(Dv, 1, -v); ADDP(Dv, 2, v); ADDP(Dv, 3, -v); } break;
        case K_L: if (with_D) { SETP(Dv, 0, -1 * c.P(op.pbase)); } break;
        case K_K: if (with_D) { double M = sqrt(c.P(op.aux0) * c.P(op.aux1)) * c.P(op.pbase); ADDP(Dv, 0, -1 * M); ADDP(Dv, 1, -1 * M); } break;
      }
#undef ADDP
#undef SETP
    }
    if (with_D && c.o->cmin > 0)
      for (int k = 0; k < p.npos; k++) if (p.pdiag[k]) Dv[k] += c.o->cmin;
    // fresh device state (diode.py:89, ekv.py:107, switch.py:98)
    double* st = c.sm + c.L->dst;
    for (int d = 0; d < p.ndev; d++) {
      const DevOp& dv = p.dev[d];
      st[2 * d] = st[2 * d + 1] = 0.0;
      if (dv.type == AHK_D) st[2 * d] = .425;
      if (dv.type == AHK_EKV) st[2 * d] = st[2 * d + 1] = 1.0;
      if (dv.type == AHK_SW) st[2 * d] = c.P(dv.pbase);
    }
  }
}

// N of generate_mna_and_N: DC current sources, then -V on the KVL rows.
AHK_DEV void build_N(Ctx& c) {
  const Prog& p = *c.p;
  double* Nc = c.sm + c.L->Nc;
  if (c.sub == 0) {
    for (int i = 0; i < p.n; i++) Nc[i] = 0.0;
    for (int k = 0; k < p.nsrc; k++) {
      const SrcOp& s = p.src[k];
      if (s.tf) continue;
      double v = (s.elem == c.dc_src) ? c.dc_val : c.P(s.pbase);
      if (s.is_v) Nc[s.row0] = -1.0 * v;
      else { if (s.row0 >= 0) Nc[s.row0] += v; if (s.row1 >= 0) Nc[s.row1] -= v; }
    }
  }
}

// ---- device evaluation: one device per lane ------------------------------------
template <int G>
AHK_DEV void eval_devices(Ctx& c, const double* x) {
  const Prog& p = *c.p;
  const Opts& o = *c.o;
  double* dout = c.sm + c.L->dout;
  double* dst = c.sm + c.L->dst;
  for (int d = c.sub; d < p.ndev; d += G) {
    const DevOp& dv = p.dev[d];
    if (!dv.active) continue;
    double* out = dout + dv.dout;
    double* st = dst + 2 * d;
    switch (dv.type) {
      case AHK_D: {
        DiodeP m;
        m.IS = c.P(dv.mbase + 0); m.N = c.P(dv.mbase + 1); m.NBV = c.P(dv.mbase + 2);
        m.ISR = c.P(dv.mbase + 3); m.NR = c.P(dv.mbase + 4); m.RS = c.P(dv.mbase + 5);
        m.BV = c.P(dv.mbase + 6); m.IKF = c.P(dv.mbase + 7); m.VT = c.P(dv.mbase + 8);
        m.AREA = c.P(dv.pbase);
        diode_eval(m, V_of(x, dv.node[0]) - V_of(x, dv.node[1]), o.vea, o.gmin, st, out);
      } break;
      case AHK_EKV: {
        EkvP m;
        m.VTO = c.P(dv.mbase + 0); m.KP = c.P(dv.mbase + 1); m.GAMMA = c.P(dv.mbase + 2);
        m.PHI = c.P(dv.mbase + 3); m.COX = c.P(dv.mbase + 4); m.LAMBDA = c.P(dv.mbase + 5);
        m.XJ = c.P(dv.mbase + 6); m.UCRIT = c.P(dv.mbase + 7);
        m.W = c.P(dv.pbase + 0); m.L = c.P(dv.pbase + 1); m.M = c.P(dv.pbase + 2);
        m.N = c.P(dv.pbase + 3); m.NP = dv.flags;
        double vb = V_of(x, dv.node[3]);
        ekv_eval(m, V_of(x, dv.node[0]) - vb, V_of(x, dv.node[1]) - vb,
                 V_of(x, dv.node[2]) - vb, o.gmin, st, out);
      } break;
      case AHK_MOSQ: {
        MosqP m;
        m.VTO = c.P(dv.mbase + 0); m.KP = c.P(dv.mbase + 1); m.GAMMA = c.P(dv.mbase + 2);
        m.PHI = c.P(dv.mbase + 3); m.LAMBDA = c.P(dv.mbase + 4); m.AVT = c.P(dv.mbase + 5);
        m.AKP = c.P(dv.mbase + 6);
        m.W = c.P(dv.pbase + 0); m.L = c.P(dv.pbase + 1); m.M = c.P(dv.pbase + 2);
        m.N = c.P(dv.pbase + 3); m.ZVT = c.P(dv.pbase + 4); m.ZKP = c.P(dv.pbase + 5);
        m.NP = dv.flags;
        double vs = V_of(x, dv.node[2]);
        mosq_eval(m, V_of(x, dv.node[0]) - vs, V_of(x, dv.node[1]) - vs,
                  V_of(x, dv.node[3]) - vs, o.iea, o.gmin, out);
      } break;
      case AHK_SW: {
        SwP m;
        m.VT = c.P(dv.mbase + 0); m.VH = c.P(dv.mbase + 1); m.RON = c.P(dv.mbase + 2);
        m.ROFF = c.P(dv.mbase + 3);
        switch_eval(m, V_of(x, dv.node[0]) - V_of(x, dv.node[1]),
                    V_of(x, dv.node[2]) - V_of(x, dv.node[3]), o.gmin, st, out);
      } break;
    }
  }
  Grp<G>::sync(c.mask);
}

// ---- assembly: A = [Bv + J | -(Bv.x + T + Tx)], row-local, no conflicts -----------
// residuo = mna.dot(x) + T + Tx (dc_analysis.py:816); the matrix solved is mna + J.
template <int G>
AHK_DEV void assemble(Ctx& c) {
  const Prog& p = *c.p;
  const int n = p.n, RS = c.L->RS;
  double* A = c.A();
  const double* x = c.x();
  const double* Bv = c.sm + c.L->Bv;
  const double* T = c.sm + c.L->T;
  const double* dout = c.sm + c.L->dout;
  double* res = c.sm + c.L->res;
  for (int r = c.sub; r < n; r += G) {
    double* Ar = A + r * RS;
    for (int j = 0; j < n; j++) Ar[j] = 0.0;
    double s = 0.0;
    for (int k = p.row_ptr[r]; k < p.row_ptr[r + 1]; k++) {
      double v = Bv[k];
      int col = p.pcol[k];
      s += v * x[col];
      for (int q = p.nl_ptr[k]; q < p.nl_ptr[k + 1]; q++) v += (double)p.nl_sgn[q] * dout[p.nl_src[q]];
      Ar[col] = v;
    }
    double tx = 0.0;
    for (int q = p.tx_ptr[r]; q < p.tx_ptr[r + 1]; q++) tx += (double)p.tx_sgn[q] * dout[p.tx_src[q]];
    double rr = s + T[r] + tx;
    res[r] = rr;
    Ar[n] = -rr;
  }
  Grp<G>::sync(c.mask);
}

// ---- dense solve: Gaussian elimination, implicit partial pivoting --------------------
// numpy.linalg.solve (LAPACK dgesv) at dc_analysis.py:822.  Same pivot rule (largest
// |a| in the column among the remaining rows); rows are never moved, the pivot order
// is recorded in prow[]/stepof[].  Returns 0 if a pivot is exactly zero (LinAlgError).
template <int G>
AHK_DEV int lu_solve(Ctx& c) {
  const int n = c.p->n, RS = c.L->RS;
  double* A = c.A();
  double* dx = c.sm + c.L->dx;
  unsigned char* prow = c.prow();
  unsigned char* stepof = c.stepof();
  unsigned long long used = 0ull;
  for (int k = 0; k < n; k++) {
    double best = -1.0;
    int bi = -1;
    for (int r = c.sub; r < n; r += G) {
      if ((used >> r) & 1ull) continue;
      double v = fabs(A[r * RS + k]);
      if (bi < 0 || v > best) { best = v; bi = r; }   // NaN never wins after the first
    }
    Grp<G>::argmax(best, bi, c.mask);
    if (best == 0.0) return 0;
    used |= 1ull << bi;
    if (c.sub == 0) { prow[k] = (unsigned char)bi; stepof[bi] = (unsigned char)k; }
    const double* Ap = A + bi * RS;
    double inv = 1.0 / Ap[k];
    for (int r = c.sub; r < n; r += G) {
      if ((used >> r) & 1ull) continue;
      double* Ar = A + r * RS;
      double l = Ar[k] * inv;
      if (l == 0.0) continue;                       // structural zeros: skip the row
      for (int j = k + 1; j <= n; j++) Ar[j] -= l * Ap[j];
    }
    Grp<G>::sync(c.mask);
  }
  for (int k = n - 1; k >= 0; k--) {
    int pr = prow[k];
    double xk = A[pr * RS + n] / A[pr * RS + k];
    if (c.sub == 0) dx[k] = xk;
    for (int r = c.sub; r < n; r += G) {
      if ((int)stepof[r] < k) { double a = A[r * RS + k]; if (a != 0.0) A[r * RS + n] -= a * xk; }
    }
    Grp<G>::sync(c.mask);
  }
  return 1;
}

// np.allclose element test (utilities.py:527-529)
AHK_DEV int close1(double a, double b, double rtol, double atol) {
  if (isnan(a) || isnan(b)) return 0;
  if (isinf(a) || isinf(b)) return a == b;
  return fabs(a - b) <= atol + rtol * fabs(b);
}

// ---- damped Newton (dc_analysis.mdn_solver) -------------------------------------------
// returns 1 converged, 0 not converged, -1 singular (x untouched in that iteration)
template <int G>
AHK_DEV int mdn_solver(Ctx& c, int MAXIT, int* iters) {
  const Prog& p = *c.p;
  const Opts& o = *c.o;
  const int n = p.n;
  double* x = c.x();
  double* dx = c.sm + c.L->dx;
  const double* res = c.sm + c.L->res;
  const double lim = o.nl_voltages_lock_factor * AHK_VTH;
  int it = 0, conv = 0;
  while (it < MAXIT) {
    it++;
    if (p.ndev) eval_devices<G>(c, x);
    assemble<G>(c);
    if (!lu_solve<G>(c)) { c.singular = 1; *iters = 0; return -1; }
    // get_td (dc_analysis.py:884-946)
    double td = 1.0;
    if (o.nr_damp_first_iters) td = it < 10 ? 1e-2 : (it < 20 ? 0.1 : 1.0);
    if (o.nl_voltages_lock && p.nlock) {
      double t = 1.0;
      for (int k = c.sub; k < p.nlock; k += G) {
        int n1 = p.lock[2 * k], n2 = p.lock[2 * k + 1];
        double d;
        if (n1 != 0) d = n2 != 0 ? fabs(dx[n1 - 1] - dx[n2 - 1]) : fabs(dx[n1 - 1]);
        else d = fabs(dx[n2 != 0 ? n2 - 1 : n - 1]);   // dx[-1] quirk, dc_analysis.py:941
        if (d > lim) t = fmin(t, lim / d);
      }
      td = fmin(td, Grp<G>::fmin_all(t, c.mask));
    }
    int ok = 1, finite = 1;
    for (int i = c.sub; i < n; i += G) {
      double d = dx[i];
      double xn = x[i] + td * d;
      x[i] = xn;
      if (!isfinite(xn)) finite = 0;
      bool isv = i < p.nv1;
      if (!close1(xn, xn + d, isv ? o.ver : o.ier, isv ? o.vea : o.iea)) ok = 0;
      if (!close1(res[i], 0.0, 0.0, isv ? o.iea : o.vea)) ok = 0;
    }
    Grp<G>::sync(c.mask);
    if (!p.ndev) { conv = 1; break; }            // linear circuit: one iteration
    if (Grp<G>::all(ok, c.mask)) { conv = 1; break; }
    if (!Grp<G>::all(finite, c.mask)) { it = MAXIT; break; }  // NaN never converges
  }
  *iters = it;
  return conv;
}

// ---- dc_solve: standard -> gmin stepping -> source stepping (dc_analysis.py:128-320) ---
// gmin: value on the node diagonals for the standard / source-stepping attempts.
template <int G>
AHK_DEV int dc_solve(Ctx& c, double gmin, bool with_tran, double time, int MAXIT,
                     int* tot_iters, int* method) {
  const Prog& p = *c.p;
  const Opts& o = *c.o;
  const int n = p.n;
  double* Nd = c.sm + c.L->Nd;
  const double* Nc = c.sm + c.L->Nc;
  const double* Ntr = c.sm + c.L->Ntr;
  double* T = c.sm + c.L->T;
  double* Bv = c.sm + c.L->Bv;
  const double* Mv = c.sm + c.L->Mv;
  const double* Dv = c.sm + c.L->Dv;
  for (int i = c.sub; i < n; i += G) Nd[i] = Nc[i];
  Grp<G>::sync(c.mask);
  if (p.ntd && c.sub == 0) {   // Tt(t), dc_analysis.py:222-237
    for (int k = 0; k < p.nsrc; k++) {
      const SrcOp& s = p.src[k];
      if (!s.tf) continue;
      double q[8];
      for (int j = 0; j < 8; j++) q[j] = c.P(s.pbase + 3 + j);
      double v = tf_eval(s.tf, q, time);
      if (s.is_v) Nd[s.row0] += -1 * v;
      else { if (s.row0 >= 0) Nd[s.row0] += v; if (s.row1 >= 0) Nd[s.row1] -= v; }
    }
  }
  int failed[3] = {0, 0, 0};
  const int use[3] = {o.use_standard_solve_method, o.use_gmin_stepping, o.use_source_stepping};
  int enabled = -1, gidx = 0, sidx = 0;
  for (int m = 0; m < 3; m++) if (use[m]) { enabled = m; break; }
  *tot_iters = 0;
  *method = 0;
  while (true) {
    if (enabled < 0) return 0;
    double g = gmin, f = 1.0;
    if (enabled == 1) {   // 10 ** -(idx+1), replaces Gmin (dc_analysis.py:262-269,452-455)
      const double g10[10] = {1e-1, 1e-2, 1e-3, 1e-4, 1e-5, 1e-6, 1e-7, 1e-8, 1e-9, 1e-10};
      g = g10[gidx];
    } else if (enabled == 2) {
      const double sfac[10] = {0.001, .005, .01, .03, .1, .3, .5, .7, .8, .9};
      f = sfac[sidx];
    }
    Grp<G>::sync(c.mask);
    for (int k = c.sub; k < p.npos; k += G) {
      double v = Mv[k];
      if (with_tran) v += c.C1 * Dv[k];
      if (p.pdiag[k]) v += g;
      Bv[k] = v;
    }
    for (int i = c.sub; i < n; i += G) T[i] = (enabled == 2 ? f * Nd[i] : Nd[i]) + (with_tran ? Ntr[i] : 0.0);
    Grp<G>::sync(c.mask);
    int it = 0;
    int r = mdn_solver<G>(c, MAXIT, &it);
    if (r >= 0) *tot_iters += it;
    *method = enabled;
    if (r <= 0) {   // set_next_solve_method, dc_analysis.py:354-408
      failed[enabled] = 1;
      enabled = -1;
      for (int m = 0; m < 3; m++) if (!failed[m] && use[m]) { enabled = m; break; }
      if (enabled < 0) return 0;
    } else {
      if (enabled == 2 && sidx != 9) sidx++;
      else if (enabled == 1 && gidx != 9) gidx++;
      else return 1;
    }
  }
}

// ---- OP: guess -> Gmin pass -> no-Gmin pass -> gmin_check (dc_analysis.py:598-687) ---
// in: x() holds x0 if have_x0; out: x() solution; returns AHK_ST_* word
template <int G>
AHK_DEV int op_analysis(Ctx& c, bool have_x0, bool use_guess, int* iters) {
  const Prog& p = *c.p;
  const Opts& o = *c.o;
  const int n = p.n;
  double* x = c.x();
  double* x1 = c.sm + c.L->x1;
  build_N(c);
  if (!have_x0) {
    for (int i = c.sub; i < n; i += G) {
      double s = 0.0;
      if (use_guess && p.nguess > 0) {   // dc_guess.py via the precomputed operator
        for (int j = 0; j < p.nguess; j++) {
          double T = p.gconst[j];
          if (p.gslot[j] >= 0) T += p.gscale[j] * c.P(p.gslot[j]);
          s += p.guessG[i * p.nguess + j] * T;
        }
      }
      x[i] = s;
    }
  }
  Grp<G>::sync(c.mask);
  int n1 = 0, n2 = 0, m1 = 0, m2 = 0, status;
  int s1 = dc_solve<G>(c, o.gmin, false, 0.0, o.dc_max_nr_iter, &n1, &m1);
  if (!s1) {
    *iters = n1;
    status = (m1 << 1);
    const double qnan = nan("");
    for (int i = c.sub; i < n; i += G) x[i] = qnan;
  } else {
    double* r1 = c.sm + c.L->xs;   // residual of pass 1 (kept for PASS2_FAIL)
    const double* res = c.sm + c.L->res;
    for (int i = c.sub; i < n; i += G) { x1[i] = x[i]; r1[i] = res[i]; }
    Grp<G>::sync(c.mask);
    int s2 = dc_solve<G>(c, 0.0, false, 0.0, o.dc_max_nr_iter, &n2, &m2);
    if (!s2) {
      double* resw = c.sm + c.L->res;
      for (int i = c.sub; i < n; i += G) { x[i] = x1[i]; resw[i] = r1[i]; }
      *iters = n1;
      status = AHK_ST_OK | AHK_ST_PASS2_FAIL | (m1 << 1);
    } else {
      *iters = n1 + n2;
      int bad = 0;
      for (int i = c.sub; i < n; i += G) {   // results.py:493-523
        bool isv = i < p.nv1;
        double er = isv ? o.ver : o.ier, ea = isv ? o.vea : o.iea;
        if (fabs(x[i] - x1[i]) > er * fmax(fabs(x1[i]), fabs(x[i])) + ea) bad = 1;
      }
      status = AHK_ST_OK | (m2 << 1);
      if (Grp<G>::any(bad, c.mask)) status |= AHK_ST_GMIN_CHECK;
    }
  }
  if (c.singular) status |= AHK_ST_SINGULAR;
  Grp<G>::sync(c.mask);
  return status;
}

// ---- transient (transient.py:121-470) ----------------------------------------------------
AHK_DEV int check_step(const Opts& o, double& tstep, double time, double tstop, double HMAX) {
  if (tstep > HMAX) tstep = HMAX;
  if (tstop - time < tstep) tstep = tstop - time;
  else if (tstep < o.hmin) return -1;
  return 0;
}

struct DfOut { double C1, x_lte, pred_lte; int has_pred; };

// History ring: entry k (k = 0 newest) lives in slot (h0 + k) % buflen.
template <int G>
AHK_DEV void get_df(Ctx& c, int method, int order, bool bootstrap, int nh, int h0, int buflen,
                    double step, bool predict, DfOut& d) {
  const int n = c.p->n;
  const double* ht = c.sm + c.L->ht;
  const double* hx = c.sm + c.L->hx;
  const double* hdx = c.sm + c.L->hdx;
  double* C0 = c.sm + c.L->C0;
  double* pred = c.sm + c.L->pred;
#define HT(k) ht[(h0 + (k)) % buflen]
#define HX(k) (hx + ((h0 + (k)) % buflen) * n)
  d.has_pred = 0;
  d.pred_lte = 0.0;
  if (bootstrap || method == AHK_IMPLICIT_EULER) {   // implicit_euler.py:125-205
    int avail = bootstrap ? 1 : nh;
    d.x_lte = 0.5 * step;
    d.C1 = 1. / step;
    const double* x0 = HX(0);
    if (predict && avail > 1) {
      const double* xm = HX(1);
      double dt = HT(0) - HT(1);
      for (int i = c.sub; i < n; i += G) pred[i] = (x0[i] - xm[i]) / dt * step + x0[i];
      d.pred_lte = -0.5 * step * (HT(0) + step - HT(1));
      d.has_pred = 1;
    }
    for (int i = c.sub; i < n; i += G) C0[i] = -1. / step * x0[i];
  } else if (method == AHK_TRAP) {                   // trap.py:98-192
    d.C1 = 2.0 / step;
    d.x_lte = 1.0 / 12 * step * step * step;
    const double* x0 = HX(0);
    for (int i = c.sub; i < n; i += G) C0[i] = -1.0 * (2.0 / step * x0[i] + hdx[i]);
    if (predict && nh > 2) {
      // quadratic through the last three points (trap.py:171-186 solves the same
      // interpolation with inv(A)); divided differences about t_n
      const double* xa = HX(1);
      const double* xb = HX(2);
      double t1 = HT(1) - HT(0), t2 = HT(2) - HT(0);
      for (int i = c.sub; i < n; i += G) {
        double d1 = (xa[i] - x0[i]) / t1, d2 = (xb[i] - x0[i]) / t2;
        double a2 = (d1 - d2) / (t1 - t2), a1 = d1 - a2 * t1;
        pred[i] = a2 * step * step + a1 * step + x0[i];
      }
      d.pred_lte = -1.0 / 6.0 * step * (HT(0) + step - HT(1)) * (HT(0) + step - HT(2));
      d.has_pred = 1;
    }
  } else {                                           // gear.py:122-208
    double s[9], alpha[9], gamma[9];
    s[0] = 0;
    for (int i = 1; i < order + 2; i++) s[i] = step + HT(0) - HT(i - 1);
    for (int k = 1; k < order + 2; k++) {
      double a = 1.0;
      for (int j = 1; j < order + 2; j++) a *= (j == k) ? 1.0 : s[j] / (s[j] - s[k]);
      alpha[k] = a;
    }
    gamma[0] = 0;
    for (int k = 1; k < order + 1; k++) {
      gamma[k] = alpha[k] * ((1.0 / s[order + 1]) - (1.0 / s[k]));
      gamma[0] -= gamma[k];
    }
    d.C1 = gamma[0];
    for (int i = c.sub; i < n; i += G) {
      double a = 0.0;
      for (int k = 0; k < order; k++) a = a + gamma[k + 1] * HX(k)[i];
      C0[i] = a;
    }
    double xl = 0.0, fact = 1.0;
    for (int k = 2; k <= order + 1; k++) fact *= k;
    for (int k = 1; k < order + 1; k++) xl += pow(s[k], (double)(order + 1)) * (-1.0 * gamma[k] / gamma[0]);
    d.x_lte = (((order + 1) & 1) ? -1.0 : 1.0) * (1.0 / fact) * xl;
    if (predict) {
      for (int i = c.sub; i < n; i += G) {
        double a = 0.0;
        for (int k = 1; k < order + 2; k++) a = a + alpha[k] * HX(k - 1)[i];
        pred[i] = a;
      }
      double pl = -1.0 / fact;
      for (int k = 1; k < order + 2; k++) pl *= s[k];
      d.pred_lte = pl;
      d.has_pred = 1;
    }
  }
#undef HT
#undef HX
  Grp<G>::sync(c.mask);
}

struct TranArgs {
  double tstart, tstep, tstop;
  int method, use_step_control;
  long long max_rows;
  double* t_out; double* x_out;
  int* rows; int* counters;
};

// in: x() holds x(tstart).  returns status word.
template <int G>
AHK_DEV int tran_analysis(Ctx& c, const TranArgs& a) {
  const Prog& p = *c.p;
  const Opts& o = *c.o;
  const int n = p.n;
  const int buflen = c.L->buflen;
  double* x = c.x();
  double* xacc = c.sm + c.L->xacc;   // last accepted x
  double* ht = c.sm + c.L->ht;
  double* hx = c.sm + c.L->hx;
  double* hdx = c.sm + c.L->hdx;
  double* C0 = c.sm + c.L->C0;
  double* pred = c.sm + c.L->pred;
  double* Ntr = c.sm + c.L->Ntr;
  const double* Dv = c.sm + c.L->Dv;
  double* xs = c.sm + c.L->xs;       // NR start point when nothing better exists
  int order, max_x, max_dx, pmax_x;
  if (a.method == AHK_IMPLICIT_EULER) { order = 1; max_x = 0; max_dx = -1; pmax_x = 1; }
  else if (a.method == AHK_TRAP) { order = 2; max_x = 0; max_dx = 0; pmax_x = 2; }
  else { order = a.method - 10; max_x = order; max_dx = -1; pmax_x = order; }
  int first_iters = 0;   // transient.py:272-278
  if (max_x > 0 || max_dx >= 0) { first_iters = max_x; if (max_dx >= 0 && max_dx + 1 > first_iters) first_iters = max_dx + 1; }
  const int start_pred = pmax_x > 0 ? pmax_x : 0;
  const double HMAX = a.tstep;
  double tstep = a.tstep;
  build_N(c);
  int h0 = 0, nh = 1;
  for (int i = c.sub; i < n; i += G) { hx[i] = x[i]; xs[i] = x[i]; xacc[i] = x[i]; }
  if (c.sub == 0) ht[0] = a.tstart;
  Grp<G>::sync(c.mask);
  double time = a.tstart;
  if (a.use_step_control) tstep = fmin((a.tstop - a.tstart) / 9999.0, HMAX);
  int iter_n = 0, solved = 1, have_x = 0, status = 0;
  long long nrows = 0;
  int n_solves = 0, n_nr = 0, n_rej = 0, n_cut = 0;
  double* t_out = a.t_out + c.inst * a.max_rows;
  double* x_out = a.x_out + c.inst * a.max_rows * n;
  while (time < a.tstop) {
    DfOut d;
    bool predict = a.use_step_control && iter_n >= start_pred;
    get_df<G>(c, a.method, order, iter_n < first_iters, nh, h0, buflen, tstep, predict, d);
    // NR start point (transient.py:335-338)
    if (o.transient_prediction_as_x0 && a.use_step_control && d.has_pred) {
      for (int i = c.sub; i < n; i += G) x[i] = pred[i];
    } else if (have_x) {
      for (int i = c.sub; i < n; i += G) x[i] = xacc[i];
    } else {
      for (int i = c.sub; i < n; i += G) x[i] = xs[i];
    }
    // Ntran = D . C0
    for (int r = c.sub; r < n; r += G) {
      double s = 0.0;
      for (int k = p.row_ptr[r]; k < p.row_ptr[r + 1]; k++) s += Dv[k] * C0[p.pcol[k]];
      Ntr[r] = s;
    }
    c.C1 = d.C1;
    Grp<G>::sync(c.mask);
    int it = 0, meth = 0;
    int ok = dc_solve<G>(c, o.gmin, true, time + tstep, o.transient_max_nr_iter, &it, &meth);
    n_solves++; n_nr += it;
    if (ok) {
      double old_step = tstep;
      if (a.use_step_control && d.has_pred) {
        double coeff = 2.0;
        double ex = 1.0 / (order + 1);
        for (int i = c.sub; i < n; i += G) {   // transient.py:356-365
          double lte = fabs((d.x_lte / (d.pred_lte - d.x_lte)) * (pred[i] - x[i]));
          if (lte != 0) {
            double re = i < p.nv1 ? o.ver : o.ier;
            double v = pow((o.vea + re * fabs(xacc[i])) / lte, ex);
            if (v < coeff) coeff = v;
          }
        }
        coeff = Grp<G>::fmin_all(coeff, c.mask);
        double new_step = tstep * coeff;
        bool reject = o.transient_use_aposteriori_step_control &&
                      coeff < o.transient_aposteriori_step_threshold;
        tstep = new_step;
        if (check_step(o, tstep, time, a.tstop, HMAX)) { status |= AHK_ST_HMIN; solved = 0; break; }
        if (reject) { n_rej++; continue; }
      }
      time = time + old_step;
      have_x = 1;
      iter_n++;
      if (nrows >= a.max_rows) { status |= AHK_ST_ROWS_FULL; solved = 0; break; }
      // accept: store the row, dxdt = C1*x + C0, push to the ring (transient.py:385-393)
      h0 = (h0 + buflen - 1) % buflen;
      double* hnew = hx + h0 * n;
      for (int i = c.sub; i < n; i += G) {
        double xi = x[i];
        xacc[i] = xi;
        hnew[i] = xi;
        hdx[i] = d.C1 * xi + C0[i];
        x_out[nrows * n + i] = xi;
      }
      if (c.sub == 0) { ht[h0] = time; t_out[nrows] = time; }
      nrows++;
      if (nh < buflen) nh++;
      Grp<G>::sync(c.mask);
    } else {
      if (a.use_step_control) {
        tstep = tstep / 5.0;
        if (check_step(o, tstep, time, a.tstop, HMAX)) { status |= AHK_ST_HMIN; solved = 0; break; }
        n_cut++;
      } else { status |= AHK_ST_NR_FAIL; solved = 0; break; }
    }
    if (o.transient_max_time_iter && iter_n == o.transient_max_time_iter) { solved = 0; break; }
  }
  if (solved) status |= AHK_ST_OK;
  if (c.singular) status |= AHK_ST_SINGULAR;
  if (c.sub == 0) {
    a.rows[c.inst] = (int)nrows;
    a.counters[4 * c.inst + 0] = n_solves; a.counters[4 * c.inst + 1] = n_nr;
    a.counters[4 * c.inst + 2] = n_rej; a.counters[4 * c.inst + 3] = n_cut;
  }
  return status;
}

}  // namespace ahk
