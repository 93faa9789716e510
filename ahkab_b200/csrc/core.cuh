// core.cuh — per-instance Newton/MNA machinery of the small-circuit engine.
//
// Execution model: a GROUP of G lanes (G = 1,2,4,8,16,32, a sub-warp) owns one
// circuit instance for the whole analysis.  A kernel variant Var<NC,R,G> fixes at
// compile time the number of matrix columns NC (>= n+1, the last one is the
// right-hand side), the rows per lane R and the lanes per instance G (R*G >= n):
// row r lives in lane r % G, slot r / G, as NC doubles IN REGISTERS for the whole
// linear solve (G > 1: Gauss-Jordan with implicit partial pivoting, fully unrolled, the
// pivot row crossing lanes through a double-buffered shared-memory broadcast; G == 1: LU
// with explicit register swaps).  The per-instance
// parameters are staged once from HBM into shared memory, together with the
// sparse values of the linear stamps, the iterate, the history ring and the
// device outputs; HBM is touched again only to write result rows.
// Var<0,0,G> is the generic fallback for n > 31: run-time n, matrix in shared
// memory, LU with back-substitution.
// Control scalars (time, step, counters) are kept redundantly in registers and
// are group-uniform by construction.
//
// This source is compiled twice: ahead of time (k_dc.cu / k_tran.cu / k_ac.cu), where it
// INTERPRETS the stamp program (`Prog`), and at run time by NVRTC with `Prog`, `Opts` and
// `Layout` replaced by compile-time constants (AHK_JIT, jit_gen.hpp), where the loops
// marked AHK_UNROLL / AHK_LANES unroll and the program folds into straight-line code.
//
// What it replaces in the reference (all per instance, pure Python there):
//   dc_analysis.mdn_solver / _update_J_and_Tx / get_td   dc_analysis.py:690-946
//   utilities.convergence_check                          utilities.py:392-549
//   dc_analysis.dc_solve (solve-method state machine)    dc_analysis.py:128-320
//   dc_analysis.op_analysis / dc_analysis                dc_analysis.py:461-687
//   transient.transient_analysis + IE/TRAP/GEAR get_df   transient.py:121-470,
//       implicit_euler.py:125-205, trap.py:98-192, gear.py:122-208
#pragma once
#ifndef __CUDACC_RTC__
#include <string.h>
#endif

#include "devices.cuh"
#include "program.h"

// Specialised build (jit_gen.hpp): the program is a set of constants, so the loops over it
// have constant bounds — unrolling them lets every index load and branch fold away.  The
// interpreted build keeps them rolled (run-time bounds).
#if defined(AHK_JIT) && defined(__CUDACC__)
#define AHK_UNROLL _Pragma("unroll")
#else
#define AHK_UNROLL
#endif
// "for (i = lane; i < cnt; i += G)": the items of a vector this lane owns.  The specialised
// build gives the loop a constant trip count (cnt and G are constants, the lane is not).
#ifdef AHK_JIT
#define AHK_LANES(i, cnt)                                                  \
  AHK_UNROLL for (int i##_q = 0; i##_q < ((cnt) + G - 1) / G; i##_q++)     \
    if (const int i = c.sub + i##_q * G; i < (cnt))
#else
#define AHK_LANES(i, cnt) for (int i = c.sub; i < (cnt); i += G)
#endif

// "items of a small vector" in the flat transient kernel.  REPLICATED shapes (every lane of the
// group holds all rows, see assemble_solve) also replicate the cheap vector work: every lane
// handles EVERY item, so that after unrolling all indices are compile-time constants (no
// constant-table look-ups by lane, no cross-lane reductions); the lanes store identical values.
#define AHK_ITEMS(i, cnt)                                                                      \
  AHK_UNROLL for (int i##_q = 0; i##_q < (REP ? (cnt) : ((cnt) + G - 1) / G); i##_q++)         \
    if (const int i = REP ? i##_q : c.sub + i##_q * G; i < (cnt))

namespace ahk {

template <int NC_, int R_, int G_>
struct Var { static const int NC = NC_, R = R_, G = G_; };

// replicated shape: several lanes per instance, each holding all rows (specialised builds only)
template <class V>
struct RepOf {
#ifdef AHK_JIT
  static const bool value = V::G > 1 && V::NC > 0 && V::R >= Prog::n;
#else
  static const bool value = false;
#endif
};

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
  static AHK_DEV double fmax_all(double v, unsigned m) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(m, v, o));
    return v;
  }
  static AHK_DEV double sum_all(double v, unsigned m) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(m, v, o);
    return v;
  }
  static AHK_DEV int all(int pred, unsigned m) { return G > 1 ? __all_sync(m, pred) : pred; }
  static AHK_DEV int any(int pred, unsigned m) { return G > 1 ? __any_sync(m, pred) : pred; }
  static AHK_DEV double bcast(double v, int src, unsigned m) { return __shfl_sync(m, v, src, G); }
  // group maximum of a 64-bit key: xor butterfly (measured faster on B200 than two
  // dependent 32-bit REDUX.MAX: c3 21.6 vs 30.8 ms, c4 217 vs 297 ms)
  static AHK_DEV unsigned long long kmax(unsigned long long k, unsigned m) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      const unsigned long long ok = __shfl_xor_sync(m, k, o);
      if (ok > k) k = ok;
    }
    return k;
  }
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
  static double fmax_all(double v, unsigned) { return v; }
  static double sum_all(double v, unsigned) { return v; }
  static int all(int p, unsigned) { return p; }
  static int any(int p, unsigned) { return p; }
  static double bcast(double v, int, unsigned) { return v; }
  static unsigned long long kmax(unsigned long long k, unsigned) { return k; }
  static void argmax(double&, int&, unsigned) {}
};
#endif

// ---- warp-uniform ("flat") kernels ------------------------------------------------------------
// In the flat transient kernel every lane of a warp executes the same instruction stream: the
// groups of a warp are in different phases of their analyses, but phase-specific work is
// predicated, never branched around cross-lane operations.  All shuffles / votes / barriers
// therefore run with the FULL mask at warp-uniform points (no sub-warp masks: those make the
// compiler guard every shuffle), and a group-level vote is a slice of the warp ballot.
#ifdef __CUDACC__
AHK_DEV bool warp_any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
AHK_DEV bool warp_loop(bool p) { return __any_sync(0xffffffffu, p) != 0; }   // loop control: same vote
template <int G>
struct GrpW {
  static AHK_DEV unsigned slice(unsigned ballot) {
    if (G == 32) return ballot;
    const unsigned lane = threadIdx.x & 31u;
    return (ballot >> (lane & ~(unsigned)(G - 1))) & ((1u << (G & 31)) - 1u);
  }
  static AHK_DEV bool all(bool p) {
    if (G == 1) return p;
    return slice(__ballot_sync(0xffffffffu, p)) == (G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u));
  }
  static AHK_DEV bool any(bool p) {
    if (G == 1) return p;
    return slice(__ballot_sync(0xffffffffu, p)) != 0u;
  }
  static AHK_DEV void sync() { if (G > 1) __syncwarp(); }
};
#else
// host emulation: one lane.  AHK_EMU_ALWAYS (tests) makes every predicated section of the flat
// loop run in every iteration, as it does on a warp whose other groups need it — a state
// update that is not properly predicated then corrupts the lane's own state and shows up
#ifdef AHK_EMU_ALWAYS
AHK_DEV bool warp_any(bool) { return true; }
#else
AHK_DEV bool warp_any(bool p) { return p; }
#endif
AHK_DEV bool warp_loop(bool p) { return p; }
template <int G>
struct GrpW {
  static bool all(bool p) { return p; }
  static bool any(bool p) { return p; }
  static void sync() {}
};
#endif

struct Ctx {
  const Prog* p;
  const Opts* o;
  const Layout* L;
  double* sm;            // this instance's shared-memory slice
  const double* vary;    // [n_vary][B] in HBM
  long long B, inst;
  int sub;               // lane within the group
  int G;                 // lanes per instance (run-time copy of Var::G)
  unsigned mask;         // lanes of the group
  int singular;
  int dc_src;            // DC sweep: element whose dc value is overridden
  double dc_val;
  double C1;             // DF coefficient of x_{n+1} (0 outside transients)
  double gadd;           // Gmin added on the fly to node diagonals (OP/DC layouts: Bv is Mv)
  double* Ag = nullptr;  // large-n layout (Layout::big): this instance's n x RS matrix in the global workspace
  int* pivrec = nullptr; // static-schedule kernels, calibration launch: 32 records, one per instance / lane (layout: slu_record)

  // persistent parameter (staged in shared memory; index = SrcOp/DevOp ppb/pmb + k)
  AHK_MEM double PP(int pslot) const { return sm[L->Pv + pslot]; }
  // any parameter, straight from HBM (set-up only)
  AHK_MEM double P(int slot) const {
    int v = p->slot_vary[slot];
    return v < 0 ? p->nominal[slot] : vary[(long long)v * B + inst];
  }
  double* Vg = nullptr;  // Layout::vglob: base of this instance's Mv / Dv / Bv block in the global workspace
#ifdef AHK_JIT           // specialised kernels exist for n <= 31 only: everything is in shared memory / registers
  AHK_MEM double* A() const { return sm + L->A; }
  AHK_MEM double* vb() const { return sm; }
#else
  AHK_MEM double* A() const { return Ag ? Ag : sm + L->A; }
  AHK_MEM double* vb() const { return Vg ? Vg : sm; }   // base the offsets Layout::Mv / Dv / Bv refer to
#endif
  AHK_MEM double* x() const { return sm + L->x; }
  AHK_MEM unsigned char* prow() const { return (unsigned char*)(sm + L->bytes); }
  AHK_MEM unsigned char* stepof() const { return (unsigned char*)(sm + L->bytes) + 64; }
};

// Pivot key of the register variants (n <= 31): the magnitude's bit pattern (monotone for
// non-negative doubles) with its 5 low mantissa bits replaced by 31 - row, so that one
// integer maximum finds the largest magnitude and breaks ties towards the lowest row
// (LAPACK idamax picks the first maximum).  Magnitudes that agree to 2^-47 count as ties.
AHK_DEV unsigned long long pivot_key(double mag, int row) {
#ifdef __CUDACC__
  unsigned long long b = (unsigned long long)__double_as_longlong(mag);
#else
  unsigned long long b; memcpy(&b, &mag, 8);
#endif
  return (b & ~31ull) | (unsigned long long)(31 - row);
}

AHK_DEV double V_of(const double* x, int node) { return node >= 0 ? x[node] : 0.0; }

// ---- prologue: stage the parameters, per-instance values of the linear stamps ---
// dc_analysis.generate_mna_and_N (dc_analysis.py:949-1069) + transient.generate_D
// (transient.py:472-561) evaluated into the compact value arrays Mv/Dv that
// parallel the merged position list; device constants and fresh device state.
template <class V>
AHK_DEV void build_values(Ctx& c, bool with_D) {
  const Prog& p = *c.p;
  const int G = V::G;
  double* Pv = c.sm + c.L->Pv;
  AHK_LANES(s, p.npersist) Pv[s] = c.P(p.pstage[s]);
#ifndef AHK_JIT                    // (the specialised build has no staging rows)
  if (V::NC > 0 && V::R == 1) {   // staging rows are zeroed once: only the sparsity
    double* A = c.A();              // pattern is rewritten per iteration
    AHK_LANES(i, p.n * c.L->RS) A[i] = 0.0;
  }
#endif
  Grp<V::G>::sync(c.mask);
  double* Mv = c.vb() + c.L->Mv;
  double* Dv = c.vb() + c.L->Dv;
  bool gathered = false;
#ifndef AHK_JIT
  if (c.L->big) {
    // Large circuits: the MNA values in GATHER form, one position per lane (the term lists of
    // program_build.hpp replay the stamp operations below per position, in the same order:
    // same sums, bit for bit).  The serial replay by one lane cost a dense 202-node mesh
    // (19 900 resistors, values in global memory) 40 % of its whole OP (ncu,
    // profiles/r2w_ncu_source_lines_large_n_mesh202.txt).
    for (int q = c.sub; q < p.npos; q += G) {
      double v = 0.0;
      for (int t = p.term_ptr[q]; t < p.term_ptr[q + 1]; t++) {
        const int kd = p.term_kind[t], ak = kd < 0 ? -kd : kd;
        const double sg = kd < 0 ? -1.0 : 1.0;
        if (ak == 1) v += sg * (1. / c.P(p.term_slot[t]));
        else if (ak == 2) v += sg * c.P(p.term_slot[t]);
        else v += sg;
      }
      Mv[q] = v;
      if (with_D) Dv[q] = 0.0;
    }
    gathered = true;
    Grp<V::G>::sync(c.mask);
  }
#endif
  if (c.sub == 0) {
    if (!gathered) {
      AHK_UNROLL
      for (int k = 0; k < p.npos; k++) { Mv[k] = 0.0; if (with_D) Dv[k] = 0.0; }
    }
    AHK_UNROLL
    for (int k = 0; k < p.nlin; k++) {
      const LinOp& op = p.lin[k];
      if (gathered && op.kind < K_C) continue;      // (only the D operations are left to replay)
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
      AHK_UNROLL
      for (int k = 0; k < p.npos; k++) if (p.pdiag[k]) Dv[k] += c.o->cmin;
  }
  // device constants + fresh device state (diode.py:89, switch.py:98): one device per lane
  double* st = c.sm + c.L->dst;
  double* dcon = c.sm + c.L->dcon;
  AHK_LANES(d, p.ndev) {
    const DevOp& dv = p.dev[d];
    if (dv.st >= 0) {
      st[dv.st] = st[dv.st + 1] = 0.0;
      if (dv.type == AHK_D) st[dv.st] = .425;
      if (dv.type == AHK_SW) st[dv.st] = c.PP(dv.ppb);
    }
    if (dv.type == AHK_D) {
      DiodeP m;
      m.IS = c.P(dv.mbase + 0); m.N = c.P(dv.mbase + 1); m.NBV = c.P(dv.mbase + 2);
      m.ISR = c.P(dv.mbase + 3); m.NR = c.P(dv.mbase + 4); m.RS = c.P(dv.mbase + 5);
      m.BV = c.P(dv.mbase + 6); m.IKF = c.P(dv.mbase + 7); m.VT = c.P(dv.mbase + 8);
      m.AREA = c.P(dv.pbase);
      diode_setup(m, dcon + dv.con);
    }
    if (dv.type == AHK_EKV) {
      EkvP m;
      m.VTO = c.P(dv.mbase + 0); m.KP = c.P(dv.mbase + 1); m.GAMMA = c.P(dv.mbase + 2);
      m.PHI = c.P(dv.mbase + 3); m.COX = c.P(dv.mbase + 4); m.LAMBDA = c.P(dv.mbase + 5);
      m.XJ = c.P(dv.mbase + 6); m.UCRIT = c.P(dv.mbase + 7);
      m.W = c.P(dv.pbase + 0); m.L = c.P(dv.pbase + 1); m.M = c.P(dv.pbase + 2);
      m.N = c.P(dv.pbase + 3); m.NP = dv.flags;
      ekv_setup(m, dcon + dv.con);
    }
  }
  Grp<V::G>::sync(c.mask);
}

// N of generate_mna_and_N: DC current sources, then -V on the KVL rows.
AHK_DEV void build_N(Ctx& c) {
  const Prog& p = *c.p;
  double* Nc = c.sm + c.L->Nc;
  if (c.sub == 0) {
    AHK_UNROLL
    for (int i = 0; i < p.n; i++) Nc[i] = 0.0;
    AHK_UNROLL
    for (int k = 0; k < p.nsrc; k++) {
      const SrcOp& s = p.src[k];
      if (s.tf) continue;
      double v = (s.elem == c.dc_src) ? c.dc_val : c.PP(s.ppb);
      if (s.is_v) Nc[s.row0] = -1.0 * v;
      else { if (s.row0 >= 0) Nc[s.row0] += v; if (s.row1 >= 0) Nc[s.row1] -= v; }
    }
  }
}

#if defined(AHK_JIT) && defined(AHK_JIT_ALL_EKV) && AHK_JIT_ALL_EKV && AHK_JIT_ILP > 1
#define AHK_EKV_BATCHED 1
// Every device of this circuit is an (active) EKV transistor: a lane evaluates its devices K at
// a time through ekv_eval_v, which interleaves the K evaluations statement by statement.
// Rounds beyond the device list (ndev not a multiple of G*K) re-evaluate the last device and
// drop the result.  The loop over the batches stays rolled: one copy of the device code.
template <int K>
AHK_DEV void eval_devices_ekv(Ctx& c, const double* x) {
  const Prog& p = *c.p;
  const int G = AHK_JIT_G;
  const int rounds = (p.ndev + G - 1) / G;
  double* dout = c.sm + c.L->dout;
  const double* dcon = c.sm + c.L->dcon;
#pragma unroll 1
  for (int q0 = 0; q0 < rounds; q0 += K) {
    const double* kk[K];
    int NP[K], dd[K];
    bool ok[K];
    double vd[K], vg[K], vs[K], o0[K], o1[K], o2[K], o3[K];
#pragma unroll
    for (int j = 0; j < K; j++) {
      const int d = c.sub + (q0 + j) * G;
      ok[j] = d < p.ndev;
      dd[j] = ok[j] ? d : p.ndev - 1;
      const DevOp& dv = p.dev[dd[j]];
      kk[j] = dcon + dv.con;
      NP[j] = dv.flags;
      const double vb = V_of(x, dv.node[3]);
      vd[j] = V_of(x, dv.node[0]) - vb; vg[j] = V_of(x, dv.node[1]) - vb; vs[j] = V_of(x, dv.node[2]) - vb;
    }
    ekv_eval_v<K>(kk, NP, vd, vg, vs, c.o->gmin, o0, o1, o2, o3);
#pragma unroll
    for (int j = 0; j < K; j++) {
      if (ok[j]) {
        double* out = dout + p.dev[dd[j]].dout;
        out[0] = o0[j]; out[1] = o1[j]; out[2] = o2[j]; out[3] = o3[j];
      }
    }
  }
}
#endif

// ---- device evaluation: one device per lane ------------------------------------
// Not a template: one copy of the device models per kernel, whatever the variant.
AHK_DEV void eval_devices(Ctx& c, const double* x) {
  const Prog& p = *c.p;
  const Opts& o = *c.o;
  double* dout = c.sm + c.L->dout;
  double* dst = c.sm + c.L->dst;
#ifdef AHK_JIT
  const int G = AHK_JIT_G;
#else
  const int G = c.G;
#endif
#ifdef AHK_EKV_BATCHED
  (void)o; (void)dout; (void)dst;
  eval_devices_ekv<AHK_JIT_ILP>(c, x);
  return;
#endif
  AHK_LANES(d, p.ndev) {
    const DevOp& dv = p.dev[d];
    if (!dv.active) continue;
    double* out = dout + dv.dout;
    double* st = dst + (dv.st >= 0 ? dv.st : 0);
    switch (dv.type) {
      case AHK_D:
        diode_eval(c.sm + c.L->dcon + dv.con, V_of(x, dv.node[0]) - V_of(x, dv.node[1]), o.vea,
                   o.gmin, st, out);
        break;
      case AHK_EKV: {
        double vb = V_of(x, dv.node[3]);
        ekv_eval(c.sm + c.L->dcon + dv.con, dv.flags, V_of(x, dv.node[0]) - vb,
                 V_of(x, dv.node[1]) - vb, V_of(x, dv.node[2]) - vb, o.gmin, out);
      } break;
      case AHK_MOSQ: {
        MosqP m;
        m.VTO = c.PP(dv.pmb + 0); m.KP = c.PP(dv.pmb + 1); m.GAMMA = c.PP(dv.pmb + 2);
        m.PHI = c.PP(dv.pmb + 3); m.LAMBDA = c.PP(dv.pmb + 4); m.AVT = c.PP(dv.pmb + 5);
        m.AKP = c.PP(dv.pmb + 6);
        m.W = c.PP(dv.ppb + 0); m.L = c.PP(dv.ppb + 1); m.M = c.PP(dv.ppb + 2);
        m.N = c.PP(dv.ppb + 3); m.ZVT = c.PP(dv.ppb + 4); m.ZKP = c.PP(dv.ppb + 5);
        m.NP = dv.flags;
        double vs = V_of(x, dv.node[2]);
        mosq_eval(m, V_of(x, dv.node[0]) - vs, V_of(x, dv.node[1]) - vs,
                  V_of(x, dv.node[3]) - vs, o.iea, o.gmin, out);
      } break;
      case AHK_SW: {
        SwP m;
        m.VT = c.PP(dv.pmb + 0); m.VH = c.PP(dv.pmb + 1); m.RON = c.PP(dv.pmb + 2);
        m.ROFF = c.PP(dv.pmb + 3);
        switch_eval(m, V_of(x, dv.node[0]) - V_of(x, dv.node[1]),
                    V_of(x, dv.node[2]) - V_of(x, dv.node[3]), o.gmin, st, out);
      } break;
    }
  }
}


// ---- assembly: [Bv + J | -(Bv.x + T + Tx)], row-local, no conflicts ----------------
// residuo = mna.dot(x) + T + Tx (dc_analysis.py:816); the matrix solved is mna + J.
// One row: writes the pattern entries into the staging row Ar, returns the rhs.
AHK_DEV double assemble_row(Ctx& c, int r, double* Ar) {
  const Prog& p = *c.p;
  const double* x = c.x();
  const double* Bv = c.vb() + c.L->Bv;
  const double* dout = c.sm + c.L->dout;
  const double* Dv = c.vb() + c.L->Dv;
  const bool tran = Dv != Bv;          // transient layout: the matrix is Mv + C1 D (+ Gmin)
  const double C1 = c.C1;
  double s = 0.0;
  const int k1 = p.row_ptr[r + 1];
  int q = p.nl_ptr[p.row_ptr[r]];
  AHK_UNROLL
  for (int k = p.row_ptr[r]; k < k1; k++) {
    double v = Bv[k];
    if (tran) v += C1 * Dv[k];
    const int pc = p.pcd[k];
    const int col = pc & 0x7fff;
    if (pc & 0x8000) v += c.gadd;
    s += v * x[col];
    const int q1 = p.nl_ptr[k + 1];
    AHK_UNROLL
    for (; q < q1; q++) v += (double)p.nl_sgn[q] * dout[p.nl_src[q]];
    Ar[col] = v;
  }
  double tx = 0.0;
  AHK_UNROLL
  for (int t = p.tx_ptr[r]; t < p.tx_ptr[r + 1]; t++) tx += (double)p.tx_sgn[t] * dout[p.tx_src[t]];
  const double rr = s + (c.sm + c.L->T)[r] + tx;
  (c.sm + c.L->res)[r] = rr;
  return -rr;
}

// ---- dense solve, register-resident (Var<NC,R,G>, NC > 0) --------------------------------
// numpy.linalg.solve (LAPACK dgesv) at dc_analysis.py:822: same pivot rule (largest |a|
// in the column among the rows not yet used), but Gauss-Jordan instead of LU +
// back-substitution: every elimination step updates all other rows, so no lane idles
// and there is no serial triangular solve; rows are never moved.  Results differ from
// dgesv by rounding only.  Returns 0 if a pivot is exactly zero (LinAlgError).
// FLAT (warp-uniform kernels): a singular matrix must not leave the function early — the other
// groups of the warp still need this group's lanes in their full-mask shuffles — so the
// elimination runs to the end on garbage and only the return value reports it.
template <class V, bool FLAT = false>
AHK_DEV int gj_solve(Ctx& c, double (&a)[V::R > 0 ? V::R : 1][V::NC > 0 ? V::NC : 1]) {
  const int NC = V::NC, R = V::R, G = V::G;
  const int n = c.p->n;
  int nonsing = 1;
  const unsigned gm = FLAT ? 0xffffffffu : c.mask;   // flat kernels: the literal full mask (unguarded shuffles)
  double* dx = c.sm + c.L->dx;
  double* pvb = c.sm + c.L->pvb;   // [2][NC] pivot-row broadcast buffers
  unsigned usedm = 0;              // bit s: local row slot s has been a pivot row
  double pd[R > 0 ? R : 1];
  int ps[R > 0 ? R : 1];
#pragma unroll
  for (int s = 0; s < R; s++) { pd[s] = 1.0; ps[s] = 0; }   // pd: RECIPROCAL of the row's pivot
#pragma unroll
  for (int k = 0; k < NC - 1; k++) {
    if (k < n) {
      unsigned long long key = 0ull;
#pragma unroll
      for (int s = 0; s < R; s++) {
        const int r = c.sub + G * s;
        const unsigned long long kb = pivot_key(fabs(a[s][k]), r);
        const bool cand = r < n && !((usedm >> s) & 1u);
        if (cand && kb > key) key = kb;
      }
      key = Grp<G>::kmax(key, gm);
      if ((key >> 5) == 0ull) {                  // every candidate is exactly zero
        if (FLAT) nonsing = 0; else return 0;
      }
      const int bi = 31 - (int)(key & 31ull);
      const int own = bi & (G - 1), slot = bi / G;
      const bool mine = c.sub == own;
#ifdef AHK_BCAST_SHFL
      if (G > 1 && R == 1) {
        // one row per lane: every element of the pivot row is fetched from the owner's
        // registers right where it is used — no pivot-row copy in registers, no
        // shared-memory round trip, no barrier
        const double pk = Grp<G>::bcast(a[0][k], own, gm);
        const double ipk = fm::rcp(pk);
        if (mine) { usedm |= 1u; pd[0] = ipk; ps[0] = k; }
        const double l = mine ? 0.0 : a[0][k] * ipk;
#pragma unroll
        for (int j = k + 1; j < NC; j++) a[0][j] -= l * Grp<G>::bcast(a[0][j], own, gm);
#else
      if (G > 1 && R == 1) {
        // one row per lane: the owner publishes its row in one of two shared-memory buffers,
        // one group barrier, then every lane reads each element right where it is used (one
        // broadcast load per element instead of two shuffles and their run-time-mask checks;
        // no pivot-row copy in registers).  Two buffers: a lane reaches the barrier of step
        // k+1 only after its reads of step k, so buffer k & 1 is free again at step k+2.
        double* buf = pvb + (k & 1) * NC;
        if (mine) {
#pragma unroll
          for (int j = k; j < NC; j++) buf[j] = a[0][j];
        }
        Grp<G>::sync(gm);
        const double ipk = fm::rcp(buf[k]);
        if (mine) { usedm |= 1u; pd[0] = ipk; ps[0] = k; }
        const double l = mine ? 0.0 : a[0][k] * ipk;
#pragma unroll
        for (int j = k + 1; j < NC; j++) a[0][j] -= l * buf[j];
#endif
      } else {
        double pv[NC > 0 ? NC : 1];
        if (G > 1) {   // the owner lane publishes its row, everybody reads it back
          double* buf = pvb + (k & 1) * NC;
          if (mine) {
#pragma unroll
            for (int s = 0; s < R; s++)
              if (slot == s) {
#pragma unroll
                for (int j = k; j < NC; j++) buf[j] = a[s][j];
              }
          }
          Grp<G>::sync(gm);
#pragma unroll
          for (int j = k; j < NC; j++) pv[j] = buf[j];
        } else {
#pragma unroll
          for (int j = k; j < NC; j++) {
            double t = a[0][j];
#pragma unroll
            for (int s = 1; s < R; s++) if (slot == s) t = a[s][j];
            pv[j] = t;
          }
        }
        const double inv = fm::rcp(pv[k]);
        if (mine) {
          usedm |= 1u << slot;
#pragma unroll
          for (int s = 0; s < R; s++) if (slot == s) { pd[s] = inv; ps[s] = k; }
        }
#pragma unroll
        for (int s = 0; s < R; s++) {
          const double l = (mine && slot == s) ? 0.0 : a[s][k] * inv;
#pragma unroll
          for (int j = k + 1; j < NC; j++) a[s][j] -= l * pv[j];
        }
      }
    }
  }
#pragma unroll
  for (int s = 0; s < R; s++) {
    const int r = c.sub + G * s;
    if (r < n) dx[ps[s]] = a[s][NC - 1] * pd[s];   // (stored pivot reciprocal: no division)
  }
  Grp<G>::sync(gm);
  return nonsing;
}

// ---- dense solve, one lane per instance (Var<NC,R,1>): LU in registers -------------------
// dgesv as LAPACK's unblocked dgetf2 + dgetrs run it: at step k the largest |a| of column k
// among rows k.. (first maximum wins, exact comparison), an explicit row swap (register
// selects, the indices stay compile-time), multipliers l = a * (1 / pivot), elimination of
// the rows below only, then back-substitution (with the stored pivot reciprocals).  About a
// third of the instructions of the Gauss-Jordan form below, which pays for keeping every
// lane of a group busy — pointless when the group is one lane.  0 = a pivot is exactly zero.
template <class V, bool REC = false>
AHK_DEV int lu_solve_thread(Ctx& c, double (&a)[V::R > 0 ? V::R : 1][V::NC > 0 ? V::NC : 1], int* rec = nullptr) {
  const int NC = V::NC, R = V::R;
  const int n = c.p->n;
  double* dx = c.sm + c.L->dx;
  double inv[R > 0 ? R : 1], xk[R > 0 ? R : 1];
#pragma unroll
  for (int k = 0; k < R; k++) { inv[k] = 0.0; xk[k] = 0.0; }
#pragma unroll
  for (int k = 0; k < R; k++) {
    if (k < n) {
      double best = fabs(a[k][k]);
      int p = k;
#pragma unroll
      for (int r = k + 1; r < R; r++) {
        const double v = fabs(a[r][k]);
        if (r < n && v > best) { best = v; p = r; }
      }
      if (best == 0.0) return 0;
      if (REC) { if (rec) rec[k] = p; }
#pragma unroll
      for (int j = k; j < NC; j++) {        // swap rows k and p (columns < k are not needed again)
        const double t = a[k][j];
        double q = t;
#pragma unroll
        for (int r = k + 1; r < R; r++) if (p == r) q = a[r][j];
        a[k][j] = q;
#pragma unroll
        for (int r = k + 1; r < R; r++) if (p == r) a[r][j] = t;
      }
      const double iv = fm::rcp(a[k][k]);
      inv[k] = iv;
#pragma unroll
      for (int r = k + 1; r < R; r++) {     // rows >= n are zero: l = 0
        const double l = a[r][k] * iv;
#pragma unroll
        for (int j = k + 1; j < NC; j++) a[r][j] -= l * a[k][j];
      }
    }
  }
#pragma unroll
  for (int k = R - 1; k >= 0; k--) {
    if (k < n) {
      double s = a[k][NC - 1];
#pragma unroll
      for (int j = k + 1; j < R; j++) s -= a[k][j] * xk[j];   // columns >= n are zero
      xk[k] = s * inv[k];
      dx[k] = xk[k];
    }
  }
  return 1;
}

#ifdef AHK_JIT_SLU
// ---- the same LU along a STATIC schedule (specialised kernels, one lane per instance or
// replicated rows) -------------------------------------------------------------------------
// The circuit fixes the sparsity pattern of mna + J, and in practice the operating region
// fixes dgesv's pivot sequence too (C3: one sequence in 11 525 solves of 16 instances, C4: one in
// 102 608).  jit_gen.hpp (slu_codegen) eliminates symbolically along the sequence the engine
// learned in its calibration launch and emits the elimination as straight-line statements
// on literal indices (AHK_JIT_SLU_BODY): no pivot search, no row swaps, only the structural
// non-zeros are touched (C3: 43 multiply-adds + 9 reciprocals instead of a dense 9 x 10
// elimination).  Every step VERIFIES that the scheduled pivot is the one dgesv would take
// (largest magnitude among the rows not yet used, first one in LAPACK's swapped row order on
// ties); skipped operations are on exact zeros, so while the pivots agree the result has the
// bits of lu_solve_thread.  A solve whose pivots differ (`same` false, or a zero pivot) is
// assembled again and goes through lu_solve_thread.
template <class V>
AHK_DEV int slu_solve_thread(Ctx& c, double (&a)[V::R > 0 ? V::R : 1][V::NC > 0 ? V::NC : 1]) {
  double* dx = c.sm + c.L->dx;
  bool same = true;
  AHK_JIT_SLU_BODY
  return same ? 1 : 0;
}
#endif

// ---- dense solve, generic fallback (Var<0,0,G>): matrix in shared memory ----------------
// Gaussian elimination with implicit partial pivoting + back-substitution, run-time n.
template <int G, bool FLAT = false>
AHK_DEV int lu_solve_smem(Ctx& c) {
  const int n = c.p->n, RS = c.L->RS;
  int nonsing = 1;
  const unsigned gm = FLAT ? 0xffffffffu : c.mask;
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
      if (bi < 0 || v > best) { best = v; bi = r; }
    }
    Grp<G>::argmax(best, bi, gm);
    if (best == 0.0) {
      if (!FLAT) return 0;
      nonsing = 0;              // (flat kernels run on, on garbage, to keep the warp together)
    }
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
    Grp<G>::sync(gm);
  }
  for (int k = n - 1; k >= 0; k--) {
    int pr = prow[k];
    double xk = A[pr * RS + n] / A[pr * RS + k];
    if (c.sub == 0) dx[k] = xk;
    for (int r = c.sub; r < n; r += G) {
      if ((int)stepof[r] < k) { double a = A[r * RS + k]; if (a != 0.0) A[r * RS + n] -= a * xk; }
    }
    Grp<G>::sync(gm);
  }
  return nonsing;
}

// ---- dense solve, large n (AHK_SMALL_NMAX < n <= AHK_BIG_NMAX): matrix in global memory ------
// numpy.linalg.solve at dc_analysis.py:822 for circuits of any size.  The same elimination as
// lu_solve_smem / dgesv (partial pivoting: largest magnitude among the rows not yet used, ties
// to the lowest row; rows are never moved) on a dense n x RS matrix that lives in L2 / HBM —
// a warp cannot afford to walk 2 KB-strided columns and full rows of zeros there (measured: a
// 252-unknown diode ladder spent its time pulling one DRAM sector per matrix element), so the
// non-zero STRUCTURE is tracked next to the values: one bitmap of non-zero columns per row, in
// shared memory, set from the stamp pattern at assembly and OR-ed with the pivot row's bitmap
// at every row update (fill-in).  Column k is read only in the rows whose bit k is set; a row
// update touches only the columns of the pivot row's bitmap (sparse pivot row: the lanes split
// the (row, bitmap word) pairs; dense pivot row: the lanes run along the columns, coalesced);
// the next assembly clears only the flagged entries.  Skipped entries are exact zeros and
// a - 0*b == a, so the result is that of the dense algorithm.  The multiplier column, the
// right-hand side and the pivot bookkeeping stay in shared memory.  One warp per instance on
// the device (G == 32), one lane on the host (tests).
AHK_DEV int big_ctz(unsigned v) {
#ifdef __CUDACC__
  return __ffs((int)v) - 1;
#else
  return __builtin_ctz(v);
#endif
}
AHK_DEV int big_popc(unsigned v) {
#ifdef __CUDACC__
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}

AHK_DEV void big_or(unsigned* p, unsigned v) {     // lanes of one warp may hit the same word
#ifdef __CUDACC__
  atomicOr(p, v);
#else
  *p |= v;
#endif
}

// once per instance: the whole matrix zero, no entry flagged
template <int G>
AHK_DEV void big_init(Ctx& c, double* A, size_t len) {
  const int n = c.p->n, W = (n + 31) >> 5;
  unsigned* nzm = (unsigned*)(c.sm + c.L->nzm);
  for (size_t i = c.sub; i < len; i += G) A[i] = 0.0;
  for (int i = c.sub; i < n * W; i += G) nzm[i] = 0u;
  Grp<G>::sync(c.mask);
}

// One panel of BIG_NB pivot steps of lu_solve_glob on a DENSE remainder, blocked (right-looking):
//   1. per column of the panel: pivot search (dgesv's rule), multipliers of ALL unused rows into
//      the shared-memory panel Lp[row][BIG_NB], eager update of the remaining PANEL columns and
//      of the right-hand side only (the next pivot search needs them);
//   2. the pivot rows' trailing parts: U12 = L11^-1 A12, a unit-lower-triangular solve with the
//      lanes along the columns (registers);
//   3. trailing update A22 -= L21 U12: 8 x 8 tiles of C over K = 8, two mma.sync m8n8k4 FP64
//      (DMMA) per tile — every element of the trailing matrix is read and written ONCE per panel
//      instead of once per pivot step.
// The operations on an element are those of the unblocked elimination in another order of
// summation (rounding-level differences); the pivot sequence is the same.
#define BIG_NB 8
template <int G, bool FLAT>
AHK_DEV int lu_panel_dense(Ctx& c, const int k0) {
  const int n = c.p->n, RS = c.L->RS, W = (n + 31) >> 5, NB = BIG_NB;
  const unsigned gm = FLAT ? 0xffffffffu : c.mask;
  int nonsing = 1;
  double* A = c.Ag;
  double* dx = c.sm + c.L->dx;
  double* b = c.sm + c.L->brhs;
  double* Lp = c.Ag + c.L->lpan;       // [n][NB] multipliers of this panel (global workspace, L1 / L2 resident)
  unsigned* nzm = (unsigned*)(c.sm + c.L->nzm);
  unsigned short* prow = (unsigned short*)(c.sm + c.L->bytes);
  unsigned short* stepof = prow + n;
  unsigned short* list = stepof + n;
  const int kend = k0 + NB;
  for (int kk = 0; kk < NB; kk++) {
    const int k = k0 + kk;
    double best = -1.0;
    int bi = -1;
    for (int r = c.sub; r < n; r += G) {
      if (stepof[r] != 0xffff) continue;
      const double a = A[(size_t)r * RS + k];
      dx[r] = a;
      const double v = fabs(a);
      if (bi < 0 || v > best) { best = v; bi = r; }
    }
    Grp<G>::argmax(best, bi, gm);
    if (best == 0.0) {
      if (!FLAT) return 0;
      nonsing = 0;
    }
    if (c.sub == 0) { prow[k] = (unsigned short)bi; stepof[bi] = (unsigned short)k; }
    Grp<G>::sync(gm);
    const double inv = 1.0 / dx[bi];
    const double* Ap = A + (size_t)bi * RS;
    const double bp = b[bi];
    double pv[BIG_NB];                  // the pivot row inside the panel (every lane: broadcast loads)
#pragma unroll
    for (int j = 0; j < NB; j++) pv[j] = (j > kk) ? Ap[k0 + j] : 0.0;
    for (int r = c.sub; r < n; r += G) {
      if (stepof[r] != 0xffff) continue;
      const double l = dx[r] * inv;
      Lp[r * NB + kk] = l;
      if (l != 0.0) {
        double* Ar = A + (size_t)r * RS + k0;
#pragma unroll
        for (int j = 0; j < NB; j++) if (j > kk) Ar[j] -= l * pv[j];
        b[r] -= l * bp;
      }
    }
    Grp<G>::sync(gm);
  }
  // 2. U12: rows prow[k0 .. kend), columns >= kend
  for (int j = kend + c.sub; j < n; j += G) {
    double u[BIG_NB];
#pragma unroll
    for (int i = 0; i < NB; i++) {
      const int pi = prow[k0 + i];
      double v = A[(size_t)pi * RS + j];
#pragma unroll
      for (int m = 0; m < NB; m++) if (m < i) v -= Lp[pi * NB + m] * u[m];
      u[i] = v;
      if (i > 0) A[(size_t)pi * RS + j] = v;
    }
  }
  // rows of the trailing update: unused, some multiplier of the panel not zero
  int cnt = 0;
  for (int r0 = 0; r0 < n; r0 += G) {
    const int r = r0 + c.sub;
    bool nz = false;
    if (r < n && stepof[r] == 0xffff) {
#pragma unroll
      for (int m = 0; m < NB; m++) nz = nz || (Lp[r * NB + m] != 0.0);
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
  Grp<G>::sync(gm);                     // U12 and the list are visible
  // 3. A22 -= L21 U12
#ifdef __CUDACC__
  if (G == 32) {
    const int g = c.sub >> 2, tig = c.sub & 3;
    const double* U0 = A + (size_t)prow[k0 + tig] * RS;          // B operands: U[tig][*], U[4 + tig][*]
    const double* U1 = A + (size_t)prow[k0 + 4 + tig] * RS;
    for (int j0 = kend; j0 < n; j0 += 8) {
      const int jb = j0 + g;
      const double b0 = jb < n ? U0[jb] : 0.0, b1 = jb < n ? U1[jb] : 0.0;
      const int jc = j0 + 2 * tig;
      // four row tiles per pass: all their loads are in flight before the first MMA needs one
      // (a warp is alone with its instance: memory-level parallelism is what hides the L2 latency)
      for (int t0 = 0; t0 < cnt; t0 += 32) {
        double a0[4], a1[4], c0[4], c1[4];
        double* Cr[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int t = t0 + 8 * q + g;
          const int r = t < cnt ? (int)list[t] : -1;
          a0[q] = a1[q] = c0[q] = c1[q] = 0.0;
          Cr[q] = nullptr;
          if (r >= 0) {
            a0[q] = -Lp[r * NB + tig]; a1[q] = -Lp[r * NB + 4 + tig];
            Cr[q] = A + (size_t)r * RS + jc;
            if (jc + 1 < n) { const double2 cc = *(const double2*)Cr[q]; c0[q] = cc.x; c1[q] = cc.y; }
            else if (jc < n) c0[q] = Cr[q][0];
          }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          if (t0 + 8 * q < cnt) {        // (warp-uniform)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[q]), "+d"(c1[q]) : "d"(a0[q]), "d"(b0));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[q]), "+d"(c1[q]) : "d"(a1[q]), "d"(b1));
          }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          if (Cr[q]) {
            if (jc + 1 < n) *(double2*)Cr[q] = make_double2(c0[q], c1[q]);
            else if (jc < n) Cr[q][0] = c0[q];
          }
        }
      }
    }
  } else
#endif
  {
    for (int t = 0; t < cnt; t++) {
      const int r = list[t];
      double* Ar = A + (size_t)r * RS;
      for (int j = kend + c.sub; j < n; j += G) {
        double v = Ar[j];
        for (int m = 0; m < NB; m++) v -= Lp[r * NB + m] * A[(size_t)prow[k0 + m] * RS + j];
        Ar[j] = v;
      }
    }
  }
  // the rows touched are dense right of the panel from here on (and keep their panel columns)
  {
    const int w0 = k0 >> 5;
    const unsigned m0 = ~0u << (k0 & 31);
    for (int item = c.sub; item < (cnt + NB) * (W - w0); item += G) {
      const int t = item / (W - w0), w = w0 + item % (W - w0);
      const int r = t < cnt ? (int)list[t] : (int)prow[k0 + t - cnt];
      unsigned full = (w == W - 1 && (n & 31)) ? ((1u << (n & 31)) - 1u) : ~0u;
      if (w == w0) full &= m0;
      nzm[r * W + w] |= full;
    }
  }
  Grp<G>::sync(gm);
  return nonsing;
}

template <int G, bool FLAT = false>
AHK_DEV int lu_solve_glob(Ctx& c) {
  const int n = c.p->n, RS = c.L->RS, W = (n + 31) >> 5;
  const unsigned gm = FLAT ? 0xffffffffu : c.mask;
  int nonsing = 1;
  double* A = c.Ag;
  double* dx = c.sm + c.L->dx;          // column k / multipliers while eliminating, then the solution
  double* b = c.sm + c.L->brhs;
  unsigned* nzm = (unsigned*)(c.sm + c.L->nzm);
  unsigned short* prow = (unsigned short*)(c.sm + c.L->bytes);
  unsigned short* stepof = prow + n;    // 0xffff: row not used as a pivot row yet
  unsigned short* list = stepof + n;
  // The same structure by COLUMNS (czm[col]: bitmap of the rows with an entry in it) and a
  // bitmap of the used rows give the candidates of a pivot step directly — ncu on the
  // 252-unknown ladder: scanning all n rows per step for the pivot and for the multipliers was
  // 1065 warp instructions per pivot step on a lone warp (7 cycles each), 60 % of the kernel.
  // Maintained while the elimination is sparse; after the first dense panel (which does not
  // maintain it) the steps fall back to the row scan (`densemode`).
  unsigned* czm = (unsigned*)(c.sm + (c.L->czm < 0 ? 0 : c.L->czm));
  unsigned* usedw = (unsigned*)((char*)prow + ((6 * n + 3) & ~3));   // behind the three 16-bit arrays, 4-byte aligned
  bool densemode = c.L->czm < 0;        // dense stamp pattern: no column bitmaps at all
  for (int r = c.sub; r < n; r += G) stepof[r] = 0xffff;
  for (int w = c.sub; w < W; w += G) usedw[w] = 0u;
  Grp<G>::sync(gm);
  for (int k = 0; k < n; k++) {
    if ((k & (BIG_NB - 1)) == 0 && k + BIG_NB + 8 <= n) {
      // A full panel of BIG_NB columns ahead: where the remaining matrix is DENSE (more than half
      // of the entries of the unused rows right of k are flagged — dense circuits, or fill-in late
      // in a sparse elimination) the panel is eliminated as a block: the trailing update is a
      // genuine dense contraction, C -= L U with K = BIG_NB, done on the FP64 tensor cores.
      double cntb = 0.0, cntr = 0.0;
      const int w0 = k >> 5;
      for (int r = c.sub; r < n; r += G) {
        if (stepof[r] != 0xffff) continue;
        cntr += 1.0;
        for (int w = w0; w < W; w++) cntb += (double)big_popc(nzm[r * W + w] & (w == w0 ? ~0u << (k & 31) : ~0u));
      }
      cntb = Grp<G>::sum_all(cntb, gm);
      cntr = Grp<G>::sum_all(cntr, gm);
      if (cntb * 2.0 > cntr * (double)(n - k)) {
        const int ok = lu_panel_dense<G, FLAT>(c, k);
        if (!ok) { if (!FLAT) return 0; nonsing = 0; }
        k += BIG_NB - 1;
        densemode = true;
        continue;
      }
    }
    double best = -1.0;
    int bi = -1;
    int cnt = 0;
    if (!densemode) {
      // ---- sparse step: the candidates are the flagged, unused rows of column k ----
      const unsigned* cz = czm + k * W;
      for (int w = c.sub; w < W; w += G) {
        unsigned bits = cz[w] & ~usedw[w];
        while (bits) {
          const int r = (w << 5) + big_ctz(bits);
          bits &= bits - 1;
          const double a = A[(size_t)r * RS + k];
          dx[r] = a;
          const double v = fabs(a);
          if (bi < 0 || v > best) { best = v; bi = r; }   // (ascending rows: the first maximum stays)
        }
      }
      Grp<G>::argmax(best, bi, gm);
      if (bi < 0 || best == 0.0) {       // no entry / only zeros in the column: singular
        if (!FLAT) return 0;
        nonsing = 0;
        if (bi < 0) {                    // (flat kernels run on: any unused row, the result is discarded)
          for (int w = 0; w < W && bi < 0; w++) {
            unsigned u = ~usedw[w];
            if (w == W - 1 && (n & 31)) u &= (1u << (n & 31)) - 1u;
            if (u) bi = (w << 5) + big_ctz(u);
          }
          if (c.sub == 0) dx[bi] = 1.0;
        }
      }
      if (c.sub == 0) {
        prow[k] = (unsigned short)bi; stepof[bi] = (unsigned short)k;
        usedw[bi >> 5] |= 1u << (bi & 31);
      }
      Grp<G>::sync(gm);
      const double inv0 = 1.0 / dx[bi];
      // rows of this step: candidates whose multiplier is not zero, in ascending order
#ifdef __CUDACC__
      if (G == 32) {               // W <= 16 words: one word per lane, offsets by a warp scan
        const int w = c.sub;
        unsigned nzb = 0u;
        if (w < W) {
          unsigned bits = cz[w] & ~usedw[w];
          while (bits) {
            const int bq = big_ctz(bits);
            bits &= bits - 1;
            const int r = (w << 5) + bq;
            const double l = dx[r] * inv0;
            if (l != 0.0) { dx[r] = l; nzb |= 1u << bq; }
          }
        }
        const int my = __popc(nzb);
        int off = my;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, off, o);
          if (c.sub >= o) off += t;
        }
        cnt = __shfl_sync(0xffffffffu, off, 31);
        off -= my;
        while (nzb) { list[off++] = (unsigned short)((w << 5) + big_ctz(nzb)); nzb &= nzb - 1; }
      } else
#endif
      {
        for (int w = 0; w < W; w++) {
          unsigned bits = cz[w] & ~usedw[w];
          while (bits) {
            const int r = (w << 5) + big_ctz(bits);
            bits &= bits - 1;
            const double l = dx[r] * inv0;
            if (l != 0.0) { dx[r] = l; list[cnt++] = (unsigned short)r; }
          }
        }
      }
      Grp<G>::sync(gm);
    } else {
      // ---- after a dense panel: scan the rows ----
      const int kw = k >> 5;
      const unsigned kb = 1u << (k & 31);
      for (int r = c.sub; r < n; r += G) {
        if (stepof[r] != 0xffff) continue;
        const double a = (nzm[r * W + kw] & kb) ? A[(size_t)r * RS + k] : 0.0;
        dx[r] = a;
        const double v = fabs(a);
        if (bi < 0 || v > best) { best = v; bi = r; }
      }
      Grp<G>::argmax(best, bi, gm);
      if (best == 0.0) {
        if (!FLAT) return 0;
        nonsing = 0;
      }
      if (c.sub == 0) { prow[k] = (unsigned short)bi; stepof[bi] = (unsigned short)k; }
      Grp<G>::sync(gm);
      const double inv = 1.0 / dx[bi];
      for (int r0 = 0; r0 < n; r0 += G) {
        const int r = r0 + c.sub;
        bool nz = false;
        if (r < n && stepof[r] == 0xffff) {
          const double l = dx[r] * inv;
          nz = l != 0.0;
          if (nz) dx[r] = l;
        }
#ifdef __CUDACC__
        if (G == 32) {
          const unsigned bal = __ballot_sync(0xffffffffu, nz);
          if (nz) list[cnt + __popc(bal & ((1u << c.sub) - 1u))] = (unsigned short)r;
          cnt += __popc(bal);
        } else
#endif
        { if (nz) list[cnt++] = (unsigned short)r; }     // one lane per instance
      }
      Grp<G>::sync(gm);
    }
    const double* Ap = A + (size_t)bi * RS;
    const unsigned* nzp = nzm + bi * W;
    const double bp = b[bi];
    const int w0 = (k + 1) >> 5, Wrem = W - w0;
    const unsigned m0 = ~0u << ((k + 1) & 31);          // columns > k of word w0
    int nnzp = 0;                                         // entries of the pivot row right of the pivot
    for (int w = w0; w < W; w++) nnzp += big_popc(nzp[w] & (w == w0 ? m0 : ~0u));
    if (k + 1 < n && cnt > 0) {
      if (nnzp * 4 > n - k - 1) {
        // dense pivot row: lanes along the columns (coalesced), one row after the other
        for (int t = 0; t < cnt; t++) {
          const int r = list[t];
          const double l = dx[r];
          double* Ar = A + (size_t)r * RS;
#pragma unroll 4
          for (int j = k + 1 + c.sub; j < n; j += G) Ar[j] -= l * Ap[j];
          for (int w = w0 + c.sub; w < W; w += G) {
            const unsigned bits = nzp[w] & (w == w0 ? m0 : ~0u), old = nzm[r * W + w];
            nzm[r * W + w] = old | bits;
            unsigned nb = densemode ? 0u : bits & ~old;      // fill-in: the column bitmaps follow
            while (nb) {
              const int j = (w << 5) + big_ctz(nb);
              nb &= nb - 1;
              big_or(&czm[j * W + (r >> 5)], 1u << (r & 31));
            }
          }
        }
      } else {
        // sparse pivot row: the lanes split the (row, bitmap word) pairs
        for (int item = c.sub; item < cnt * Wrem; item += G) {
          const int t = item / Wrem, w = w0 + item % Wrem;
          unsigned bits = nzp[w] & (w == w0 ? m0 : ~0u);
          if (!bits) continue;
          const int r = list[t];
          const double l = dx[r];
          double* Ar = A + (size_t)r * RS;
          const unsigned old = nzm[r * W + w];
          nzm[r * W + w] = old | bits;
          unsigned nb = densemode ? 0u : bits & ~old;        // fill-in: the column bitmaps follow
          while (nb) {
            const int j = (w << 5) + big_ctz(nb);
            nb &= nb - 1;
            big_or(&czm[j * W + (r >> 5)], 1u << (r & 31));
          }
          while (bits) {
            const int j = (w << 5) + big_ctz(bits);
            bits &= bits - 1;
            Ar[j] -= l * Ap[j];
          }
        }
      }
    }
    for (int t = c.sub; t < cnt; t += G) { const int r = list[t]; b[r] -= dx[r] * bp; }
    Grp<G>::sync(gm);
  }
  // back-substitution, row-wise: dx[k] = (b[p] - sum_{j > k} U[p][j] dx[j]) / U[p][k]
  for (int k = n - 1; k >= 0; k--) {
    const int pr = prow[k];
    const double* Ap = A + (size_t)pr * RS;
    const unsigned* nzp = nzm + pr * W;
    const int w0 = (k + 1) >> 5;
    const unsigned m0 = ~0u << ((k + 1) & 31);
    double s = 0.0;
    if (k + 1 < n) {
      for (int w = w0 + c.sub; w < W; w += G) {
        unsigned bits = nzp[w] & (w == w0 ? m0 : ~0u);
        while (bits) {
          const int j = (w << 5) + big_ctz(bits);
          bits &= bits - 1;
          s += Ap[j] * dx[j];
        }
      }
    }
    s = Grp<G>::sum_all(s, gm);
    if (c.sub == 0) dx[k] = (b[pr] - s) / Ap[k];
    Grp<G>::sync(gm);
  }
  return nonsing;
}

#ifdef AHK_JIT
// Specialised build: the row is a COMPILE-TIME constant, so its column indices, gather
// lists and signs fold into straight-line code that writes the matrix row directly into
// the registers of the solve (no staging row in shared memory).  With several lanes per
// instance the lane's row is only known at run time: a switch runs each row's code for
// the lanes that own it (divergent, but a few dozen instructions per row, and no index
// loads, loop control or shared-memory round trip).
template <int ROW, int NC>
AHK_DEV double assemble_row_const(Ctx& c, double (&row)[NC]) {
#pragma unroll
  for (int j = 0; j < NC - 1; j++) row[j] = 0.0;
  return assemble_row(c, ROW, row);
}
template <int NC>
AHK_DEV double assemble_row_switch(Ctx& c, int r, double (&row)[NC]) {
  switch (r) {
#define AHK_ROW_CASE(i) case i: if constexpr (i < Prog::n) return assemble_row_const<i, NC>(c, row); else break;
    AHK_ROW_CASE(0) AHK_ROW_CASE(1) AHK_ROW_CASE(2) AHK_ROW_CASE(3) AHK_ROW_CASE(4) AHK_ROW_CASE(5)
    AHK_ROW_CASE(6) AHK_ROW_CASE(7) AHK_ROW_CASE(8) AHK_ROW_CASE(9) AHK_ROW_CASE(10) AHK_ROW_CASE(11)
    AHK_ROW_CASE(12) AHK_ROW_CASE(13) AHK_ROW_CASE(14) AHK_ROW_CASE(15) AHK_ROW_CASE(16) AHK_ROW_CASE(17)
    AHK_ROW_CASE(18) AHK_ROW_CASE(19) AHK_ROW_CASE(20) AHK_ROW_CASE(21) AHK_ROW_CASE(22) AHK_ROW_CASE(23)
    AHK_ROW_CASE(24) AHK_ROW_CASE(25) AHK_ROW_CASE(26) AHK_ROW_CASE(27) AHK_ROW_CASE(28) AHK_ROW_CASE(29)
    AHK_ROW_CASE(30)
#undef AHK_ROW_CASE
  }
#pragma unroll
  for (int j = 0; j < NC; j++) row[j] = 0.0;
  return 0.0;
}
#endif

// Calibration record of the static-schedule kernels (engine.cu: slu_calibrate), written by ONE
// thread (instance 0): rec[0] solves that left the schedule, rec[1] solves, rec[2] distinct
// sequences seen among them (at most 4), rec[8 + i] how often, rec[16 + 32 i + k] sequence i.
AHK_DEV void slu_record(int* rec, const int* seq, int n) {
  rec[0]++;
  const int ns = rec[2];
  for (int i = 0; i < ns; i++) {
    bool eq = true;
    for (int k = 0; k < n; k++) if (rec[16 + 32 * i + k] != seq[k]) eq = false;
    if (eq) { rec[8 + i]++; return; }
  }
  if (ns < 4) {
    for (int k = 0; k < n; k++) rec[16 + 32 * ns + k] = seq[k];
    rec[8 + ns] = 1; rec[2] = ns + 1;
  }
}

#ifdef AHK_JIT_SLU
// a solve that left the static schedule (or is singular): the matrix again, through the pivoting
// elimination.  Not inlined: its dense register matrix must not weigh on the scheduled path.
template <class V>
#ifdef __CUDACC__
__device__ __noinline__
#else
static
#endif
int slu_fallback(Ctx& c, int* rec) {
  const int NC = V::NC, R = V::R, n = c.p->n;
  double a[R > 0 ? R : 1][NC > 0 ? NC : 1];
#pragma unroll
  for (int s = 0; s < R; s++) {
    if (s < n) {
#pragma unroll
      for (int j = 0; j < NC - 1; j++) a[s][j] = 0.0;
      a[s][NC - 1] = assemble_row(c, s, a[s]);
    } else {
#pragma unroll
      for (int j = 0; j < NC; j++) a[s][j] = 0.0;
    }
  }
  return lu_solve_thread<V, true>(c, a, rec);
}
#endif

// one Newton linear step: assemble at x() and solve into dx(); 0 = singular
template <class V, bool FLAT = false>
AHK_DEV int assemble_solve(Ctx& c) {
  const int n = c.p->n, RS = c.L->RS, G = V::G;
  double* A = c.A();
  if constexpr (V::NC > 0) {
    const int NC = V::NC, R = V::R;
    double a[R > 0 ? R : 1][NC > 0 ? NC : 1];
    // REPLICATED solve: several lanes per instance (they share the device evaluations), but
    // every lane holds ALL rows and runs the one-lane LU redundantly — for a small matrix
    // behind expensive devices that beats distributing rows (no pivot search across lanes, no
    // pivot-row broadcast, no idle lanes when n is not a multiple of G)
    constexpr bool REP = RepOf<V>::value;
#pragma unroll
    for (int s = 0; s < R; s++) {
      const int r = REP ? s : c.sub + G * s;
#ifdef AHK_JIT
      (void)A; (void)RS;
      if (G == 1 || REP) {   // r == s: a compile-time constant
        if (s < n) {
#pragma unroll
          for (int j = 0; j < NC - 1; j++) a[s][j] = 0.0;
          a[s][NC - 1] = assemble_row(c, s, a[s]);
        }
      } else if (r < n) {
        a[s][NC - 1] = assemble_row_switch<NC>(c, r, a[s]);
      }
      if ((G == 1 || REP) ? s >= n : r >= n) {
#pragma unroll
        for (int j = 0; j < NC; j++) a[s][j] = 0.0;
      }
#else
      if (r < n) {
        // R == 1: the lane's own staging row, zeroed once (only the pattern is rewritten);
        // R > 1: ONE scratch row per lane, cleared before each of its rows
        double* Ar = A + (R == 1 ? r : c.sub) * RS;
        if (R > 1) {
#pragma unroll
          for (int j = 0; j < NC - 1; j++) Ar[j] = 0.0;
        }
        a[s][NC - 1] = assemble_row(c, r, Ar);
#pragma unroll
        for (int j = 0; j < NC - 1; j++) a[s][j] = Ar[j];
      } else {
#pragma unroll
        for (int j = 0; j < NC; j++) a[s][j] = 0.0;
      }
#endif
    }
    if constexpr (V::G == 1 || REP) {
#ifdef AHK_JIT_SLU
      int okk = slu_solve_thread<V>(c, a);
      // calibration launch (engine.cu): the first 32 instances each count their solves and the
      // ones that left the schedule, in their own record (pivots may depend on the instance)
      int* prec = (c.pivrec != nullptr && c.inst < 32 && c.sub == 0) ? c.pivrec + 256 * (int)c.inst : nullptr;
      if (prec) prec[1]++;
      if (!okk) {
        int rec[R > 0 ? R : 1];
        okk = slu_fallback<V>(c, rec);
        if (prec && okk) slu_record(prec, rec, n);   // ... and leave the sequence dgesv took
      }
#else
      const int okk = lu_solve_thread<V>(c, a);   // (REP: every lane stores the same dx)
#endif
      Grp<G>::sync(FLAT ? 0xffffffffu : c.mask);
      return okk;
    } else return gj_solve<V, FLAT>(c, a);
  } else {
    if (c.L->big) {       // large n: matrix in the global workspace, right-hand side in shared memory
      const int W = (n + 31) >> 5;
      unsigned* nzm = (unsigned*)(c.sm + c.L->nzm);
      // clear what the previous factorisation left behind: the flagged entries only
      for (int item = c.sub; item < n * W; item += G) {
        unsigned bits = nzm[item];
        if (!bits) continue;
        nzm[item] = 0u;
        double* Ar = A + (size_t)(item / W) * RS + ((item % W) << 5);
        while (bits) { Ar[big_ctz(bits)] = 0.0; bits &= bits - 1; }
      }
      const bool cz = c.L->czm >= 0;
      unsigned* czm = (unsigned*)(c.sm + (cz ? c.L->czm : 0));
      if (cz) for (int i = c.sub; i < n * W; i += G) czm[i] = 0u;
      Grp<G>::sync(FLAT ? 0xffffffffu : c.mask);
      double* b = c.sm + c.L->brhs;
      const Prog& p = *c.p;
      for (int r = c.sub; r < n; r += G) {
        b[r] = assemble_row(c, r, A + (size_t)r * RS);
        for (int k = p.row_ptr[r]; k < p.row_ptr[r + 1]; k++) {     // the stamp pattern of the row
          const int col = p.pcol[k];
          nzm[r * W + (col >> 5)] |= 1u << (col & 31);
          if (cz) big_or(&czm[col * W + (r >> 5)], 1u << (r & 31));
        }
      }
      Grp<G>::sync(FLAT ? 0xffffffffu : c.mask);
      return lu_solve_glob<G, FLAT>(c);
    }
    for (int r = c.sub; r < n; r += G) {
      double* Ar = A + r * RS;
      for (int j = 0; j < n; j++) Ar[j] = 0.0;
      Ar[n] = assemble_row(c, r, Ar);
    }
    Grp<G>::sync(FLAT ? 0xffffffffu : c.mask);
    return lu_solve_smem<G, FLAT>(c);
  }
}

// ---- damped Newton (dc_analysis.mdn_solver) -------------------------------------------
// returns 1 converged, 0 not converged, -1 singular (x untouched in that iteration)
template <class V>
AHK_DEV int mdn_solver(Ctx& c, int MAXIT, int* iters) {
  const int G = V::G;
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
    if (p.ndev) { eval_devices(c, x); Grp<G>::sync(c.mask); }
    if (!assemble_solve<V>(c)) { c.singular = 1; *iters = 0; return -1; }
    // get_td (dc_analysis.py:884-946)
    double td = 1.0;
    if (o.nr_damp_first_iters) td = it < 10 ? 1e-2 : (it < 20 ? 0.1 : 1.0);
    if (o.nl_voltages_lock && p.nlock) {
      // min over the pairs of lim / d  ==  lim / (max over the pairs of d): the quotient is
      // monotone in d, so ONE division per iteration gives the identical value
      double dmax = 0.0;
      AHK_LANES(k, p.nlock) {
        int n1 = p.lock[2 * k], n2 = p.lock[2 * k + 1];
        double d;
        if (n1 != 0) d = n2 != 0 ? fabs(dx[n1 - 1] - dx[n2 - 1]) : fabs(dx[n1 - 1]);
        else d = fabs(dx[n2 != 0 ? n2 - 1 : n - 1]);   // dx[-1] quirk, dc_analysis.py:941
        if (d > dmax) dmax = d;                         // (a NaN step never wins: as `d > lim` before)
      }
      dmax = Grp<G>::fmax_all(dmax, c.mask);
      if (dmax > lim) td = fmin(td, fm::div(lim, dmax));
    }
    // convergence_check (utilities.py:392-549): np.allclose(x, x + dx) per unit class and
    // |residual| <= the other class's absolute tolerance; NaN never passes
    int ok = 1, finite = 1;
    AHK_LANES(i, n) {
      const double d = dx[i];
      const double xn = x[i] + td * d;
      x[i] = xn;
      if (!isfinite(xn)) finite = 0;
      const bool isv = i < p.nv1;
      if (!(fabs(d) <= (isv ? o.vea : o.iea) + (isv ? o.ver : o.ier) * fabs(xn + d))) ok = 0;
      if (!(fabs(res[i]) <= (isv ? o.iea : o.vea))) ok = 0;
    }
    Grp<G>::sync(c.mask);
    if (!p.ndev) { conv = 1; break; }            // linear circuit: one iteration
    if (Grp<G>::all(ok, c.mask)) { conv = 1; break; }
    if (!Grp<G>::all(finite, c.mask)) { it = MAXIT; break; }  // NaN never converges
  }
  *iters = it;
  return conv;
}

// 10 ** -(idx+1), idx = 0..9 (dc_analysis.py:452-455) and the source-stepping factors
// (dc_analysis.py:456-457); literal constants, no local-memory tables
AHK_DEV double gmin_step_value(int i) {
  return i == 0 ? 1e-1 : i == 1 ? 1e-2 : i == 2 ? 1e-3 : i == 3 ? 1e-4 : i == 4 ? 1e-5 :
         i == 5 ? 1e-6 : i == 6 ? 1e-7 : i == 7 ? 1e-8 : i == 8 ? 1e-9 : 1e-10;
}
AHK_DEV double source_step_value(int i) {
  return i == 0 ? 0.001 : i == 1 ? .005 : i == 2 ? .01 : i == 3 ? .03 : i == 4 ? .1 :
         i == 5 ? .3 : i == 6 ? .5 : i == 7 ? .7 : i == 8 ? .8 : .9;
}

// ---- dc_solve: standard -> gmin stepping -> source stepping (dc_analysis.py:128-320) ---
// gmin: value on the node diagonals for the standard / source-stepping attempts.
template <class V>
AHK_DEV int dc_solve(Ctx& c, double gmin, bool with_tran, double time, int MAXIT,
                     int* tot_iters, int* method) {
  const int G = V::G;
  const Prog& p = *c.p;
  const Opts& o = *c.o;
  const int n = p.n;
  double* Nd = c.sm + c.L->Nd;
  const double* Nc = c.sm + c.L->Nc;
  const double* Ntr = c.sm + c.L->Ntr;
  double* T = c.sm + c.L->T;
  double* Bv = c.vb() + c.L->Bv;
  const double* Mv = c.vb() + c.L->Mv;
  const double* Dv = c.vb() + c.L->Dv;
  AHK_LANES(i, n) Nd[i] = Nc[i];
  if (p.ntd) {   // Tt(t), dc_analysis.py:222-237
    Grp<G>::sync(c.mask);
    if (c.sub == 0) {
      AHK_UNROLL
      for (int k = 0; k < p.nsrc; k++) {
        const SrcOp& s = p.src[k];
        if (!s.tf) continue;
        // time None (OP / DC sweep, with_tran false): VSource.V / ISource.I return dc_value
        // (VSource.py:92-95; a missing dc_value is staged as the waveform at t = 0)
        double v = with_tran ? tf_eval(s.tf, c.sm + c.L->Pv + s.ppb + 1, time)
                             : ((s.elem == c.dc_src) ? c.dc_val : c.PP(s.ppb));
        if (s.is_v) Nd[s.row0] += -1 * v;
        else { if (s.row0 >= 0) Nd[s.row0] += v; if (s.row1 >= 0) Nd[s.row1] -= v; }
      }
    }
  }
  int failed = 0;   // bit m: method m failed
  const int use = (o.use_standard_solve_method ? 1 : 0) | (o.use_gmin_stepping ? 2 : 0) |
                  (o.use_source_stepping ? 4 : 0);
  int enabled = -1, gidx = 0, sidx = 0;
  for (int m = 0; m < 3; m++) if ((use >> m) & 1) { enabled = m; break; }
  *tot_iters = 0;
  *method = 0;
  while (true) {
    if (enabled < 0) return 0;
    double g = gmin, f = 1.0;
    if (enabled == 1) g = gmin_step_value(gidx);   // replaces Gmin (dc_analysis.py:262-269,452-455)
    else if (enabled == 2) f = source_step_value(sidx);
    Grp<G>::sync(c.mask);
    c.gadd = g;     // Gmin (and C1 D in a transient) are added while assembling
    AHK_LANES(i, n) T[i] = (enabled == 2 ? f * Nd[i] : Nd[i]) + (with_tran ? Ntr[i] : 0.0);
    Grp<G>::sync(c.mask);
    int it = 0;
    int r = mdn_solver<V>(c, MAXIT, &it);
    if (r >= 0) *tot_iters += it;
    *method = enabled;
    if (r <= 0) {   // set_next_solve_method, dc_analysis.py:354-408
      failed |= 1 << enabled;
      enabled = -1;
      for (int m = 0; m < 3; m++) if (!((failed >> m) & 1) && ((use >> m) & 1)) { enabled = m; break; }
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
template <class V>
AHK_DEV int op_analysis(Ctx& c, bool have_x0, bool use_guess, int* iters) {
  const int G = V::G;
  const Prog& p = *c.p;
  const Opts& o = *c.o;
  const int n = p.n;
  double* x = c.x();
  double* x1 = c.sm + c.L->x1;
  build_N(c);
  if (!have_x0) {
    AHK_LANES(i, n) {
      double s = 0.0;
      if (use_guess && p.nguess > 0) {   // dc_guess.py via the precomputed operator
        AHK_UNROLL
        for (int j = 0; j < p.nguess; j++) {
          double T = p.gconst[j];
          if (p.gpslot[j] >= 0) T += p.gscale[j] * c.PP(p.gpslot[j]);
          s += p.guessG[i * p.nguess + j] * T;
        }
      }
      x[i] = s;
    }
  }
  Grp<G>::sync(c.mask);
  // the two passes share ONE dc_solve call site (not unrolled): the Newton loop, the device
  // models and the unrolled solve are instantiated once per kernel — instruction-cache
  // misses were the largest stall of the C2 kernel when each pass had its own copy
  int n1 = 0, n2 = 0, m1 = 0, m2 = 0, s1 = 0, s2 = 0, status;
  double* r1 = c.sm + c.L->xs;   // residual of pass 1 (kept for PASS2_FAIL)
#pragma unroll 1
  for (int pass = 0; pass < 2; pass++) {
    int it = 0, me = 0;
    const int sres = dc_solve<V>(c, pass == 0 ? o.gmin : 0.0, false, 0.0, o.dc_max_nr_iter, &it, &me);
    if (pass == 0) {
      s1 = sres; n1 = it; m1 = me;
      if (!s1) break;
      const double* res = c.sm + c.L->res;
      AHK_LANES(i, n) { x1[i] = x[i]; r1[i] = res[i]; }
      Grp<G>::sync(c.mask);
    } else {
      s2 = sres; n2 = it; m2 = me;
    }
  }
  if (!s1) {
    *iters = n1;
    status = (m1 << 1);
    const double qnan = nan("");
    AHK_LANES(i, n) x[i] = qnan;
  } else if (!s2) {
    double* resw = c.sm + c.L->res;
    AHK_LANES(i, n) { x[i] = x1[i]; resw[i] = r1[i]; }
    *iters = n1;
    status = AHK_ST_OK | AHK_ST_PASS2_FAIL | (m1 << 1);
  } else {
    *iters = n1 + n2;
    int bad = 0;
    AHK_LANES(i, n) {   // results.py:493-523
      bool isv = i < p.nv1;
      double er = isv ? o.ver : o.ier, ea = isv ? o.vea : o.iea;
      if (fabs(x[i] - x1[i]) > er * fmax(fabs(x1[i]), fabs(x[i])) + ea) bad = 1;
    }
    status = AHK_ST_OK | (m2 << 1);
    if (Grp<G>::any(bad, c.mask)) status |= AHK_ST_GMIN_CHECK;
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

// rmin ** (1 / (order + 1)) of the LTE step factor (transient.py:362)
#ifdef __CUDACC__
AHK_DEV double fm_pow(double b, double e) { return fm::exp(fm::log(b) * e); }
#else
AHK_DEV double fm_pow(double b, double e) { return pow(b, e); }
#endif

AHK_DEV double ipow(double b, int e) { double r = 1.0; for (int i = 0; i < e; i++) r *= b; return r; }

// Differentiation formula of the next step: C1, C0[], the LTE coefficients and the predictor.
// History ring: entry k (k = 0 newest) lives in slot (h0 + k) % buflen.
// No cross-lane traffic: every lane forms the (group-uniform) coefficients itself and writes
// the components of C0 / pred it owns; the caller synchronises the group afterwards.
template <class V>
AHK_DEV void get_df(Ctx& c, int method, int order, bool bootstrap, int nh, int h0, int buflen,
                    double step, bool predict, DfOut& d) {
  const int G = V::G;
  constexpr bool REP = RepOf<V>::value;
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
      AHK_ITEMS(i, n) pred[i] = (x0[i] - xm[i]) / dt * step + x0[i];
      d.pred_lte = -0.5 * step * (HT(0) + step - HT(1));
      d.has_pred = 1;
    }
    AHK_ITEMS(i, n) C0[i] = -1. / step * x0[i];
  } else if (method == AHK_TRAP) {                   // trap.py:98-192
    d.C1 = 2.0 / step;
    d.x_lte = 1.0 / 12 * step * step * step;
    const double* x0 = HX(0);
    AHK_ITEMS(i, n) C0[i] = -1.0 * (2.0 / step * x0[i] + hdx[i]);
    if (predict && nh > 2) {
      // quadratic through the last three points (trap.py:171-186 solves the same
      // interpolation with inv(A)); divided differences about t_n
      const double* xa = HX(1);
      const double* xb = HX(2);
      double t1 = HT(1) - HT(0), t2 = HT(2) - HT(0);
      AHK_ITEMS(i, n) {
        double d1 = (xa[i] - x0[i]) / t1, d2 = (xb[i] - x0[i]) / t2;
        double a2 = (d1 - d2) / (t1 - t2), a1 = d1 - a2 * t1;
        pred[i] = a2 * step * step + a1 * step + x0[i];
      }
      d.pred_lte = -1.0 / 6.0 * step * (HT(0) + step - HT(1)) * (HT(0) + step - HT(2));
      d.has_pred = 1;
    }
  } else {                                           // gear.py:122-208
    // s / alpha / gamma: order + 2 entries each (a specialised build knows the order: the
    // loops unroll and the arrays live in registers)
    double s[8], alpha[8], gamma[8];
    s[0] = 0;
    AHK_UNROLL
    for (int i = 1; i < order + 2; i++) s[i] = step + HT(0) - HT(i - 1);
    AHK_UNROLL
    for (int k = 1; k < order + 2; k++) {
      double a = 1.0;
      AHK_UNROLL
      for (int j = 1; j < order + 2; j++) a *= (j == k) ? 1.0 : fm::div(s[j], s[j] - s[k]);
      alpha[k] = a;
    }
    double g0 = 0;
    AHK_UNROLL
    for (int k = 1; k < order + 1; k++) {
      gamma[k] = alpha[k] * (fm::rcp(s[order + 1]) - fm::rcp(s[k]));
      g0 -= gamma[k];
    }
    gamma[0] = g0;
    double xl = 0.0, fact = 1.0;
    AHK_UNROLL
    for (int k = 2; k <= order + 1; k++) fact *= k;
    AHK_UNROLL
    for (int k = 1; k < order + 1; k++) xl += ipow(s[k], order + 1) * fm::div(-1.0 * gamma[k], g0);
    d.x_lte = (((order + 1) & 1) ? -1.0 : 1.0) * (1.0 / fact) * xl;
    d.C1 = g0;
    AHK_ITEMS(i, n) {
      double a = 0.0;
      AHK_UNROLL
      for (int k = 0; k < order; k++) a = a + gamma[k + 1] * HX(k)[i];
      C0[i] = a;
    }
    if (predict) {
      double pl = -1.0 / fact;
      AHK_UNROLL
      for (int k = 1; k < order + 2; k++) pl *= s[k];
      AHK_ITEMS(i, n) {
        double a = 0.0;
        AHK_UNROLL
        for (int k = 1; k < order + 2; k++) a = a + alpha[k] * HX(k - 1)[i];
        pred[i] = a;
      }
      d.pred_lte = pl;
      d.has_pred = 1;
    }
  }
#undef HT
#undef HX
}

struct TranArgs {
  double tstart, tstep, tstop;
  int method, use_step_control;
  long long max_rows;
  double* t_out; double* x_out;
  int* rows; int* counters;
};

// phases of a group inside the flat transient loop
enum { PH_STEP = 0, PH_SOLVE = 1, PH_NR = 2, PH_DSFAIL = 3, PH_DONE = 4 };

// transient.transient_analysis for one instance, in WARP-UNIFORM ("flat") form.
// in: x() holds x(tstart).  returns the status word.
//
// The reference nests three loops per instance — time steps (transient.py:323-405), the
// homotopy stages of dc_solve (dc_analysis.py:258-320) and the Newton iterations of mdn_solver
// (dc_analysis.py:796-836).  A warp carries 32 / G instances whose trip counts differ at every
// level (Newton iterations per solve, rejected steps, step counts), so nested loops make the
// groups of a warp wait for the slowest at EVERY loop exit.  Here the nest is one loop whose
// body is ONE Newton iteration for every group, preceded / followed by the step and stage
// bookkeeping of the groups that need it in this pass (predicated on the group's phase): no
// group ever waits for another one's iteration count, only the instance with the most Newton
// iterations in total bounds the warp.  `live` = false: padding group of a partial warp (no
// instance): it runs along (the warp's shuffles need all lanes) and never touches HBM.
template <class V>
AHK_DEV int tran_analysis(Ctx& c, const TranArgs& a, bool live = true) {
  const int G = V::G;
  constexpr bool REP = RepOf<V>::value;
  const Prog& p = *c.p;
  const Opts& o = *c.o;
  const int n = p.n;
  const int buflen = c.L->buflen;
#ifdef AHK_JIT   // specialised build: the integration method is a constant of the kernel
  const int method = AHK_JIT_METHOD, use_step_control = AHK_JIT_USC;
#else
  const int method = a.method, use_step_control = a.use_step_control;
#endif
  double* x = c.x();
  double* dx = c.sm + c.L->dx;
  const double* res = c.sm + c.L->res;
  double* xin = c.sm + c.L->xin;     // x at the entry of mdn_solver (restored if the matrix is singular)
  double* xacc = c.sm + c.L->xacc;   // last accepted x
  double* ht = c.sm + c.L->ht;
  double* hx = c.sm + c.L->hx;
  double* hdx = c.sm + c.L->hdx;
  double* C0 = c.sm + c.L->C0;
  double* pred = c.sm + c.L->pred;
  double* Ntr = c.sm + c.L->Ntr;
  double* Nd = c.sm + c.L->Nd;
  const double* Nc = c.sm + c.L->Nc;
  double* T = c.sm + c.L->T;
  const double* Dv = c.vb() + c.L->Dv;
  double* xs = c.sm + c.L->xs;       // NR start point when nothing better exists
  int order, max_x, max_dx, pmax_x;
  if (method == AHK_IMPLICIT_EULER) { order = 1; max_x = 0; max_dx = -1; pmax_x = 1; }
  else if (method == AHK_TRAP) { order = 2; max_x = 0; max_dx = 0; pmax_x = 2; }
  else { order = method - 10; max_x = order; max_dx = -1; pmax_x = order; }
  int first_iters = 0;   // transient.py:272-278
  if (max_x > 0 || max_dx >= 0) { first_iters = max_x; if (max_dx >= 0 && max_dx + 1 > first_iters) first_iters = max_dx + 1; }
  const int start_pred = pmax_x > 0 ? pmax_x : 0;
  const double HMAX = a.tstep;
  const double lim = o.nl_voltages_lock_factor * AHK_VTH;
  const int MAXIT = o.transient_max_nr_iter;
  const int use = (o.use_standard_solve_method ? 1 : 0) | (o.use_gmin_stepping ? 2 : 0) |
                  (o.use_source_stepping ? 4 : 0);
  double tstep = a.tstep;
  build_N(c);
  int h0 = 0, nh = 1;
  AHK_ITEMS(i, n) { hx[i] = x[i]; xs[i] = x[i]; xacc[i] = x[i]; }
  if (c.sub == 0) ht[0] = a.tstart;
  GrpW<G>::sync();
  double time = a.tstart;
  if (use_step_control) tstep = fmin((a.tstop - a.tstart) / 9999.0, HMAX);
  int iter_n = 0, solved = 1, have_x = 0, status = 0;
  long long nrows = 0;
  int n_solves = 0, n_nr = 0, n_rej = 0, n_cut = 0;
  double* t_out = a.t_out + c.inst * a.max_rows;
  double* x_out = a.x_out + c.inst * a.max_rows * n;
  // state of the running dc_solve / mdn_solver
  int failed = 0, enabled = -1, gidx = 0, sidx = 0, ds_iters = 0, it = 0;
  DfOut d;
  d.C1 = 0.0; d.x_lte = 0.0; d.pred_lte = 0.0; d.has_pred = 0;
  int phase = (live && time < a.tstop) ? PH_STEP : PH_DONE;
  while (warp_loop(phase != PH_DONE)) {
    // ===== (1) a new step attempt: DF coefficients, NR start point, Ntran; dc_solve entry =====
    if (warp_any(phase == PH_STEP)) {
      const bool m = phase == PH_STEP;
      if (m) {
        const bool predict = use_step_control && iter_n >= start_pred;
        get_df<V>(c, method, order, iter_n < first_iters, nh, h0, buflen, tstep, predict, d);
      }
      GrpW<G>::sync();
      if (m) {
        // NR start point (transient.py:335-338)
        if (o.transient_prediction_as_x0 && use_step_control && d.has_pred) {
          AHK_ITEMS(i, n) x[i] = pred[i];
        } else if (have_x) {
          AHK_ITEMS(i, n) x[i] = xacc[i];
        } else {
          AHK_ITEMS(i, n) x[i] = xs[i];
        }
        // Ntran = D . C0
        AHK_ITEMS(r, n) {
          double s = 0.0;
          AHK_UNROLL
          for (int k = p.row_ptr[r]; k < p.row_ptr[r + 1]; k++) s += Dv[k] * C0[p.pcol[k]];
          Ntr[r] = s;
        }
        c.C1 = d.C1;
        // dc_solve entry (dc_analysis.py:222-257): N + Tt(t), first enabled method
        if (p.ntd) {
          AHK_ITEMS(i, n) Nd[i] = Nc[i];
        }
      }
      if (p.ntd) {   // Tt(t), dc_analysis.py:222-237
        GrpW<G>::sync();
        if (m && c.sub == 0) {
          AHK_UNROLL
          for (int k = 0; k < p.nsrc; k++) {
            const SrcOp& sr = p.src[k];
            if (!sr.tf) continue;
            double v = tf_eval(sr.tf, c.sm + c.L->Pv + sr.ppb + 1, time + tstep);
            if (sr.is_v) Nd[sr.row0] += -1 * v;
            else { if (sr.row0 >= 0) Nd[sr.row0] += v; if (sr.row1 >= 0) Nd[sr.row1] -= v; }
          }
        }
      }
      if (m) {
        failed = 0; gidx = 0; sidx = 0; ds_iters = 0;
        enabled = -1;
        for (int q = 0; q < 3; q++) if ((use >> q) & 1) { enabled = q; break; }
        phase = enabled < 0 ? PH_DSFAIL : PH_SOLVE;
      }
    }
    // ===== (2) a homotopy stage of dc_solve: Gmin, source factor, T; mdn_solver entry =====
    if (warp_any(phase == PH_SOLVE)) {
      const bool m = phase == PH_SOLVE;
      GrpW<G>::sync();
      if (m) {
        double g = o.gmin, f = 1.0;
        if (enabled == 1) g = gmin_step_value(gidx);   // replaces Gmin (dc_analysis.py:262-269,452-455)
        else if (enabled == 2) f = source_step_value(sidx);
        c.gadd = g;     // Gmin and C1 D are added while assembling
        AHK_ITEMS(i, n) {
          T[i] = (enabled == 2 ? f * Nd[i] : Nd[i]) + Ntr[i];
          xin[i] = x[i];
        }
        it = 0;
        phase = PH_NR;
      }
      GrpW<G>::sync();
    }
    // ===== (3) ONE Newton iteration of every group (dc_analysis.mdn_solver) =====
    // groups that are not iterating (done / padding) compute along on their frozen state and
    // commit nothing
    int r = 2;   // 2 = keep iterating, 1 converged, 0 not converged, -1 singular
    {
      const bool m = phase == PH_NR;
      if (m) it++;
      if (p.ndev) { eval_devices(c, x); GrpW<G>::sync(); }
      const int lin_ok = assemble_solve<V, true>(c);
      // get_td (dc_analysis.py:884-946)
      double td = 1.0;
      if (o.nr_damp_first_iters) td = it < 10 ? 1e-2 : (it < 20 ? 0.1 : 1.0);
      if (o.nl_voltages_lock && p.nlock) {
        double dmax = 0.0;
        AHK_ITEMS(k, p.nlock) {
          int n1 = p.lock[2 * k], n2 = p.lock[2 * k + 1];
          double dd;
          if (n1 != 0) dd = n2 != 0 ? fabs(dx[n1 - 1] - dx[n2 - 1]) : fabs(dx[n1 - 1]);
          else dd = fabs(dx[n2 != 0 ? n2 - 1 : n - 1]);   // dx[-1] quirk, dc_analysis.py:941
          if (dd > dmax) dmax = dd;
        }
        if (!REP) dmax = Grp<G>::fmax_all(dmax, 0xffffffffu);
        if (dmax > lim) td = fmin(td, fm::div(lim, dmax));
      }
      // x += td dx; convergence_check (utilities.py:392-549)
      bool ok = true, finite = true;
      if (m && lin_ok) {
        AHK_ITEMS(i, n) {
          const double di = dx[i];
          const double xn = x[i] + td * di;
          x[i] = xn;
          if (!isfinite(xn)) finite = false;
          const bool isv = i < p.nv1;
          if (!(fabs(di) <= (isv ? o.vea : o.iea) + (isv ? o.ver : o.ier) * fabs(xn + di))) ok = false;
          if (!(fabs(res[i]) <= (isv ? o.iea : o.vea))) ok = false;
        }
      }
      const bool all_ok = REP ? ok : GrpW<G>::all(ok), all_fin = REP ? finite : GrpW<G>::all(finite);
      if (m) {
        if (!lin_ok) {            // LinAlgError: x as it was when mdn_solver was entered
          c.singular = 1;
          AHK_ITEMS(i, n) x[i] = xin[i];
          r = -1;
        } else if (!p.ndev || all_ok) r = 1;
        else if (!all_fin) { it = MAXIT; r = 0; }   // NaN never converges
        else if (it >= MAXIT) r = 0;
      }
      GrpW<G>::sync();
    }
    // ===== (4) mdn_solver returned: next homotopy stage, or dc_solve is finished =====
    int ds = 2;   // dc_solve: 2 = still running, 1 = solved, 0 = failed
    if (phase == PH_DSFAIL) ds = 0;
    if (phase == PH_NR && r != 2) {
      if (r >= 0) ds_iters += it;
      if (r <= 0) {   // set_next_solve_method, dc_analysis.py:354-408
        failed |= 1 << enabled;
        enabled = -1;
        for (int q = 0; q < 3; q++) if (!((failed >> q) & 1) && ((use >> q) & 1)) { enabled = q; break; }
        if (enabled < 0) ds = 0; else phase = PH_SOLVE;
      } else {
        if (enabled == 2 && sidx != 9) { sidx++; phase = PH_SOLVE; }
        else if (enabled == 1 && gidx != 9) { gidx++; phase = PH_SOLVE; }
        else ds = 1;
      }
    }
    // ===== (5) the step attempt is over: LTE step control, accept / reject / cut =====
    if (warp_any(ds != 2)) {
      const bool sd = ds != 2;
      const bool lte_on = sd && ds == 1 && use_step_control && d.has_pred;
      // transient.py:356-365: coeff = min(2, min_i ((aerr + rerr |x_prev_i|) / lte_i)^(1/(order+1)));
      // the power is monotone, so the minimum ratio is found first and raised once.
      double rmin = INFINITY;
      if (lte_on) {
        const double sc = fm::div(d.x_lte, d.pred_lte - d.x_lte);
        AHK_ITEMS(i, n) {
          double lte = fabs(sc * (pred[i] - x[i]));
          if (lte != 0) {
            double re = i < p.nv1 ? o.ver : o.ier;
            double v = fm::div(o.vea + re * fabs(xacc[i]), lte);
            if (v < rmin) rmin = v;
          }
        }
      }
      if (!REP) rmin = Grp<G>::fmin_all(rmin, 0xffffffffu);
      bool stop = false;
      if (sd) {
        n_solves++; n_nr += ds_iters;
        bool accept = false;
        if (ds == 1) {
          const double old_step = tstep;
          bool reject = false;
          if (lte_on) {
            double coeff = 2.0;
            if (rmin < INFINITY) {
              double v = fm_pow(rmin, 1.0 / (order + 1));
              if (v < coeff) coeff = v;
            }
            reject = o.transient_use_aposteriori_step_control &&
                     coeff < o.transient_aposteriori_step_threshold;
            tstep = tstep * coeff;
            if (check_step(o, tstep, time, a.tstop, HMAX)) { status |= AHK_ST_HMIN; solved = 0; stop = true; }
            else if (reject) n_rej++;
          }
          if (!stop && !reject) {
            time = time + old_step;
            have_x = 1;
            iter_n++;
            if (nrows >= a.max_rows) { status |= AHK_ST_ROWS_FULL; solved = 0; stop = true; }
            else accept = true;
          }
          if (accept) {
            // store the row, dxdt = C1*x + C0, push to the ring (transient.py:385-393)
            h0 = (h0 + buflen - 1) % buflen;
            double* hnew = hx + h0 * n;
            AHK_ITEMS(i, n) {
              double xi = x[i];
              xacc[i] = xi;
              hnew[i] = xi;
              hdx[i] = d.C1 * xi + C0[i];
              x_out[nrows * n + i] = xi;
            }
            if (c.sub == 0) { ht[h0] = time; t_out[nrows] = time; }
            nrows++;
            if (nh < buflen) nh++;
            if (o.transient_max_time_iter && iter_n == o.transient_max_time_iter) { solved = 0; stop = true; }
          }
        } else {
          if (use_step_control) {
            tstep = tstep / 5.0;
            if (check_step(o, tstep, time, a.tstop, HMAX)) { status |= AHK_ST_HMIN; solved = 0; stop = true; }
            else n_cut++;
          } else { status |= AHK_ST_NR_FAIL; solved = 0; stop = true; }
          if (!stop && o.transient_max_time_iter && iter_n == o.transient_max_time_iter) { solved = 0; stop = true; }
        }
        phase = (stop || !(time < a.tstop)) ? PH_DONE : PH_STEP;
      }
      GrpW<G>::sync();
    }
  }
  if (solved) status |= AHK_ST_OK;
  if (c.singular) status |= AHK_ST_SINGULAR;
  if (live && c.sub == 0) {
    a.rows[c.inst] = (int)nrows;
    a.counters[4 * c.inst + 0] = n_solves; a.counters[4 * c.inst + 1] = n_nr;
    a.counters[4 * c.inst + 2] = n_rej; a.counters[4 * c.inst + 3] = n_cut;
  }
  return status;
}

}  // namespace ahk
