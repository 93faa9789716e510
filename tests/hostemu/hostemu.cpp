// hostemu.cpp — TEST-ONLY.  Compiles the CUDA engine's per-instance source
// (ahkab_b200/csrc/{core,entry,devices}.cuh, program_build.hpp) for the HOST
// with one lane per instance (G = 1), so that the kernel logic can be checked
// against the oracle in a container that has no GPU.  It is built into
// tests/hostemu/_build/ and loaded only by tests/test_hostemu.py; the product
// package never loads it (there is no CPU fallback in ahkab_b200).
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>
using std::isnan; using std::isinf; using std::isfinite;
#include "../../ahkab_b200/csrc/entry.cuh"
#include "../../ahkab_b200/csrc/program_build.hpp"
#include "../../ahkab_b200/csrc/ac.cuh"
#include "../../ahkab_b200/csrc/jit_gen.hpp"
#include "../../ahkab_b200/csrc/biglin3.cuh"
#include "../../ahkab_b200/csrc/biglin3_gen.hpp"

using namespace ahk;

static void copy_opts(const ahk_options* o, Opts& q) { std::memcpy(&q, o, sizeof(Opts)); }

// Host-testable kernel variants: one lane per instance (no cross-lane traffic).
//   0: Var<0,0,1> generic shared-memory fallback   1: Var<4,3,1>   2: Var<6,5,1>
static int g_variant = 0;
static int variant_nc() { return g_variant == 1 ? 4 : (g_variant == 2 ? 6 : 0); }

struct Emu {
  HostProg H; Prog p; Opts o; Layout L; std::vector<double> sm;
  Emu(const ahk_circuit* c, const ahk_options* opts, int n_vary, const int32_t* slots, int mode,
      int method, int usc) {
    build_program(c, H);
    set_vary_slots(H, n_vary, slots);
    p = H.view();
    copy_opts(opts, o);
    // n > AHK_SMALL_NMAX: the large-n layout of the generic variant (matrix in a "global" workspace)
    const bool big = variant_nc() == 0 && H.hdr.n > AHK_SMALL_NMAX;
    L = make_layout(H, mode, method, usc, 1, variant_nc(), variant_nc() ? variant_nc() - 1 : 1, false, big);
    sm.assign(L.stride + 64, 0.0);
    if (big) Ag.assign((size_t)H.hdr.n * L.RS + L.vlen + (size_t)H.hdr.n * 8, 0.0);
  }
  std::vector<double> Ag;
  Ctx ctx(const double* vary, long long B, long long b) {
    Ctx c; c.p = &p; c.o = &o; c.L = &L; c.sm = sm.data(); c.vary = vary; c.B = B; c.inst = b;
    c.sub = 0; c.G = 1; c.mask = 1u; c.Ag = Ag.empty() ? nullptr : Ag.data();
    c.Vg = (c.Ag && L.vglob) ? c.Ag + (size_t)p.n * L.RS : nullptr;
    return c;
  }
  bool fits() const {
    int nc = variant_nc();
    return nc == 0 ? p.n <= AHK_BIG_NMAX : p.n <= nc - 1;
  }
};

#define DISPATCH(CALL)                                   \
  switch (g_variant) {                                   \
    case 1: { typedef Var<4, 3, 1> V; CALL; } break;     \
    case 2: { typedef Var<6, 5, 1> V; CALL; } break;     \
    default: { typedef Var<0, 0, 1> V; CALL; } break;    \
  }

extern "C" {

int emu_set_variant(int v) { if (v < 0 || v > 2) return -1; g_variant = v; return 0; }

int emu_op(const ahk_circuit* c, const ahk_options* o, int64_t B, int n_vary, const int32_t* slots,
           const double* vary, const double* x0, int use_guess, double* x, double* resid,
           int32_t* iters, int32_t* status) {
  Emu e(c, o, n_vary, slots, MODE_OP, 0, 0);
  if (!e.fits()) return -2;
  DcArgs a = {-1, nullptr, 1, x0, use_guess, x, resid, iters, status};
  for (int64_t b = 0; b < B; b++) { Ctx cx = e.ctx(vary, B, b); DISPATCH(run_dc<V>(cx, a)) }
  return 0;
}

int emu_dc_sweep(const ahk_circuit* c, const ahk_options* o, int64_t B, int n_vary,
                 const int32_t* slots, const double* vary, int src_elem, const double* sweep,
                 int npts, const double* x0, int use_guess, double* x_out, int32_t* iters,
                 int32_t* status) {
  Emu e(c, o, n_vary, slots, MODE_OP, 0, 0);
  if (!e.fits()) return -2;
  DcArgs a = {src_elem, sweep, npts, x0, use_guess, x_out, nullptr, iters, status};
  for (int64_t b = 0; b < B; b++) { Ctx cx = e.ctx(vary, B, b); DISPATCH(run_dc<V>(cx, a)) }
  return 0;
}

int emu_tran(const ahk_circuit* c, const ahk_options* o, int64_t B, int n_vary, const int32_t* slots,
             const double* vary, const double* x0, double tstart, double tstep, double tstop,
             int method, int usc, int64_t max_rows, double* t_out, double* x_out, int32_t* rows,
             int32_t* counters, int32_t* status) {
  Emu e(c, o, n_vary, slots, MODE_TRAN, method, usc);
  if (!e.fits()) return -2;
  TranArgs a = {tstart, tstep, tstop, method, usc, max_rows, t_out, x_out, rows, counters};
  for (int64_t b = 0; b < B; b++) { Ctx cx = e.ctx(vary, B, b); DISPATCH(run_tran<V>(cx, x0, a, status)) }
  return 0;
}

int emu_ac(const ahk_circuit* c, const ahk_options* o, int64_t B, int n_vary, const int32_t* slots,
           const double* vary, const double* x_op, const double* omegas, int npts, double* x_out,
           int32_t* status) {
  Emu e(c, o, n_vary, slots, MODE_OP, 0, 0);
  const int nc = g_variant == 1 ? 4 : 0;   // AC register variant on the host: Var<4,3,1>
  if (nc && e.p.n > nc - 1) return -2;
  const bool big = nc == 0 && e.p.n > AHK_SMALL_NMAX;
  AcLayout AL = make_ac_layout(e.H, 1, nc, big);
  std::vector<double> sm(AL.stride + 64, 0.0);
  std::vector<double> Ag(big ? (size_t)2 * e.p.n * AL.RS + AL.L.vlen : 0, 0.0);
  AcArgs a = {x_op, omegas, npts, 0, npts, x_out, status};
  for (int64_t b = 0; b < B; b++) {
    status[b] = AHK_ST_OK;
    Ctx cx = e.ctx(vary, B, b);
    cx.sm = sm.data();
    cx.Ag = big ? Ag.data() : nullptr;
    cx.Vg = (big && AL.L.vglob) ? Ag.data() + (size_t)2 * e.p.n * AL.RS : nullptr;
    if (nc) run_ac<Var<4, 3, 1> >(cx, AL, a); else run_ac<Var<0, 0, 1> >(cx, AL, a);
  }
  return 0;
}

// Large linear circuits: the static-schedule path (biglin3.cuh) run on the host, one loop
// iteration per device thread.  Returns the number of instances the schedule could not serve
// (pivot sequence differs from the nominal one: the device hands those to k_biglin), -1 if no
// schedule could be built; flags[B] receives BIG3_OK / BIG3_SINGULAR / BIG3_DYNAMIC.
int emu_big3_sweep(const ahk_circuit* c, const ahk_options* o, int64_t B, int n_vary, const int32_t* slots,
                   const double* vary, int src_elem, const double* sweep, int npts, double* x_out,
                   double* resid, int32_t* iters, int32_t* status, int32_t* flags) {
  HostProg H;
  build_program(c, H);
  set_vary_slots(H, n_vary, slots);
  Opts q; copy_opts(o, q);
  Big3Sched S;
  if (!big3_build(H, q, S)) return -1;
  const int n = S.n;
  std::vector<double> V((size_t)2 * S.nslot * B), Mv((size_t)S.npos * B), bs((size_t)n * B * npts), xs((size_t)n * B * npts);
  std::vector<int> flag(B), ovf(B + 1, 0);
  std::vector<double> G((size_t)S.rslot.size() * B);
  Big3K k;
  k.p = H.view(); k.o = q; k.vary = vary; k.B = B; k.src_elem = src_elem; k.sweep = sweep; k.npts = npts;
  k.pt0 = 0; k.ptc = npts; k.idx32 = 0; k.b_smem = 0;
  k.x_out = x_out; k.iters = iters; k.status = status; k.resid = resid;
  k.n = n; k.npos = S.npos; k.nslot = S.nslot;
  k.step = S.step.data(); k.prow = S.prow.data(); k.cand_slot = S.cand_slot.data(); k.cand_row = S.cand_row.data();
  k.l_slot = S.l_slot.data(); k.l_row = S.l_row.data(); k.u_slot = S.u_slot.data(); k.u_col = S.u_col.data();
  k.upd = S.upd.data();
  k.rslot = S.rslot.data(); k.rindex = S.rindex.data(); k.nrec = (int)S.rslot.size(); k.G = G.data();
  k.V = V.data(); k.Mv = Mv.data(); k.flag = flag.data(); k.bs = bs.data(); k.xs = xs.data(); k.ovf = ovf.data();
  int dyn = 0;
  for (int64_t b = 0; b < B; b++) { flag[b] = big3_factor_one<long long>(k, b); if (flag[b] == BIG3_DYNAMIC) dyn++; flags[b] = flag[b]; }
  std::vector<double> bpriv(n);
  for (int64_t t = 0; t < B * npts; t++) big3_solve_one<long long>(k, t, bpriv.data(), 1LL);   // private right-hand side, stride 1 (the device uses shared memory)
  return dyn;
}

// Text of the COMPILED static schedule (biglin3_gen.hpp) for this circuit / batch: returns its
// length (truncated to cap - 1 bytes), -1 if no schedule / too large; *nf = factor values per
// matrix, *npos = matrix positions (sizes of the F / Mv buffers the caller must provide).
int emu_big3_gen_source(const ahk_circuit* c, const ahk_options* o, int n_vary, const int32_t* slots,
                        char* buf, int64_t cap, int32_t* nf, int32_t* npos) {
  try {
    HostProg H;
    build_program(c, H);
    set_vary_slots(H, n_vary, slots);
    Opts q; copy_opts(o, q);
    Big3Sched S;
    if (!big3_build(H, q, S)) return -1;
    Big3Gen G;
    if (!big3_codegen(H, q, S, G)) return -1;
    if (nf) *nf = G.nf;
    if (npos) *npos = S.npos;
    if (buf && cap > 0) {
      size_t m = std::min<size_t>(G.src.size(), (size_t)cap - 1);
      std::memcpy(buf, G.src.data(), m); buf[m] = 0;
    }
    return (int)G.src.size();
  } catch (...) { return -1; }
}

// Text of the generated jit_spec.h (jit_gen.hpp) for this circuit / batch / analysis / variant:
// returns its length (the text is truncated to cap - 1 bytes), < 0 on error.
int emu_jit_source(const ahk_circuit* c, const ahk_options* o, int n_vary, const int32_t* slots,
                   int mode, int method, int usc, int NC, int R, int G, int threads, int min_blocks,
                   char* buf, int64_t cap, int ilp) {
  try {
    HostProg H;
    build_program(c, H);
    set_vary_slots(H, n_vary, slots);
    Opts q; copy_opts(o, q);
    Layout L = make_layout(H, mode, method, usc, G, NC, R, true);
    JitCfg cfg = {NC, R, G, threads, min_blocks, mode, method, usc, ilp > 1 ? ilp : 1};
    if (const char* f = getenv("AHK_SLU_SEQ")) {   // static-schedule LU along these pivot sequences, ';' between them (tests)
      std::string cur;
      for (const char* z = f;; z++) {
        if (*z == ';' || *z == 0) {
          std::istringstream is(cur); int v; std::vector<int> sq;
          while (is >> v) sq.push_back(v);
          if (!sq.empty()) cfg.slu.push_back(sq);
          cur.clear();
          if (*z == 0) break;
        } else cur += *z;
      }
    }
    std::string t = jit_spec_source(H, q, L, cfg);
    if (buf && cap > 0) {
      size_t m = std::min<size_t>(t.size(), (size_t)cap - 1);
      std::memcpy(buf, t.data(), m); buf[m] = 0;
    }
    return (int)t.size();
  } catch (...) { return -1; }
}

// ---- single-device entry points (same device code the kernels inline) -------------------
void emu_ekv_eval(const double* model8, int np, const double* dev4, double vd, double vg,
                  double vs, double gmin, double* out4) {
  EkvP m;
  m.VTO = model8[0]; m.KP = model8[1]; m.GAMMA = model8[2]; m.PHI = model8[3]; m.COX = model8[4];
  m.LAMBDA = model8[5]; m.XJ = model8[6]; m.UCRIT = model8[7];
  m.W = dev4[0]; m.L = dev4[1]; m.M = dev4[2]; m.N = dev4[3]; m.NP = np;
  double k[EK_NCONST];
  ekv_setup(m, k);
  ekv_eval(k, np, vd, vg, vs, gmin, out4);
}
double emu_ekv_q(double v) { return ekv_q(v); }
void emu_mosq_eval(const double* model7, int np, const double* dev6, double vds, double vgs,
                   double vbs, double iea, double gmin, double* out9) {
  MosqP m;
  m.VTO = model7[0]; m.KP = model7[1]; m.GAMMA = model7[2]; m.PHI = model7[3]; m.LAMBDA = model7[4];
  m.AVT = model7[5]; m.AKP = model7[6];
  m.W = dev6[0]; m.L = dev6[1]; m.M = dev6[2]; m.N = dev6[3]; m.ZVT = dev6[4]; m.ZKP = dev6[5];
  m.NP = np;
  mosq_eval(m, vds, vgs, vbs, iea, gmin, out9);
}
void emu_diode_eval(const double* model9, double area, double v, double vea, double gmin,
                    double* last_vd, double* out2) {
  DiodeP m;
  m.IS = model9[0]; m.N = model9[1]; m.NBV = model9[2]; m.ISR = model9[3]; m.NR = model9[4];
  m.RS = model9[5]; m.BV = model9[6]; m.IKF = model9[7]; m.VT = model9[8]; m.AREA = area;
  double k[DK_NCONST];
  diode_setup(m, k);
  diode_eval(k, v, vea, gmin, last_vd, out2);
}
void emu_switch_eval(const double* model4, double vout, double vin, double* is_on, double gmin,
                     double* out3) {
  SwP m; m.VT = model4[0]; m.VH = model4[1]; m.RON = model4[2]; m.ROFF = model4[3];
  switch_eval(m, vout, vin, gmin, is_on, out3);
}

}  // extern "C"
