// jit_gen.hpp — circuit-specialised kernels: the stamp program as compile-time constants.
//
// The AOT kernels (k_dc.cu, k_tran.cu) INTERPRET the stamp program: every Newton
// iteration re-reads row pointers, column indices, gather lists and layout offsets
// from memory, and ncu shows that this interpretation, not the arithmetic, is what the
// small-circuit kernels spend their time on (profiles/r1_ncu_source_lines_c2g.txt: 82 %
// of the warp instructions in core.cuh, 10 % in the device models).  For a batch that is
// worth a few hundred milliseconds of compilation the engine therefore compiles THE SAME
// per-instance source (core.cuh, devices.cuh, entry.cuh) once more with NVRTC for
// sm_100a, with `Prog`, `Opts` and `Layout` replaced by types whose members are
// constants: scalars become `static constexpr`, tables become `__device__ const` arrays
// behind an indexing functor.  After unrolling, every index load, loop bound, layout
// offset, option test and device-type switch folds away; what is left is the
// straight-line arithmetic of this circuit.  Same source => same results as the AOT
// kernels, bit for bit.
//
// This header only GENERATES the constants (plain C++, no CUDA): `jit_spec_source`
// returns the text of "jit_spec.h".  jit.cpp feeds it to NVRTC; tests/hostemu compiles
// it for the host to check the specialised source against the interpreted one.
#pragma once
#include <cmath>
#include <cstdio>
#include <algorithm>
#include <sstream>
#include <string>
#include <vector>

#include "program_build.hpp"

namespace ahk {

struct JitCfg {
  int NC, R, G;            // kernel variant
  int threads, min_blocks; // __launch_bounds__
  int mode;                // MODE_OP / MODE_TRAN / JIT_MODE_AC
  int method, use_step_control;   // transient only
  int ilp = 1;             // devices a lane evaluates interleaved (ekv_eval_v), 1 = one at a time
  std::vector<std::vector<int> > slu;   // pivot sequences of the static-schedule LU (slu_codegen), empty = pivoting LU
};
enum { JIT_MODE_AC = 2 };
// the complex arrays of the AC kernels' layout (ac.cuh: AcLayout) — offsets in doubles
struct JitAc { int RS, Ac, xc, dxc, Nac, pvc, stride; };

namespace jitgen {

inline std::string num(double v) {
  // (NVRTC has neither INFINITY nor __builtin_huge_val; the bit pattern is a constant expression)
  if (std::isnan(v)) return "__builtin_bit_cast(double, 0x7ff8000000000000ULL)";
  if (std::isinf(v)) return v > 0 ? "__builtin_bit_cast(double, 0x7ff0000000000000ULL)"
                                  : "__builtin_bit_cast(double, 0xfff0000000000000ULL)";
  char b[64];
  std::snprintf(b, sizeof b, "%a", v);   // hexadecimal floating literal: exact
  return b;
}
inline std::string num(int v) { return std::to_string(v); }
inline std::string num(unsigned short v) { return std::to_string((int)v); }
inline std::string num(short v) { return std::to_string((int)v); }
inline std::string num(unsigned char v) { return std::to_string((int)v); }
inline std::string num(signed char v) { return std::to_string((int)v); }

template <typename T>
void arr(std::ostringstream& s, const char* type, const char* name, const std::vector<T>& v) {
  s << "AHK_JD_CONST " << type << " " << name << "[" << (v.empty() ? 1 : v.size()) << "] = {";
  if (v.empty()) s << "0";
  for (size_t i = 0; i < v.size(); i++) s << (i ? "," : "") << num(v[i]);
  s << "};\n";
}

}  // namespace jitgen

// ---- static elimination schedule of the one-lane / replicated LU (core.cuh: slu_solve_thread,
// ac.cuh: the complex form) ---------------------------------------------------------------------
// seq[k] = the POSITION LAPACK's dgetf2 / zgetf2 swaps into place k at step k (what the pivot
// search of lu_solve_thread / zlu_solve_thread returns); the rows are tracked through those
// swaps, the elimination is done symbolically on the stamp pattern (every structural entry
// counts as non-zero), and the statements come out on literal row / column indices.
// SEVERAL sequences form a trie: a common prefix is emitted once, and where they part the code
// finds out which of the known continuations dgesv's rule takes for THIS matrix
// (`if (p is the choice) {...} else if (q is the choice) {...} else same = false`) — C1's AC solves
// take one sequence for 84 % of the (instance, frequency) pairs and another, different in the
// last two steps only, for 16 %.  cx: complex entries `m[r][c]` (ac.cuh: cd, magnitude
// |re| + |im| as izamax), else real `a[r][c]`.  Returns "" if no sequence is valid for the pattern.
struct SluEmit {
  const int n; const bool cx;
  const HostProg* lazyH = nullptr;   // complex form: rows are ASSEMBLED inside the body, right before
                                     // the first step that reads them (short live ranges: the whole
                                     // 14 x 15 complex matrix of C1 does not fit in registers)
  std::vector<std::string> out;
  int ids = 0;
  SluEmit(int n_, bool cx_) : n(n_), cx(cx_) {}
  void need_row(unsigned& done, int r, int ind) {
    if (!lazyH || ((done >> r) & 1u)) return;
    done |= 1u << r;
    const HostProg& H = *lazyH;
    line(ind, "AHK_SLU_FENCE { cd acc = Nac[" + std::to_string(r) + "];");
    for (int q = H.row_ptr[r]; q < H.row_ptr[r + 1]; q++) {
      const int c = H.pcol[q];
      line(ind + 1, "{ const cd v = cd{Bv[" + std::to_string(q) + "], om * Dv[" + std::to_string(q) + "]}; " + A(r, c) +
                    " = v; acc = cadd(acc, cmul(v, x[" + std::to_string(c) + "])); }");
    }
    line(ind + 1, A(r, n) + " = cd{-acc.re, -acc.im}; }");
  }
  std::string A(int r, int c) const { return std::string(cx ? "m[" : "a[") + std::to_string(r) + "][" + std::to_string(c) + "]"; }
  std::string mag(const std::string& e) const { return cx ? "(fabs(" + e + ".re) + fabs(" + e + ".im))" : "fabs(" + e + ")"; }
  void line(int ind, const std::string& t) { out.push_back(std::string((size_t)ind * 2, ' ') + t); }

  struct State {
    std::vector<unsigned> pat; std::vector<int> rowat, prow; std::vector<unsigned> ucols; std::vector<std::string> inv;
    unsigned done = 0;     // rows assembled so far (lazy form)
  };
  static bool valid(const std::vector<unsigned>& pat0, const std::vector<int>& seq, int n) {
    if ((int)seq.size() != n) return false;
    std::vector<unsigned> pat = pat0; std::vector<int> rowat(n);
    for (int i = 0; i < n; i++) rowat[i] = i;
    for (int k = 0; k < n; k++) {
      const int p = seq[k];
      if (p < k || p >= n) return false;
      const int pr = rowat[p];
      if (!((pat[pr] >> k) & 1u)) return false;
      const unsigned uc = pat[pr] & ~((2u << k) - 1u);
      for (int q = k; q < n; q++) {
        const int r = rowat[q];
        if (r != pr && ((pat[r] >> k) & 1u)) pat[r] = (pat[r] | uc) & ~(1u << k);
      }
      rowat[p] = rowat[k]; rowat[k] = pr;
    }
    return true;
  }
  // elimination of step k with the pivot at position p (state advanced in place)
  void eliminate(State& S, int k, int p, const std::string& ivname, int ind) {
    const int pr = S.rowat[p];
    std::vector<int> cand;
    for (int q = k; q < n; q++) { const int r = S.rowat[q]; if (r != pr && ((S.pat[r] >> k) & 1u)) cand.push_back(r); }
    S.rowat[p] = S.rowat[k]; S.rowat[k] = pr; S.prow[k] = pr;
    const unsigned uc = S.pat[pr] & ~((2u << k) - 1u);
    S.ucols[k] = uc; S.inv[k] = ivname;
    for (int r : cand) {
      line(ind, std::string("{ const ") + (cx ? "cd l = cmul(" + A(r, k) + ", " + ivname + ");" : "double l = " + A(r, k) + " * " + ivname + ";"));
      for (int j = k + 1; j <= n; j++)
        if (j == n || ((uc >> j) & 1u))
          line(ind + 1, cx ? A(r, j) + " = csub(" + A(r, j) + ", cmul(l, " + A(pr, j) + "));" : A(r, j) + " -= l * " + A(pr, j) + ";");
      line(ind, "}");
      S.pat[r] = (S.pat[r] | uc) & ~(1u << k);
    }
  }
  void back(const State& S, int ind) {
    for (int k = n - 1; k >= 0; k--) {
      const int pr = S.prow[k];
      const std::string xk = "slux" + std::to_string(k);
      line(ind, std::string("{ ") + (cx ? "cd" : "double") + " s = " + A(pr, n) + ";");
      for (int j = k + 1; j < n; j++)
        if ((S.ucols[k] >> j) & 1u)
          line(ind + 1, cx ? "s = csub(s, cmul(" + A(pr, j) + ", slux" + std::to_string(j) + "));" : "s -= " + A(pr, j) + " * slux" + std::to_string(j) + ";");
      line(ind + 1, cx ? xk + " = cmul(s, " + S.inv[k] + "); dx[" + std::to_string(k) + "] = " + xk + "; }"
                       : xk + " = s * " + S.inv[k] + "; dx[" + std::to_string(k) + "] = " + xk + "; }");
    }
  }
  void gen(State S, int k, std::vector<const std::vector<int>*> seqs, int ind) {
    if (k == n) { back(S, ind); return; }
    std::vector<int> kids;                       // distinct positions the sequences take at step k
    for (auto* q : seqs) if (std::find(kids.begin(), kids.end(), (*q)[k]) == kids.end()) kids.push_back((*q)[k]);
    std::sort(kids.begin(), kids.end());
    // magnitudes of the structural candidates of column k, in LAPACK's search order
    std::vector<int> cq;
    for (int q = k; q < n; q++) if ((S.pat[S.rowat[q]] >> k) & 1u) cq.push_back(q);
    for (int q : cq) need_row(S.done, S.rowat[q], ind);
    const int id = ids++;
    auto mg = [&](int q) { return "mg" + std::to_string(id) + "_" + std::to_string(q); };
    for (int q : cq) line(ind, "const double " + mg(q) + " = " + mag(A(S.rowat[q], k)) + ";");
    auto cond = [&](int p) {                     // position p holds the first maximum, and it is non-zero
      std::string c = "(" + mg(p) + " > 0.0)";
      for (int q : cq) if (q != p) c += " && (" + mg(q) + (q < p ? " < " : " <= ") + mg(p) + ")";
      return c;
    };
    const std::string iv = "iv" + std::to_string(id);
    auto rcp = [&](int p) { return cx ? "crecip(" + A(S.rowat[p], k) + ")" : "fm::rcp(" + A(S.rowat[p], k) + ")"; };
    if (kids.size() == 1) {                      // one continuation: verified, not branched on
      const int p = kids[0];
      line(ind, "if (!(" + cond(p) + ")) same = false;");
      line(ind, std::string("const ") + (cx ? "cd " : "double ") + iv + " = " + rcp(p) + ";");
      eliminate(S, k, p, iv, ind);
      gen(S, k + 1, seqs, ind);
      return;
    }
    for (size_t i = 0; i < kids.size(); i++) {
      const int p = kids[i];
      line(ind, std::string(i ? "} else if (" : "if (") + cond(p) + ") {");
      State T = S;
      line(ind + 1, std::string("const ") + (cx ? "cd " : "double ") + iv + " = " + rcp(p) + ";");
      eliminate(T, k, p, iv, ind + 1);
      std::vector<const std::vector<int>*> sub;
      for (auto* q : seqs) if ((*q)[k] == p) sub.push_back(q);
      gen(T, k + 1, sub, ind + 1);
    }
    line(ind, "} else same = false;");
  }
};

inline std::string slu_codegen(const HostProg& H, const std::vector<std::vector<int> >& seqs_in, bool cx = false) {
  const int n = H.hdr.n;
  if (n > 31 || seqs_in.empty()) return "";
  std::vector<unsigned> pat(n, 0u);
  for (int r = 0; r < n; r++)
    for (int q = H.row_ptr[r]; q < H.row_ptr[r + 1]; q++) pat[r] |= 1u << H.pcol[q];
  std::vector<const std::vector<int>*> seqs;
  for (auto& q : seqs_in) {
    if (!SluEmit::valid(pat, q, n)) continue;
    bool dup = false;
    for (auto* o : seqs) if (*o == q) dup = true;
    if (!dup) seqs.push_back(&q);
  }
  if (seqs.empty()) return "";
  SluEmit E(n, cx);
  if (cx) E.lazyH = &H;
  for (int k = 0; k < n; k++) E.line(0, std::string(cx ? "cd" : "double") + " slux" + std::to_string(k) + (cx ? " = cd{0.0, 0.0};" : " = 0.0;"));
  SluEmit::State S;
  S.pat = pat; S.rowat.resize(n); S.prow.assign(n, 0); S.ucols.assign(n, 0u); S.inv.assign(n, "");
  for (int i = 0; i < n; i++) S.rowat[i] = i;
  E.gen(S, 0, seqs, 0);
  std::string body;
  for (auto& t : E.out) body += "  " + t + " \\\n";
  return body;
}
inline std::string slu_codegen(const HostProg& H, const std::vector<int>& seq, bool cx = false) {
  return slu_codegen(H, std::vector<std::vector<int> >{seq}, cx);
}

// Nominal guess of the pivot sequence (before the calibration launch has seen the real one):
// dgetf2's search on the linear stamps at their nominal values, with every Jacobian / dynamic
// entry of the pattern at a small conductance and, for a transient, the dynamic elements at the
// DF coefficient of the nominal step.  Only a starting point: a wrong guess costs one
// more compilation, never a wrong result (slu_solve_thread verifies every pivot).
inline std::vector<int> slu_guess(const HostProg& H, double gmin, double c1 = 0.0) {
  const int n = H.hdr.n, np = H.hdr.npos;
  std::vector<double> A((size_t)n * n, 0.0);
  for (int q = 0; q < np; q++) {
    double v = 0.0;
    for (int t = H.term_ptr[q]; t < H.term_ptr[q + 1]; t++) {
      const int kd = H.term_kind[t], ak = kd < 0 ? -kd : kd;
      const double sg = kd < 0 ? -1.0 : 1.0, P = H.nominal[H.term_slot[t]];
      if (ak == 1) v += sg * (1. / P); else if (ak == 2) v += sg * P; else v += sg;
    }
    if (H.pdiag[q]) v += gmin;
    A[(size_t)H.prowi[q] * n + H.pcol[q]] = v;
  }
  // transient: + C1 D with the DF coefficient of the nominal step (core.cuh: build_values)
  if (c1 != 0.0)
    for (const LinOp& op : H.lin) {
      auto add = [&](int i, double v) { if (op.pos[i] >= 0) A[(size_t)H.prowi[op.pos[i]] * n + H.pcol[op.pos[i]]] += c1 * v; };
      if (op.kind == K_C) { const double v = H.nominal[op.pbase]; add(0, v); add(1, -v); add(2, v); add(3, -v); }
      else if (op.kind == K_L) add(0, -H.nominal[op.pbase]);
      else if (op.kind == K_K) { const double M = std::sqrt(H.nominal[op.aux0] * H.nominal[op.aux1]) * H.nominal[op.pbase]; add(0, -M); add(1, -M); }
    }
  for (int q = 0; q < np; q++) {   // an entry only the devices write: a small conductance
    double& v = A[(size_t)H.prowi[q] * n + H.pcol[q]];
    if (v == 0.0) v = 1e-4;
  }
  std::vector<int> seq(n, 0);
  for (int k = 0; k < n; k++) {
    int p = k; double mx = std::fabs(A[(size_t)k * n + k]);
    for (int i = k + 1; i < n; i++) { const double v = std::fabs(A[(size_t)i * n + k]); if (v > mx) { mx = v; p = i; } }
    seq[k] = p;
    if (mx == 0.0) continue;
    if (p != k) for (int j = 0; j < n; j++) std::swap(A[(size_t)k * n + j], A[(size_t)p * n + j]);
    for (int i = k + 1; i < n; i++) {
      const double l = A[(size_t)i * n + k] / A[(size_t)k * n + k];
      if (l == 0.0) continue;
      for (int j = k + 1; j < n; j++) {
        const double u = A[(size_t)k * n + j];
        if (u != 0.0) { double& t = A[(size_t)i * n + j]; t -= l * u; if (t == 0.0) t = 1e-300; }   // fill stays structural
      }
    }
  }
  return seq;
}

// Text of jit_spec.h for one (circuit, varied slots, options, layout, variant).
inline std::string jit_spec_source(const HostProg& H, const Opts& o, const Layout& L, const JitCfg& cfg,
                                   const JitAc* ac = nullptr) {
  using namespace jitgen;
  std::ostringstream s;
  const Prog& h = H.hdr;
  s << "// generated by jit_gen.hpp — do not edit\n#pragma once\n#define AHK_JIT 1\n";
  s << "#define AHK_JIT_NC " << cfg.NC << "\n#define AHK_JIT_R " << cfg.R << "\n#define AHK_JIT_G " << cfg.G << "\n";
  s << "#define AHK_JIT_THREADS " << cfg.threads << "\n#define AHK_JIT_MINB " << cfg.min_blocks << "\n";
  s << "#define AHK_JIT_MODE " << cfg.mode << "\n#define AHK_JIT_METHOD " << cfg.method
    << "\n#define AHK_JIT_USC " << cfg.use_step_control << "\n";
  {
    bool all_ekv = !H.dev.empty();
    for (const DevOp& q : H.dev) if (q.type != AHK_EKV || !q.active) all_ekv = false;
    s << "#define AHK_JIT_ILP " << (cfg.ilp > 1 ? cfg.ilp : 1) << "\n#define AHK_JIT_ALL_EKV " << (all_ekv ? 1 : 0) << "\n";
  }
  if (!cfg.slu.empty() && cfg.R >= h.n) {   // every lane holds all rows (AC: lane per frequency chunk)
    const std::string body = slu_codegen(H, cfg.slu, cfg.mode == JIT_MODE_AC);
    if (!body.empty()) s << "#define AHK_JIT_SLU 1\n#define AHK_JIT_SLU_BODY \\\n" << body << "\n";
  }
  s << "#include \"program.h\"\n";
  s << "#ifdef __CUDACC__\n#define AHK_JD_CONST __constant__ const\n#define AHK_JFN __device__ __forceinline__\n"
       "#else\n#define AHK_JD_CONST static const\n#define AHK_JFN inline\n#endif\n";
  s << "namespace ahk {\nnamespace jd {\n";
  arr(s, "int", "pstage", H.pstage); arr(s, "int", "gpslot", H.gpslot);
  arr(s, "int", "row_ptr", H.row_ptr);
  arr(s, "unsigned short", "pcol", H.pcol); arr(s, "unsigned short", "pcd", H.pcd);
  arr(s, "unsigned short", "prowi", H.prowi); arr(s, "unsigned char", "pdiag", H.pdiag);
  arr(s, "int", "nl_ptr", H.nl_ptr); arr(s, "short", "nl_src", H.nl_src); arr(s, "signed char", "nl_sgn", H.nl_sgn);
  arr(s, "int", "tx_ptr", H.tx_ptr); arr(s, "short", "tx_src", H.tx_src); arr(s, "signed char", "tx_sgn", H.tx_sgn);
  arr(s, "int", "term_ptr", H.term_ptr); arr(s, "int", "term_slot", H.term_slot);
  arr(s, "signed char", "term_kind", H.term_kind);
  arr(s, "int", "lock", H.lock);
  arr(s, "double", "guessG", H.guessG); arr(s, "int", "gslot", H.gslot);
  arr(s, "double", "gscale", H.gscale); arr(s, "double", "gconst", H.gconst);
  {
    // the nominal value of a VARIED slot is never read (the batch overrides it): zero it, so
    // that the text — and with it the cache key of the compiled kernel — does not depend on
    // the values of a particular batch
    std::vector<double> nom = H.nominal;
    for (size_t i = 0; i < nom.size() && i < H.slot_vary.size(); i++) if (H.slot_vary[i] >= 0) nom[i] = 0.0;
    arr(s, "double", "nominal", nom);
  }
  arr(s, "int", "slot_vary", H.slot_vary);
  s << "AHK_JD_CONST LinOp lin[" << (H.lin.empty() ? 1 : H.lin.size()) << "] = {";
  if (H.lin.empty()) s << "{0,0,0,0,{-1,-1,-1,-1,-1,-1}}";
  for (size_t i = 0; i < H.lin.size(); i++) {
    const LinOp& q = H.lin[i];
    s << (i ? "," : "") << "{" << q.kind << "," << q.pbase << "," << q.aux0 << "," << q.aux1 << ",{";
    for (int k = 0; k < 6; k++) s << (k ? "," : "") << q.pos[k];
    s << "}}";
  }
  s << "};\n";
  s << "AHK_JD_CONST SrcOp src[" << (H.src.empty() ? 1 : H.src.size()) << "] = {";
  if (H.src.empty()) s << "{0,0,0,0,0,0,0,0,0}";
  for (size_t i = 0; i < H.src.size(); i++) {
    const SrcOp& q = H.src[i];
    s << (i ? "," : "") << "{" << q.is_v << "," << q.elem << "," << q.pbase << "," << q.tf << "," << q.has_ac
      << "," << q.ppb << "," << q.ntf << "," << q.row0 << "," << q.row1 << "}";
  }
  s << "};\n";
  s << "AHK_JD_CONST DevOp dev[" << (H.dev.empty() ? 1 : H.dev.size()) << "] = {";
  if (H.dev.empty()) s << "{0,0,0,0,{-1,-1,-1,-1},0,-1,-1,-1,-1,0}";
  for (size_t i = 0; i < H.dev.size(); i++) {
    const DevOp& q = H.dev[i];
    s << (i ? "," : "") << "{" << q.type << "," << q.pbase << "," << q.mbase << "," << q.flags << ",{"
      << q.node[0] << "," << q.node[1] << "," << q.node[2] << "," << q.node[3] << "}," << q.dout << ","
      << q.con << "," << q.ppb << "," << q.pmb << "," << q.st << "," << q.active << "}";
  }
  s << "};\n}  // namespace jd\n";
  s << "#define AHK_JARR(T, name) struct name##_t { AHK_JFN const T& operator[](int i) const "
       "{ return jd::name[i]; } } name\n";
  s << "struct Prog {\n";
#define SC(f) s << "  static constexpr int " #f " = " << h.f << ";\n"
  SC(n); SC(nv1); SC(npos); SC(nlin); SC(nsrc); SC(ntd); SC(ndev); SC(ndout); SC(nlock); SC(nguess);
  SC(nparams); SC(nac); SC(npersist);
#undef SC
  s << "  AHK_JARR(int, pstage); AHK_JARR(int, gpslot); AHK_JARR(int, row_ptr);\n"
       "  AHK_JARR(unsigned short, pcol); AHK_JARR(unsigned short, pcd); AHK_JARR(unsigned short, prowi);\n"
       "  AHK_JARR(unsigned char, pdiag); AHK_JARR(int, nl_ptr); AHK_JARR(short, nl_src);\n"
       "  AHK_JARR(signed char, nl_sgn); AHK_JARR(int, tx_ptr); AHK_JARR(short, tx_src);\n"
       "  AHK_JARR(signed char, tx_sgn); AHK_JARR(int, term_ptr); AHK_JARR(int, term_slot);\n"
       "  AHK_JARR(signed char, term_kind); AHK_JARR(LinOp, lin); AHK_JARR(SrcOp, src); AHK_JARR(DevOp, dev);\n"
       "  AHK_JARR(int, lock); AHK_JARR(double, guessG); AHK_JARR(int, gslot); AHK_JARR(double, gscale);\n"
       "  AHK_JARR(double, gconst); AHK_JARR(double, nominal); AHK_JARR(int, slot_vary);\n};\n";
  s << "struct Opts {\n";
#define SD(f) s << "  static constexpr double " #f " = " << num(o.f) << ";\n"
#define SI(f) s << "  static constexpr int " #f " = " << o.f << ";\n"
  SD(vea); SD(ver); SD(iea); SD(ier); SD(gmin); SD(cmin); SD(hmin); SD(nl_voltages_lock_factor);
  SD(transient_aposteriori_step_threshold);
  SI(nl_voltages_lock); SI(nr_damp_first_iters); SI(dc_max_nr_iter); SI(transient_max_nr_iter);
  SI(ac_max_nr_iter); SI(use_standard_solve_method); SI(use_gmin_stepping); SI(use_source_stepping);
  SI(transient_prediction_as_x0); SI(transient_use_aposteriori_step_control); SI(transient_max_time_iter);
#undef SD
#undef SI
  s << "};\nstruct Layout {\n";
#define SL(f) s << "  static constexpr int " #f " = " << L.f << ";\n"
  SL(RS); SL(ARS); SL(A); SL(x); SL(dx); SL(T); SL(res); SL(Bv); SL(Mv); SL(Dv); SL(Nd); SL(Nc); SL(Ntr);
  SL(dout); SL(dst); SL(x1); SL(xs); SL(xacc); SL(hx); SL(hdx); SL(ht); SL(C0); SL(pred); SL(bytes); SL(Pv);
  SL(dcon); SL(pvb); SL(xin); SL(big); SL(brhs); SL(nzm); SL(czm); SL(lpan); SL(vglob); SL(vlen); SL(buflen); SL(stride);
#undef SL
  s << "};\n";
  if (ac) {
    s << "struct AcLayout {\n  Layout L;\n";
#define SA(f) s << "  static constexpr int " #f " = " << ac->f << ";\n"
    SA(RS); SA(Ac); SA(xc); SA(dxc); SA(Nac); SA(pvc); SA(stride);
#undef SA
    s << "};\n";
  }
  s << "}  // namespace ahk\n";
  return s.str();
}

}  // namespace ahk
