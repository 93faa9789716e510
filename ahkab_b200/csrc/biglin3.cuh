// biglin3.cuh — large linear circuits (n > 64), third generation: a STATIC elimination
// schedule, one THREAD per instance, structure-of-arrays values, uniform control flow.
//
// Same mathematics as biglin.cuh / biglin2.cuh (read their headers): per instance the two
// matrices mna + Gmin and mna are factored once with dgesv's pivot rule, every sweep point
// then costs two pairs of triangular solves.  biglin2 discovers the sparsity and the pivots
// of every instance dynamically (slot lists in shared memory, one warp per instance, 2-3
// busy lanes per pivot step, 6.2 ms on C5).  But all instances of a batch share the matrix
// STRUCTURE, and — component tolerances being a few per cent — almost always the PIVOT
// SEQUENCE of the nominal instance too.  So:
//
//  * host, once per (circuit, varied slots, options): factor the nominal matrices densely
//    to get the pivot sequence, then eliminate SYMBOLICALLY along it (every structural
//    entry counts as non-zero, whatever its value) -> a flat schedule of
//        l = val[lslot] / piv ... val[t] -= l * val[u]
//    plus the forward / backward substitution lists (big3_build);
//  * device: thread-per-instance kernels execute that schedule over values stored
//    [slot][instance] — every load and store is coalesced, every thread of a warp does
//    the same thing (no divergence, no shuffles, no shared memory);
//  * every pivot step VERIFIES that the scheduled pivot is the one dgesv's rule would pick
//    for this instance (largest magnitude in the column among the rows not yet used, ties to
//    the lowest row).  The same operations in the same order on the same values give the
//    same bits as biglin2 (extra operations on exact zeros change nothing: a - 0*b == a).
//    An instance whose pivots differ is appended to the overflow list and solved by the
//    dynamic kernel (k_biglin), like a slot overflow of biglin2.
//
// The solve kernel runs one thread per (instance, sweep point): the points of a linear
// circuit are independent (biglin2.cuh).  Scratch (right-hand side, iterate) is [n][thread].
#pragma once
#include <cmath>
#include <map>
#include <set>
#include <vector>

#include "biglin2.cuh"

namespace ahk {

struct Big3Step { int piv, cand0, ncand, l0, nl, u0, nu, upd0; };   // per (m, kk)

struct Big3Sched {          // host copy
  int n = 0, npos = 0, nslot = 0;          // value slots per matrix: npos initial + fills
  std::vector<Big3Step> step;              // [2][n]
  std::vector<int> prow;                   // [2][n] pivot row of step kk
  std::vector<int> cand_slot, cand_row;    // other candidates of the pivot search
  std::vector<int> l_slot, l_row;          // multipliers (stored in place)
  std::vector<int> u_slot, u_col;          // pivot-row entries right of the pivot, ascending column
  std::vector<int> upd;                    // [nl][nu] target slots
  std::vector<int> rslot;                  // parameter slots that enter a stamp as 1 / value (resistors)
  std::vector<int> rindex;                 // [nparams] index into rslot, or -1
  bool ok = false;
};

// Pivot sequence from the nominal values (dense, partial pivoting: largest magnitude among
// the rows not yet used, ties to the lowest row), symbolic elimination along it.
inline bool big3_build(const HostProg& H, const Opts& o, Big3Sched& S) {
  const Prog& h = H.hdr;
  const int n = h.n, np = h.npos;
  S = Big3Sched();
  if (n > 4096) return false;
  S.n = n; S.npos = np;
  std::vector<double> Mv(np, 0.0);
  for (int q = 0; q < np; q++) {
    double v = 0.0;
    for (int t = H.term_ptr[q]; t < H.term_ptr[q + 1]; t++) {
      const int kd = H.term_kind[t], ak = kd < 0 ? -kd : kd;
      const double sg = kd < 0 ? -1.0 : 1.0, P = H.nominal[H.term_slot[t]];
      if (ak == 1) v += sg * (1. / P); else if (ak == 2) v += sg * P; else v += sg;
    }
    Mv[q] = v;
  }
  S.rindex.assign(h.nparams > 0 ? h.nparams : 1, -1);
  for (int t = 0; t < H.term_ptr[np]; t++) {
    const int kd = H.term_kind[t], ak = kd < 0 ? -kd : kd, sl = H.term_slot[t];
    if (ak == 1 && S.rindex[sl] < 0) { S.rindex[sl] = (int)S.rslot.size(); S.rslot.push_back(sl); }
  }
  if (S.rslot.empty()) S.rslot.push_back(0);
  S.step.assign(2 * n, Big3Step());
  S.prow.assign(2 * n, 0);
  int nslot_max = np;
  for (int m = 0; m < 2; m++) {
    // numeric: dense copy
    std::vector<double> A((size_t)n * n, 0.0);
    for (int q = 0; q < np; q++)
      A[(size_t)H.prowi[q] * n + H.pcol[q]] = Mv[q] + ((m == 0 && H.pdiag[q]) ? o.gmin : 0.0);
    // symbolic: per row (col -> slot), per column the rows holding an entry
    std::vector<std::map<int, int> > row(n);
    std::vector<std::set<int> > col(n);
    for (int q = 0; q < np; q++) { row[H.prowi[q]][H.pcol[q]] = q; col[H.pcol[q]].insert(H.prowi[q]); }
    int nslot = np;
    std::vector<char> used(n, 0);
    for (int kk = 0; kk < n; kk++) {
      // candidates = structural entries of column kk in unused rows (ascending row)
      int bi = -1; double best = -1.0;
      for (int r : col[kk]) {
        if (used[r]) continue;
        const double v = std::fabs(A[(size_t)r * n + kk]);
        if (bi < 0 || v > best) { best = v; bi = r; }
      }
      if (bi < 0 || !(best > 0.0) || !std::isfinite(best)) return false;   // nominal instance singular
      used[bi] = 1;
      S.prow[m * n + kk] = bi;
      Big3Step& st = S.step[m * n + kk];
      st.piv = row[bi].at(kk);
      st.cand0 = (int)S.cand_slot.size(); st.l0 = (int)S.l_slot.size();
      st.u0 = (int)S.u_slot.size(); st.upd0 = (int)S.upd.size();
      std::vector<int> lrows;
      for (int r : col[kk]) {
        if (used[r]) continue;                       // (bi is used by now)
        S.cand_slot.push_back(row[r].at(kk)); S.cand_row.push_back(r);
        S.l_slot.push_back(row[r].at(kk)); S.l_row.push_back(r);
        lrows.push_back(r);
      }
      st.ncand = st.nl = (int)lrows.size();
      std::vector<int> ucols;
      for (auto& cs : row[bi]) if (cs.first > kk) { ucols.push_back(cs.first); S.u_slot.push_back(cs.second); S.u_col.push_back(cs.first); }
      st.nu = (int)ucols.size();
      const double inv = 1.0 / A[(size_t)bi * n + kk];
      for (int r : lrows) {
        const double l = A[(size_t)r * n + kk] * inv;
        for (int c : ucols) {
          auto it = row[r].find(c);
          if (it == row[r].end()) { it = row[r].insert(std::make_pair(c, nslot++)).first; col[c].insert(r); }
          S.upd.push_back(it->second);
          A[(size_t)r * n + c] -= l * A[(size_t)bi * n + c];
        }
        A[(size_t)r * n + kk] = 0.0;
        row[r].erase(kk);                            // column kk is done for row r
      }
    }
    nslot_max = std::max(nslot_max, nslot);
  }
  S.nslot = nslot_max;
  S.ok = true;
  return true;
}

// device view of the schedule + buffers
struct Big3K {
  Prog p; Opts o;
  const double* vary; long long B;
  int src_elem; const double* sweep; int npts;
  int idx32;        // every [slot][instance] / [row][thread] array has < 2^31 elements
  int b_smem;       // solve kernel: right-hand sides in shared memory (n * blockDim doubles)
  int pt0, ptc;     // this launch solves the points pt0 .. pt0 + ptc - 1 (scratch holds ptc points)
  double* x_out; int* iters; int* status; double* resid;
  int n, npos, nslot;
  const Big3Step* step; const int* prow;
  const int* cand_slot; const int* cand_row; const int* l_slot; const int* l_row;
  const int* u_slot; const int* u_col; const int* upd;
  const int* rslot; const int* rindex; int nrec;   // reciprocal table (see big3_factor_one)
  double* G;        // [nrec][B]  1 / value of every resistor-like parameter, computed once
  double* V;        // [2][nslot][B]  factors (in place)
  double* Mv;       // [npos][B]      matrix values (for the residual mat-vec)
  int* flag;        // [B] 0 ok, 1 singular, 2 handed to the dynamic kernel
  double* bs;       // [n][B*ptc] right-hand sides
  double* xs;       // [n][B*ptc] iterates
  int* ovf;         // ovf[0] = count, ovf[1..] = instances for k_biglin
};

enum { BIG3_OK = 0, BIG3_SINGULAR = 1, BIG3_DYNAMIC = 2 };

// ---- per-instance factorisation (one thread = one instance) --------------------------------
// IT: index type of the [slot][instance] arrays — 32-bit when they are small enough (one
// IMAD.WIDE per address instead of a 64-bit multiply), 64-bit otherwise
template <typename IT>
AHK_DEV int big3_factor_one(const Big3K& k, long long inst) {
  const Prog& p = k.p;
  const long long B = k.B;
  const int n = k.n;
  auto P = [&](int slot) -> double {
    const int v = p.slot_vary[slot];
    return v < 0 ? p.nominal[slot] : k.vary[(IT)v * (IT)B + inst];
  };
  double* V0 = k.V + inst;
  double* V1 = k.V + (IT)k.nslot * (IT)B + inst;
  // a resistor enters four stamps as 1 / R: ONE division per parameter instead of one per
  // stamp term (the same quotient, so the sums below have the same bits)
  double* G = k.G + inst;
  for (int i = 0; i < k.nrec; i++) G[(IT)i * (IT)B] = 1. / P(k.rslot[i]);
  for (int q = 0; q < k.npos; q++) {     // deterministic gather, as biglin2
    double v = 0.0;
    for (int t = p.term_ptr[q]; t < p.term_ptr[q + 1]; t++) {
      const int kd = p.term_kind[t];
      const double sg = kd < 0 ? -1.0 : 1.0;
      const int ak = kd < 0 ? -kd : kd;
      if (ak == 1) v += sg * G[(IT)k.rindex[p.term_slot[t]] * (IT)B];
      else if (ak == 2) v += sg * P(p.term_slot[t]);
      else v += sg;
    }
    k.Mv[(IT)q * (IT)B + inst] = v;
    V0[(IT)q * (IT)B] = v + (p.pdiag[q] ? k.o.gmin : 0.0);
    V1[(IT)q * (IT)B] = v;
  }
  for (int q = k.npos; q < k.nslot; q++) { V0[(IT)q * (IT)B] = 0.0; V1[(IT)q * (IT)B] = 0.0; }
  for (int m = 0; m < 2; m++) {
    double* V = m ? V1 : V0;
    for (int kk = 0; kk < n; kk++) {
      const Big3Step st = k.step[m * n + kk];
      const int bi = k.prow[m * n + kk];
      const double piv = V[(IT)st.piv * (IT)B], ap = fabs(piv);
      // dgesv's choice for THIS instance must be the scheduled row
      bool same = true, allzero = ap == 0.0;
      for (int i = 0; i < st.ncand; i++) {
        const double ac = fabs(V[(IT)k.cand_slot[st.cand0 + i] * (IT)B]);
        const int r = k.cand_row[st.cand0 + i];
        if (r < bi ? !(ac < ap) : !(ac <= ap)) same = false;
        if (ac != 0.0) allzero = false;
      }
      if (allzero) return BIG3_SINGULAR;            // every candidate exactly zero (biglin2: best == 0)
      if (!same || !(ap > 0.0)) return BIG3_DYNAMIC;
      const double inv = 1.0 / piv;
      for (int i = 0; i < st.nl; i++) {
        const long long ls = (IT)k.l_slot[st.l0 + i] * (IT)B;
        const double l = V[ls] * inv;
        V[ls] = l;
        const int* up = k.upd + st.upd0 + i * st.nu;
        for (int j = 0; j < st.nu; j++) V[(IT)up[j] * (IT)B] -= l * V[(IT)k.u_slot[st.u0 + j] * (IT)B];
      }
      // the pivot's slot keeps its RECIPROCAL: the substitutions multiply (one division per
      // pivot per instance instead of one per pivot per sweep point and pass; the slot of a used
      // row's pivot is never read by a later elimination step)
      V[(IT)st.piv * (IT)B] = inv;
    }
  }
  return BIG3_OK;
}

// ---- per-(instance, point) solve (one thread = one right-hand side) -------------------------
// tid = pt * B + inst: consecutive threads = consecutive instances (coalesced V / Mv / vary)
// b / bT: the thread's right-hand side and its stride — global scratch ([n][B*ptc], stride T)
// or the block's shared memory ([n][blockDim], stride blockDim): every forward / backward
// substitution step reads what the previous one wrote, so the latency of this array is the
// latency of the kernel.
template <typename IT>
AHK_DEV void big3_solve_one(const Big3K& k, long long tid, double* b, IT bT) {
  const Prog& p = k.p;
  const long long B = k.B, T = B * k.ptc;
  const long long inst = tid % B;
  const int pt = k.pt0 + (int)(tid / B), n = k.n;
  const int fl = k.flag[inst];
  if (fl == BIG3_DYNAMIC) return;
  const long long orow = inst * k.npts + pt;
  if (fl == BIG3_SINGULAR) {
    for (int i = 0; i < n; i++) k.x_out[orow * n + i] = nan("");
    k.iters[orow] = 0; k.status[orow] = AHK_ST_SINGULAR;
    return;
  }
  auto P = [&](int slot) -> double {
    const int v = p.slot_vary[slot];
    return v < 0 ? p.nominal[slot] : k.vary[(IT)v * (IT)B + inst];
  };
  double* x = k.xs + tid;
  for (int r = 0; r < n; r++) x[(IT)r * (IT)T] = 0.0;
  int bad = 0;
  for (int m = 0; m < 2; m++) {
    const double* V = k.V + (IT)m * k.nslot * (IT)B + inst;
    const int* pr = k.prow + m * n;
    // b = -(M x + N): N first (dc_analysis.py:1004-1040), then the mat-vec
    for (int r = 0; r < n; r++) b[(IT)r * bT] = 0.0;
    for (int s = 0; s < p.nsrc; s++) {
      const SrcOp& so = p.src[s];
      // (a source with a waveform enters OP / DC with its dc_value: time is None, VSource.py:92-95)
      const double v = (so.elem == k.src_elem) ? k.sweep[pt] : P(so.pbase);
      if (so.is_v) b[(IT)so.row0 * bT] = v;      // N[row] = -V  =>  -N = V
      else { if (so.row0 >= 0) b[(IT)so.row0 * bT] -= v; if (so.row1 >= 0) b[(IT)so.row1 * bT] += v; }
    }
    if (m == 1) {
      for (int r = 0; r < n; r++) {
        double s = 0.0;
        for (int q = p.row_ptr[r]; q < p.row_ptr[r + 1]; q++)
          s += k.Mv[(IT)q * (IT)B + inst] * x[(IT)p.pcol[q] * (IT)T];
        const double rr = s - b[(IT)r * bT];   // residuo = M.x + N
        b[(IT)r * bT] = -rr;
        if (k.resid) k.resid[orow * n + r] = rr;
      }
    }
    // forward: b[r] -= l * b[pivot row]
    for (int kk = 0; kk < n; kk++) {
      const Big3Step st = k.step[m * n + kk];
      if (st.nl) {
        const double bp = b[(IT)pr[kk] * bT];
        for (int i = 0; i < st.nl; i++)
          b[(IT)k.l_row[st.l0 + i] * bT] -= V[(IT)k.l_slot[st.l0 + i] * (IT)B] * bp;
      }
    }
    // backward; dx[kk] is stored over b[pivot row of kk]
    for (int kk = n - 1; kk >= 0; kk--) {
      const Big3Step st = k.step[m * n + kk];
      double s = 0.0;
      for (int j = 0; j < st.nu; j++)
        s += V[(IT)k.u_slot[st.u0 + j] * (IT)B] * b[(IT)pr[k.u_col[st.u0 + j]] * bT];
      b[(IT)pr[kk] * bT] = (b[(IT)pr[kk] * bT] - s) * V[(IT)st.piv * (IT)B];   // (stored reciprocal)
    }
    // x <- x + dx; gmin_check on pass 2 (results.py:493-523)
    for (int c = 0; c < n; c++) {
      const double xo = x[(IT)c * (IT)T], xn = xo + b[(IT)pr[c] * bT];
      if (m == 1) {
        const bool isv = c < p.nv1;
        const double er = isv ? k.o.ver : k.o.ier, ea = isv ? k.o.vea : k.o.iea;
        if (fabs(xn - xo) > er * fmax(fabs(xo), fabs(xn)) + ea) bad = 1;
      }
      x[(IT)c * (IT)T] = xn;
    }
  }
  for (int c = 0; c < n; c++) k.x_out[orow * n + c] = x[(IT)c * (IT)T];
  k.iters[orow] = 2;
  k.status[orow] = AHK_ST_OK | (bad ? AHK_ST_GMIN_CHECK : 0);
}

#ifdef __CUDACC__
__global__ void __launch_bounds__(128) k_big3_factor(const __grid_constant__ Big3K k) {
  const long long inst = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= k.B) return;
  const int fl = k.idx32 ? big3_factor_one<unsigned>(k, inst) : big3_factor_one<long long>(k, inst);
  k.flag[inst] = fl;
  if (fl == BIG3_DYNAMIC) { const int at = atomicAdd(&k.ovf[0], 1); k.ovf[1 + at] = (int)inst; }
}

__global__ void __launch_bounds__(128) k_big3_solve(const __grid_constant__ Big3K k) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= k.B * k.ptc) return;
  extern __shared__ __align__(16) double big3_sb[];
  if (k.b_smem) {       // right-hand side in shared memory: [n][blockDim]
    double* b = big3_sb + threadIdx.x;
    if (k.idx32) big3_solve_one<unsigned>(k, tid, b, (unsigned)blockDim.x);
    else big3_solve_one<long long>(k, tid, b, (long long)blockDim.x);
  } else {
    double* b = k.bs + tid;
    const long long T = k.B * k.ptc;
    if (k.idx32) big3_solve_one<unsigned>(k, tid, b, (unsigned)T);
    else big3_solve_one<long long>(k, tid, b, T);
  }
}
#endif

}  // namespace ahk
