// biglin2.cuh — large linear circuits (n > 64), second generation: OP and DC sweep,
// one WARP per instance, the sparse factorisation entirely in shared memory.
//
// Same mathematics as biglin.cuh (read its header): per instance the two matrices
// mna + Gmin and mna are factored ONCE by right-looking Gaussian elimination with
// dgesv's pivot rule, skipping zero multipliers and zero pivot-row entries (bit-identical
// to the dense algorithm, a - 0*b == a); every sweep point then costs two pairs of
// triangular solves.  What changed is where the work happens:
//
//  * factorisation: each row is a small unordered slot list (column, value) in shared
//    memory and each column a small list of the rows that hold an entry in it.  A pivot
//    step reads the column's list, looks the few candidate values up in their rows, and
//    updates / appends the few fill-ins — a few hundred cycles, no global memory, where
//    biglin.cuh scanned a dense n-vector out of L2 per step (C5: 2 x 257 steps x 257 rows).
//    A row or column that outgrows its slots sends the instance to the biglin.cuh kernel
//    (overflow list, second launch that exits at once when the list is empty).
//  * shared memory per warp bounds the warps resident per SM, and the kernel is latency
//    bound: 52 KB (row slots 12 deep, column bitmaps, iterates in shared memory) gave 4
//    warps/SM and 10.9 ms on C5; 30 KB (slots 6 deep, column lists, iterates in x_out) gives
//    7 warps/SM = exactly 4 waves of the 4096 instances and 6.3 ms.
//  * solves: the sweep points are independent for a linear circuit (the reference warm
//    starts point k from point k-1, which only changes the start of an exact one-step
//    Newton), so up to PC points are solved AT ONCE, one lane per right-hand side, each
//    lane running its own forward / backward substitution with no warp synchronisation;
//    L and U are streamed from the per-warp global lists (warp-uniform loads); the
//    iterates live in their rows of x_out (L2), only the right-hand sides in shared memory.
#pragma once
#include "biglin.cuh"

namespace ahk {

struct Big2K {
  Prog p; Opts o;
  const double* vary; long long B;
  int src_elem; const double* sweep; int npts;
  const double* x0;
  double* x_out; int* iters; int* status; double* resid;
  double* Lv; double* Uv; unsigned short* Li; unsigned short* Ui;   // [nwarps][2][cap2]
  int cap2, SC, PC, W;
  int* ovf;   // ovf[0] = count, ovf[1..] = instance list
};

struct Big2Smem {   // byte offsets inside one warp's dynamic shared memory
  int Mv, prow, nLc, nUc, un;   // fixed part, then the union
  int scol, sval, scnt, cbits, used, pend_r, pend_v, pu_c, pu_v;   // factorisation
  int b, x1;                                                       // solves
  int total;
};

inline Big2Smem big2_smem(int n, int npos, int SC, int PC, int W) {
  Big2Smem s; int off = 0;
  auto take = [&](int bytes) { int o = off; off += (bytes + 15) & ~15; return o; };
  s.Mv = take(8 * npos); s.prow = take(2 * 2 * n); s.nLc = take(2 * 2 * n); s.nUc = take(2 * 2 * n);
  s.un = off;
  // (cbits = the column lists: SC row indices per column + a 32-bit count per column)
  s.sval = take(8 * n * SC); s.pend_v = take(8 * 32); s.pu_v = take(8 * (SC + 1));
  s.scol = take(2 * n * SC); s.scnt = take(n); s.cbits = take(2 * n * SC + 4 * n); s.used = take(4 * W);
  s.pend_r = take(2 * 32); s.pu_c = take(2 * (SC + 1));
  int fac_end = off;
  off = s.un;
  s.b = take(8 * n * PC); s.x1 = s.b;   // the iterates live in x_out (global)
  s.total = off > fac_end ? off : fac_end;
  return s;
}

#ifdef __CUDACC__

__global__ void __launch_bounds__(32) k_biglin2(const __grid_constant__ Big2K k, const Big2Smem S) {
  extern __shared__ __align__(16) unsigned char bsm[];
  const Prog& p = k.p;
  const int n = p.n, lane = threadIdx.x, SC = k.SC, W = k.W;
  const unsigned FULL = 0xffffffffu, lt = (1u << lane) - 1u;
  double* Mv = (double*)(bsm + S.Mv);
  unsigned short* prow = (unsigned short*)(bsm + S.prow);
  unsigned short* nLc = (unsigned short*)(bsm + S.nLc);
  unsigned short* nUc = (unsigned short*)(bsm + S.nUc);
  double* sval = (double*)(bsm + S.sval); double* pend_v = (double*)(bsm + S.pend_v);
  double* pu_v = (double*)(bsm + S.pu_v);
  unsigned short* scol = (unsigned short*)(bsm + S.scol);
  unsigned char* scnt = bsm + S.scnt;
  int* ccnt = (int*)(bsm + S.cbits);                                  // [n] entries per column
  unsigned short* ccol = (unsigned short*)(bsm + S.cbits + 4 * n);    // [n][SC] rows of a column
  unsigned* used = (unsigned*)(bsm + S.used);
  unsigned short* pend_r = (unsigned short*)(bsm + S.pend_r);
  unsigned short* pu_c = (unsigned short*)(bsm + S.pu_c);
  double* bb = (double*)(bsm + S.b);
  double* Lv = k.Lv + (size_t)blockIdx.x * 2 * k.cap2;
  double* Uv = k.Uv + (size_t)blockIdx.x * 2 * k.cap2;
  unsigned short* Li = k.Li + (size_t)blockIdx.x * 2 * k.cap2;
  unsigned short* Ui = k.Ui + (size_t)blockIdx.x * 2 * k.cap2;

  for (long long inst = blockIdx.x; inst < k.B; inst += gridDim.x) {
    auto P = [&](int slot) -> double {
      int v = p.slot_vary[slot];
      return v < 0 ? p.nominal[slot] : k.vary[(long long)v * k.B + inst];
    };
    __syncwarp();
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
    int singular = 0, ovf = 0;
    // ---- factor mna + Gmin (m = 0) and mna (m = 1) -------------------------------------
    for (int m = 0; m < 2 && !singular && !ovf; m++) {
      for (int r = lane; r < n; r += 32) {
        const int q0 = p.row_ptr[r], c = p.row_ptr[r + 1] - q0;
        if (c > SC) { ovf = 1; continue; }
        scnt[r] = (unsigned char)c;
        for (int t = 0; t < c; t++) {
          scol[r * SC + t] = p.pcol[q0 + t];
          sval[r * SC + t] = Mv[q0 + t] + ((m == 0 && p.pdiag[q0 + t]) ? k.o.gmin : 0.0);
        }
      }
      for (int i = lane; i < n; i += 32) ccnt[i] = 0;
      if (lane < W) used[lane] = 0u;
      ovf = __any_sync(FULL, ovf);
      if (ovf) break;
      __syncwarp();
      for (int q = lane; q < p.npos; q += 32) {
        const int col = p.pcol[q];
        const int t = atomicAdd(&ccnt[col], 1);
        if (t < SC) ccol[col * SC + t] = p.prowi[q]; else ovf = 1;
      }
      ovf = __any_sync(FULL, ovf);
      if (ovf) break;
      __syncwarp();
      int Lbase = m * k.cap2, Ubase = m * k.cap2;
      for (int kk = 0; kk < n; kk++) {
        // candidate rows: the column's list, minus the rows that are pivot rows already;
        // lane t looks candidate t up in its row
        const int ncol = ccnt[kk];
        int my_r = -1;
        double my_v = 0.0;
        if (lane < ncol) {
          const int r = ccol[kk * SC + lane];
          if (!((used[r >> 5] >> (r & 31)) & 1u)) {
            my_r = r;
            const int c = scnt[r];
            for (int t = 0; t < c; t++) if (scol[r * SC + t] == kk) { my_v = sval[r * SC + t]; break; }
          }
        }
        const unsigned cm = __ballot_sync(FULL, my_r >= 0);
        const int ncand = __popc(cm);
        if (my_r >= 0) { const int i = __popc(cm & lt); pend_r[i] = (unsigned short)my_r; pend_v[i] = my_v; }
        double best = my_r >= 0 ? fabs(my_v) : -1.0, sv = my_v;
        int bi = my_r;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double ov = __shfl_xor_sync(FULL, best, o), osv = __shfl_xor_sync(FULL, sv, o);
          const int oi = __shfl_xor_sync(FULL, bi, o);
          const bool take = (oi >= 0) && (bi < 0 || ov > best || (ov == best && oi < bi));
          if (take) { best = ov; bi = oi; sv = osv; }
        }
        __syncwarp();
        if (bi < 0 || best == 0.0) { singular = 1; break; }
        if (Lbase + ncand > (m + 1) * k.cap2 || Ubase + SC + 1 > (m + 1) * k.cap2) { ovf = 1; break; }
        const double piv = sv, inv = 1.0 / piv;
        if (lane == 0) { prow[m * n + kk] = (unsigned short)bi; used[bi >> 5] |= 1u << (bi & 31); }
        // L: rows with a non-zero multiplier (ascending row order), compacted in place
        int nL = 0;
        for (int i0 = 0; i0 < ncand; i0 += 32) {
          const int i = i0 + lane;
          int r = 0; double v = 0.0; bool f = false;
          if (i < ncand) { r = pend_r[i]; v = pend_v[i]; f = r != bi && v != 0.0; }
          const unsigned bal = __ballot_sync(FULL, f);
          if (f) {
            const int j = nL + __popc(bal & lt);
            const double l = v * inv;
            pend_r[j] = (unsigned short)r; pend_v[j] = l;
            Li[Lbase + j] = (unsigned short)r; Lv[Lbase + j] = l;
          }
          nL += __popc(bal);
        }
        // U: the pivot first, then the pivot row's non-zero entries right of it, by column
        int nU = 0;
        {
          const int c = scnt[bi];
          int col = 0; double u = 0.0; bool f = false;
          if (lane < c) { col = scol[bi * SC + lane]; u = sval[bi * SC + lane]; f = col > kk && u != 0.0; }
          const unsigned bal = __ballot_sync(FULL, f);
          nU = __popc(bal);
          if (f) {
            int rank = 0;
            for (int t = 0; t < c; t++) {
              const int ct = scol[bi * SC + t];
              if (ct > kk && ct < col && sval[bi * SC + t] != 0.0) rank++;
            }
            pu_c[rank] = (unsigned short)col; pu_v[rank] = u;
            Ui[Ubase + 1 + rank] = (unsigned short)col; Uv[Ubase + 1 + rank] = u;
          }
          if (lane == 0) { Ui[Ubase] = (unsigned short)kk; Uv[Ubase] = piv; nLc[m * n + kk] = (unsigned short)nL; nUc[m * n + kk] = (unsigned short)nU; }
        }
        __syncwarp();
        // elimination: one lane per L row
        for (int i = lane; i < nL; i += 32) {
          const int r = pend_r[i];
          const double l = pend_v[i];
          int c = scnt[r];
          for (int t = 0; t < c; t++)           // drop the (r, kk) entry: its column is done
            if (scol[r * SC + t] == kk) { c--; scol[r * SC + t] = scol[r * SC + c]; sval[r * SC + t] = sval[r * SC + c]; break; }
          for (int j = 0; j < nU; j++) {
            const int col = pu_c[j];
            const double u = pu_v[j];
            int t = 0;
            for (; t < c; t++) if (scol[r * SC + t] == col) break;
            if (t < c) sval[r * SC + t] -= l * u;
            else if (c < SC) {
              scol[r * SC + c] = (unsigned short)col; sval[r * SC + c] = 0.0 - l * u; c++;
              const int t = atomicAdd(&ccnt[col], 1);
              if (t < SC) ccol[col * SC + t] = (unsigned short)r; else ovf = 1;
            } else ovf = 1;
          }
          scnt[r] = (unsigned char)c;
        }
        Lbase += nL; Ubase += nU + 1;
        ovf = __any_sync(FULL, ovf);
        if (ovf) break;
        __syncwarp();
      }
    }
    if (ovf) {   // hand the instance to the dense-workspace kernel
      if (lane == 0) { const int at = atomicAdd(&k.ovf[0], 1); k.ovf[1 + at] = (int)inst; }
      continue;
    }
    __syncwarp();
    // ---- solves: PC sweep points at once, one lane per right-hand side ---------------------
    for (int k0 = 0; k0 < k.npts; k0 += k.PC) {
      const int pc = min(k.PC, k.npts - k0), PC = k.PC;
      const bool act = lane < pc;
      const long long orow = inst * k.npts + k0 + lane;
      if (singular) {
        for (int j = 0; j < pc; j++) {
          for (int i = lane; i < n; i += 32) k.x_out[(inst * k.npts + k0 + j) * n + i] = nan("");
          if (lane == 0) { k.iters[inst * k.npts + k0 + j] = 0; k.status[inst * k.npts + k0 + j] = AHK_ST_SINGULAR; }
        }
        continue;
      }
      // The iterate of each point lives in its own row of x_out (global, L2-resident):
      // shared memory only holds the right-hand sides.  x <- x0 (zeros if none), coalesced.
      double* xg = k.x_out + (inst * k.npts + k0) * n;      // [pc][n]
      for (int j = 0; j < pc; j++)
        for (int i = lane; i < n; i += 32) xg[(long long)j * n + i] = k.x0 ? k.x0[inst * n + i] : 0.0;
      __syncwarp();
      int bad = 0;
      for (int m = 0; m < 2; m++) {
        const unsigned short* pr = prow + m * n;
        if (act) {
          const int j = lane;
          const double* xj = xg + (long long)j * n;
          // b = -(M x + N): N first (dc_analysis.py:1004-1040), then the mat-vec
          for (int r = 0; r < n; r++) bb[r * PC + j] = 0.0;
          for (int s = 0; s < p.nsrc; s++) {
            const SrcOp& so = p.src[s];
            // (a source with a waveform enters OP / DC with its dc_value: time is None, VSource.py:92-95)
            const double v = (so.elem == k.src_elem) ? k.sweep[k0 + j] : P(so.pbase);
            if (so.is_v) bb[so.row0 * PC + j] = v;      // N[row] = -V  =>  -N = V
            else { if (so.row0 >= 0) bb[so.row0 * PC + j] -= v; if (so.row1 >= 0) bb[so.row1 * PC + j] += v; }
          }
          if (m == 1 || k.x0) {
            for (int r = 0; r < n; r++) {
              double s = 0.0;
              for (int q = p.row_ptr[r]; q < p.row_ptr[r + 1]; q++)
                s += (Mv[q] + ((m == 0 && p.pdiag[q]) ? k.o.gmin : 0.0)) * xj[p.pcol[q]];
              // residuo = M.x + N = s - (-N)
              const double rr = s - bb[r * PC + j];
              bb[r * PC + j] = -rr;
              if (m == 1 && k.resid) k.resid[orow * n + r] = rr;
            }
          }
          // forward: b[r] -= l * b[pivot row]
          int e = m * k.cap2;
          for (int kk = 0; kk < n; kk++) {
            const int cnt = nLc[m * n + kk];
            if (cnt) {
              const double bp = bb[(int)pr[kk] * PC + j];
              if (bp != 0.0)
                for (int t = 0; t < cnt; t++) bb[(int)Li[e + t] * PC + j] -= Lv[e + t] * bp;
              e += cnt;
            }
          }
          // backward; dx[kk] is stored over b[pivot row of kk] (that entry is dead by then)
          int ue = m * k.cap2;
          for (int kk = 0; kk < n; kk++) ue += nUc[m * n + kk] + 1;
          for (int kk = n - 1; kk >= 0; kk--) {
            const int cnt = nUc[m * n + kk];
            ue -= cnt + 1;
            double s = 0.0;
            for (int t = 1; t <= cnt; t++) s += Uv[ue + t] * bb[(int)pr[Ui[ue + t]] * PC + j];
            bb[(int)pr[kk] * PC + j] = (bb[(int)pr[kk] * PC + j] - s) / Uv[ue];
          }
        }
        __syncwarp();
        // x <- x + dx (td = 1: a linear circuit has no locked nodes), all lanes, coalesced;
        // gmin_check on pass 2 (results.py:493-523)
        for (int j = 0; j < pc; j++) {
          double* xj = xg + (long long)j * n;
          int badj = 0;
          for (int c = lane; c < n; c += 32) {
            const double xo = xj[c], xn = xo + bb[(int)pr[c] * PC + j];
            if (m == 1) {
              const bool isv = c < p.nv1;
              const double er = isv ? k.o.ver : k.o.ier, ea = isv ? k.o.vea : k.o.iea;
              if (fabs(xn - xo) > er * fmax(fabs(xo), fabs(xn)) + ea) badj = 1;
            }
            xj[c] = xn;
          }
          badj = __any_sync(FULL, badj);
          if (lane == j && badj) bad = 1;
        }
        __syncwarp();
      }
      if (act) { k.iters[orow] = 2; k.status[orow] = AHK_ST_OK | (bad ? AHK_ST_GMIN_CHECK : 0); }
      __syncwarp();
    }
  }
}
#endif  // __CUDACC__

}  // namespace ahk
