// program.h — the compiled, topology-only "stamp program" shared by every
// instance of a batch, plus the shared-memory layout of one instance.
// Built on the host from the ahk_circuit element table (program.cpp); read
// on the device through pointers into one read-only global allocation.
#pragma once
#include "../../include/ahkab_b200.h"

namespace ahk {

#define AHK_SMALL_NMAX 64   // shared-memory engine: pivot mask is one 64-bit word
// general large-n engine (core.cuh: lu_solve_glob): the per-instance VECTORS stay in shared
// memory, the dense matrix lives in a global-memory workspace (L2 resident), one warp per
// instance.  The reference solves any circuit densely (dc_analysis.py:796-822); its own limit
// for switching to sparse storage is 400 unknowns (options.py:68).
#define AHK_BIG_NMAX 512
#define AHK_VGLOB_NPOS 512     // large-n layouts: more matrix positions than this -> value arrays in global memory (read once per Newton iteration; shared memory there is worth more as occupancy)

enum { K_R = 1, K_G, K_VDE, K_E, K_H, K_F, K_C, K_L, K_K };

struct LinOp {      // one linear element's contribution to MNA (Mv) or D (Dv)
  int kind, pbase, aux0, aux1;
  int pos[6];       // indices into the merged position list, -1 = ground
};

struct SrcOp {      // independent source: N, Tt(t), Nac
  int is_v, elem, pbase, tf, has_ac;
  int ppb;          // first PERSISTENT slot of [dc] (+ waveform block)
  int ntf;          // slots of the waveform block (0, 8, or 4 + 2*npts for pwl)
  int row0, row1;   // V: row0 = KVL row; I: rows of n1, n2 (-1 = ground)
};

struct DevOp {      // non-linear device
  int type, pbase, mbase, flags;
  int node[4];      // reduced indices (-1 ground): D n1,n2; MOS d,g,s,b; SW n1,n2,sn1,sn2
  int dout;         // first slot of this device's outputs
  int con;          // first slot of its per-instance constants (EKV), else -1
  int ppb, pmb;     // first PERSISTENT slot of its own / its model's parameters, else -1
  int st;           // first slot of its 2-double state (diode last_vd, switch is_on), else -1
  int active;       // 0: both output nodes grounded -> never evaluated (dc_analysis.py:863)
};

#ifndef AHK_JIT   // specialised build (jit_gen.hpp): Opts / Prog / Layout are compile-time constants
struct Opts {
  double vea, ver, iea, ier, gmin, cmin, hmin;
  double nl_voltages_lock_factor, transient_aposteriori_step_threshold;
  int nl_voltages_lock, nr_damp_first_iters, dc_max_nr_iter, transient_max_nr_iter,
      ac_max_nr_iter, use_standard_solve_method, use_gmin_stepping, use_source_stepping,
      transient_prediction_as_x0, transient_use_aposteriori_step_control,
      transient_max_time_iter, reserved;
};

struct Prog {
  int n, nv1, npos, nlin, nsrc, ntd, ndev, ndout, nlock, nguess, nparams, nac;
  int npersist;                   // parameters staged in shared memory for the whole analysis
  const int* pstage;              // [npersist] global parameter slot of each persistent slot
  const int* gpslot;              // [nguess] persistent slot of each OP-guess term, or -1
  const int* row_ptr;             // [n+1] positions are sorted by (row, col)
  const unsigned short* pcol;     // [npos]
  const unsigned short* pcd;      // [npos] pcol | 0x8000 on node-row diagonals (n < 32768)
  const unsigned short* prowi;    // [npos] row of each position
  const unsigned char* pdiag;     // [npos] node-row diagonal: gets Gmin / cmin
  const int* nl_ptr;              // [npos+1] CSR of device contributions per position
  const short* nl_src;            //   index into dout
  const signed char* nl_sgn;
  const int* tx_ptr;              // [n+1] CSR of device currents per row
  const short* tx_src;
  const signed char* tx_sgn;
  // deterministic gather form of the linear MNA stamps (used by the large-n path):
  // value(pos) = sum over terms of  kind 1: sgn/P(slot)  kind 2: sgn*P(slot)  kind 3: sgn
  const int* term_ptr;            // [npos+1]
  const int* term_slot;
  const signed char* term_kind;   // +-1, +-2, +-3
  const LinOp* lin;
  const SrcOp* src;
  const DevOp* dev;
  const int* lock;                // [nlock][2] raw node numbers (0 = ground)
  const double* guessG;           // [n][nguess]
  const int* gslot;
  const double* gscale;
  const double* gconst;
  const double* nominal;          // [nparams]
  const int* slot_vary;           // [nparams] row of vary[] or -1
};

struct Layout {     // offsets in doubles inside one instance's slice
  int RS;           // row stride of A: NC for a register variant, else odd >= n+1
  int ARS;          // row stride of the staging rows of a register variant (NC - 1)
  int A, x, dx, T, res, Bv, Mv, Dv, Nd, Nc, Ntr, dout, dst, x1, xs, xacc;
  int hx, hdx, ht, C0, pred, bytes;
  int Pv;           // staged parameter values [nparams]
  int dcon;         // per-device constants [ndev][EK_NCONST]
  int pvb;          // pivot-row broadcast buffers [2][NC]
  int xin;          // x at the entry of mdn_solver [n] (restored when the matrix turns out singular)
  int big;          // 1: large-n layout — the matrix is in the global workspace (Ctx::Ag), not at A
  int brhs;         // large-n layout: right-hand side column [n]
  int nzm;          // large-n layout: per-row bitmaps of the non-zero columns, [n][(n + 31) / 32] words
  int czm;          // large-n layout: per-COLUMN bitmaps of the rows with an entry, [n][(n + 31) / 32] words; -1: none (dense pattern)
  int lpan;         // large-n layout: multipliers of the current dense panel, [n][8] — offset in the GLOBAL workspace (from Ctx::Ag)
  int vglob;        // large-n layout of a circuit with many matrix positions (dense patterns): the
                    // per-position value arrays Mv / Dv / Bv live in the global workspace too
                    // (offsets relative to Ctx::Vg), vlen doubles per instance
  int vlen;
  int buflen;
  int stride;       // doubles per instance (padded: stride % 16 == (G*RS) % 16)
};
#endif  // AHK_JIT

}  // namespace ahk
