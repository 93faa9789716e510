// kernels.cuh — launch-side declarations shared by the kernel translation units
// (k_dc.cu, k_tran.cu, k_ac.cu) and the host API (engine.cu).  Each kernel is a
// persistent-per-instance kernel: a group of G lanes carries its instance through
// the whole analysis (every Newton iteration of both OP passes / every sweep point /
// every time step / every frequency), touching HBM only for the parameter reads and
// the result rows.  See core.cuh for the per-instance algorithm.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "ac.cuh"
#include "entry.cuh"

namespace ahk {

// (tried and dropped: Var<6,2,4> for C4 — one wave instead of 1.7 but 208 vs 202 ms, the
// kernel is issue bound;  Var<10,5,2> for C3 — 1.2 KB of spills at 128 registers, 32 vs 22 ms)
// Kernel variants (id, NC, R, G): NC matrix columns incl. the rhs (0 = run-time n,
// shared-memory fallback), R rows per lane, G lanes per instance.  R*G >= n, NC > n.
#define AHK_REAL_VARIANTS(X)                                                          \
  X(0, 0, 0, 8) X(1, 0, 0, 32)                                                        \
  X(2, 4, 3, 1) X(3, 6, 5, 1)                                                         \
  X(4, 4, 1, 4) X(5, 6, 1, 8) X(6, 8, 2, 4) X(7, 10, 3, 4) X(8, 12, 3, 4)             \
  X(9, 16, 2, 8) X(10, 16, 1, 16) X(11, 24, 3, 8) X(12, 32, 2, 16)
#define AHK_N_REAL_VARIANTS 13
// complex128 rows cost twice the registers: one row per lane beyond n = 3 (two rows per lane,
// Var<16,2,8>, spills 1 KB at 168 registers and measured 2.01 vs 1.89 ms on C1)
#define AHK_AC_VARIANTS(X)                                                            \
  X(0, 0, 0, 8) X(1, 0, 0, 32)                                                        \
  X(2, 4, 3, 1) X(3, 8, 1, 8) X(4, 16, 1, 16) X(5, 24, 1, 32) X(6, 32, 1, 32)
#define AHK_N_AC_VARIANTS 7

#ifndef AHK_MINB
#define AHK_MINB 4   // resident 128-thread blocks per SM the small real variants are compiled for
#endif
// __launch_bounds__(128, min_blocks(NC)) caps the registers so that many blocks fit
inline int real_min_blocks(int NC) { return NC <= 12 ? AHK_MINB : (NC <= 16 ? 3 : 2); }
inline int ac_min_blocks(int NC) { return (NC > 0 && NC <= 16) ? 4 : (NC == 0 ? 3 : 2); }

struct VarInfo { int NC, R, G; };
inline const VarInfo* real_variants() {
  static const VarInfo t[] = {
#define X(id, NC, R, G) {NC, R, G},
      AHK_REAL_VARIANTS(X)
#undef X
  };
  return t;
}
inline const VarInfo* ac_variants() {
  static const VarInfo t[] = {
#define X(id, NC, R, G) {NC, R, G},
      AHK_AC_VARIANTS(X)
#undef X
  };
  return t;
}

struct KBase {
  Prog p; Opts o; Layout L;
  const double* vary; long long B;
  double* Ag;            // large-n layout: [B][n x RS] matrix workspace in global memory, else null
  long long Astride;     // doubles per instance of Ag
  long long Voff;        // Layout::vglob: offset of the Mv / Dv / Bv block inside the instance's workspace
};
struct KDc { KBase b; DcArgs a; };
struct KTran { KBase b; const double* x0; TranArgs a; int* status; };
struct KAc { Prog p; Opts o; AcLayout AL; const double* vary; long long B; AcArgs a; int nchunks; double* Ag; long long Astride, Voff; };

struct Launch { int vid, G, threads, blocks; size_t smem; };

// defined in k_dc.cu / k_tran.cu / k_ac.cu; return cudaError_t as int
int launch_dc(const KDc& k, const Launch& L, cudaStream_t s);
int launch_tran(const KTran& k, const Launch& L, cudaStream_t s);
int launch_ac(const KAc& k, const Launch& L, cudaStream_t s);

#ifdef __CUDACC__
template <typename KernelT, typename Args>
int do_launch(KernelT kern, const Args& args, const Launch& L, cudaStream_t s) {
  if (L.smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
    if (e != cudaSuccess) return (int)e;
  }
  kern<<<L.blocks, L.threads, L.smem, s>>>(args);
  return (int)cudaGetLastError();
}
#endif

}  // namespace ahk
