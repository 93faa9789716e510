// jit_kernels.cuh — entry points of the circuit-specialised kernels (see jit_gen.hpp).
// Compiled at run time by NVRTC for sm_100a as
//     #include "jit_spec.h"       (generated: Prog / Opts / Layout as constants)
//     #include "jit_kernels.cuh"
// One translation unit = one (circuit, options, analysis, kernel variant).  The
// per-instance algorithm is the one of the interpreted kernels: entry.cuh / core.cuh.
#pragma once
#ifndef AHK_JIT
#error "jit_kernels.cuh needs the generated jit_spec.h first"
#endif
#include "entry.cuh"
#if AHK_JIT_MODE == 2
#include "ac.cuh"
#endif

namespace ahk {

#ifdef __CUDACC__
__device__ const Prog jit_prog{};
__device__ const Opts jit_opts{};
__device__ const Layout jit_layout{};
typedef Var<AHK_JIT_NC, AHK_JIT_R, AHK_JIT_G> JitVar;

#if AHK_JIT_MODE == 0
extern "C" __global__ void __launch_bounds__(AHK_JIT_THREADS, AHK_JIT_MINB)
ahk_jit_dc(const __grid_constant__ KDcJ k) {
  if constexpr (JitVar::G == 1) {
    if (k.a.npts == 1) {          // operating point, one lane per instance: coalesced result rows
      Ctx c; bool live;
      make_ctx_uniform<1>(c, &jit_prog, &jit_opts, &jit_layout, Layout::stride, k.vary, k.B, live);
      c.pivrec = live ? k.pivrec : nullptr;
      run_op_rows<JitVar>(c, k.a, live);
      return;
    }
  }
  Ctx c; long long u;
  if (!make_ctx<JitVar::G>(c, &jit_prog, &jit_opts, &jit_layout, Layout::stride, k.vary, k.B, k.B, u)) return;
  c.pivrec = k.pivrec;
  run_dc<JitVar>(c, k.a);
}
#elif AHK_JIT_MODE == 2
__device__ const AcLayout jit_aclayout{};
extern "C" __global__ void __launch_bounds__(AHK_JIT_THREADS, AHK_JIT_MINB)
ahk_jit_ac(const __grid_constant__ KAcJ k) {   // unit = (instance, frequency chunk), as k_ac.cu
  Ctx c; long long u;
#ifdef AHK_JIT_SLU
  // static-schedule kernel: unit = instance, the G lanes of its group are its frequency chunks
  if (!make_ctx<JitVar::G>(c, &jit_prog, &jit_opts, &jit_aclayout.L, AcLayout::stride, k.vary, k.B, k.B, u)) return;
  c.pivrec = k.pivrec;
  run_ac<JitVar>(c, jit_aclayout, k.a);
#else
  if (!make_ctx<JitVar::G>(c, &jit_prog, &jit_opts, &jit_aclayout.L, AcLayout::stride, k.vary, k.B,
                           k.B * k.nchunks, u)) return;
  c.inst = u / k.nchunks;
  const int ch = (int)(u % k.nchunks);
  AcArgs a = k.a;
  const int per = (a.npts + k.nchunks - 1) / k.nchunks;
  a.f0 = ch * per;
  a.f1 = min(a.npts, a.f0 + per);
  run_ac<JitVar>(c, jit_aclayout, a);
#endif
}
#else
extern "C" __global__ void __launch_bounds__(AHK_JIT_THREADS, AHK_JIT_MINB)
ahk_jit_tran(const __grid_constant__ KTranJ k) {
  Ctx c; bool live;
  make_ctx_uniform<JitVar::G>(c, &jit_prog, &jit_opts, &jit_layout, Layout::stride, k.vary, k.B, live);
  c.pivrec = live ? k.pivrec : nullptr;   // (padding groups repeat the last instance: they must not record)
  run_tran<JitVar>(c, k.x0, k.a, k.status, live);
}
#endif
#endif

}  // namespace ahk
