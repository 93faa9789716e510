// hostemu_jit.cpp — TEST-ONLY.  Host build of the engine's per-instance source in its
// SPECIALISED form: compiled together with a generated jit_spec.h (jit_gen.hpp, found
// through -I), so that Prog / Opts / Layout are the compile-time constants the NVRTC
// build of the same circuit sees.  tests/test_jit.py compares it bit for bit with the
// interpreted host build (hostemu.cpp).  One lane per instance (G = 1).
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>
using std::isnan; using std::isinf; using std::isfinite;
#include "jit_spec.h"
#include "../../ahkab_b200/csrc/entry.cuh"

using namespace ahk;

static Prog g_p; static Opts g_o; static Layout g_L;
typedef Var<AHK_JIT_NC, AHK_JIT_R, AHK_JIT_G> V;

static Ctx make(std::vector<double>& sm, const double* vary, long long B, long long b) {
  Ctx c; c.p = &g_p; c.o = &g_o; c.L = &g_L; c.sm = sm.data(); c.vary = vary; c.B = B; c.inst = b;
  c.sub = 0; c.G = 1; c.mask = 1u; return c;
}

extern "C" {

int emuj_dc(int64_t B, const double* vary, int src_elem, const double* sweep, int npts, const double* x0,
            int use_guess, double* x, double* resid, int32_t* iters, int32_t* status) {
  static_assert(AHK_JIT_G == 1, "host build: one lane per instance");
  std::vector<double> sm(Layout::stride + 64, 0.0);
  DcArgs a = {src_elem, sweep, npts, x0, use_guess, x, resid, iters, status};
  for (int64_t b = 0; b < B; b++) { Ctx c = make(sm, vary, B, b); run_dc<V>(c, a); }
  return 0;
}

int emuj_tran(int64_t B, const double* vary, const double* x0, double tstart, double tstep, double tstop,
              int64_t max_rows, double* t_out, double* x_out, int32_t* rows, int32_t* counters,
              int32_t* status) {
  std::vector<double> sm(Layout::stride + 64, 0.0);
  TranArgs a = {tstart, tstep, tstop, AHK_JIT_METHOD, AHK_JIT_USC, max_rows, t_out, x_out, rows, counters};
  for (int64_t b = 0; b < B; b++) { Ctx c = make(sm, vary, B, b); run_tran<V>(c, x0, a, status); }
  return 0;
}

}  // extern "C"
