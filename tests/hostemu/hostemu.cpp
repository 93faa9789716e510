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

using namespace ahk;

static void copy_opts(const ahk_options* o, Opts& q) { std::memcpy(&q, o, sizeof(Opts)); }

struct Emu {
  HostProg H; Prog p; Opts o; Layout L; std::vector<double> sm;
  Emu(const ahk_circuit* c, const ahk_options* opts, int n_vary, const int32_t* slots, int mode,
      int method, int usc) {
    build_program(c, H);
    set_vary_slots(H, n_vary, slots);
    p = H.view();
    copy_opts(opts, o);
    L = make_layout(H, mode, method, usc, 1);
    sm.assign(L.stride + 64, 0.0);
  }
  Ctx ctx(const double* vary, long long B, long long b) {
    Ctx c; c.p = &p; c.o = &o; c.L = &L; c.sm = sm.data(); c.vary = vary; c.B = B; c.inst = b;
    c.sub = 0; c.mask = 1u; return c;
  }
};

extern "C" {

int emu_op(const ahk_circuit* c, const ahk_options* o, int64_t B, int n_vary, const int32_t* slots,
           const double* vary, const double* x0, int use_guess, double* x, double* resid,
           int32_t* iters, int32_t* status) {
  Emu e(c, o, n_vary, slots, MODE_OP, 0, 0);
  OpArgs a = {x0, use_guess, x, resid, iters, status};
  for (int64_t b = 0; b < B; b++) { Ctx cx = e.ctx(vary, B, b); run_op<1>(cx, a); }
  return 0;
}

int emu_dc_sweep(const ahk_circuit* c, const ahk_options* o, int64_t B, int n_vary,
                 const int32_t* slots, const double* vary, int src_elem, const double* sweep,
                 int npts, const double* x0, int use_guess, double* x_out, int32_t* iters,
                 int32_t* status) {
  Emu e(c, o, n_vary, slots, MODE_OP, 0, 0);
  DcArgs a = {src_elem, sweep, npts, x0, use_guess, x_out, iters, status};
  for (int64_t b = 0; b < B; b++) { Ctx cx = e.ctx(vary, B, b); run_dc<1>(cx, a); }
  return 0;
}

int emu_tran(const ahk_circuit* c, const ahk_options* o, int64_t B, int n_vary, const int32_t* slots,
             const double* vary, const double* x0, double tstart, double tstep, double tstop,
             int method, int usc, int64_t max_rows, double* t_out, double* x_out, int32_t* rows,
             int32_t* counters, int32_t* status) {
  Emu e(c, o, n_vary, slots, MODE_TRAN, method, usc);
  TranArgs a = {tstart, tstep, tstop, method, usc, max_rows, t_out, x_out, rows, counters};
  for (int64_t b = 0; b < B; b++) { Ctx cx = e.ctx(vary, B, b); run_tran<1>(cx, x0, a, status); }
  return 0;
}

int emu_ac(const ahk_circuit* c, const ahk_options* o, int64_t B, int n_vary, const int32_t* slots,
           const double* vary, const double* x_op, const double* omegas, int npts, double* x_out,
           int32_t* status) {
  Emu e(c, o, n_vary, slots, MODE_OP, 0, 0);
  AcLayout AL = make_ac_layout(e.H, 1);
  std::vector<double> sm(AL.stride + 64, 0.0);
  AcArgs a = {x_op, omegas, npts, 0, npts, x_out, status};
  for (int64_t b = 0; b < B; b++) {
    status[b] = AHK_ST_OK;
    Ctx cx = e.ctx(vary, B, b);
    cx.sm = sm.data();
    run_ac<1>(cx, AL, a);
  }
  return 0;
}

}  // extern "C"
