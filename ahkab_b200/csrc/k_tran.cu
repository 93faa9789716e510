// k_tran.cu — transient kernels (transient.transient_analysis with the
// implicit_euler / trap / gear DF modules), one instantiation per variant.  sm_100a.
#include "kernels.cuh"

namespace ahk {

template <class V>
__global__ void __launch_bounds__(128, (V::NC <= 12) ? AHK_MINB : (V::NC <= 16 ? 3 : 2)) k_tran(const __grid_constant__ KTran k) {
  Ctx c; bool live;
  make_ctx_uniform<V::G>(c, &k.b.p, &k.b.o, &k.b.L, k.b.L.stride, k.b.vary, k.b.B, live);
  if (V::NC == 0 && k.b.Ag) {   // large n: matrix (and, for dense patterns, the value arrays) in global memory
    c.Ag = k.b.Ag + (size_t)c.inst * k.b.Astride;
    if (k.b.L.vglob) c.Vg = c.Ag + k.b.Voff;
  }
  run_tran<V>(c, k.x0, k.a, k.status, live);
}

int launch_tran(const KTran& k, const Launch& L, cudaStream_t s) {
  switch (L.vid) {
#define X(id, NC, R, G) case id: return do_launch(k_tran<Var<NC, R, G> >, k, L, s);
    AHK_REAL_VARIANTS(X)
#undef X
  }
  return (int)cudaErrorInvalidValue;
}

}  // namespace ahk
