// k_dc.cu — operating point / DC sweep kernels (dc_analysis.op_analysis,
// dc_analysis.dc_analysis), one instantiation per kernel variant.  sm_100a.
#include "kernels.cuh"

namespace ahk {

template <class V>
__global__ void __launch_bounds__(128, (V::NC <= 12) ? AHK_MINB : (V::NC <= 16 ? 3 : 2)) k_dc(const __grid_constant__ KDc k) {
  Ctx c; long long u;
  if (!make_ctx<V::G>(c, &k.b.p, &k.b.o, &k.b.L, k.b.L.stride, k.b.vary, k.b.B, k.b.B, u)) return;
  if (V::NC == 0 && k.b.Ag) {   // large n: matrix (and, for dense patterns, the value arrays) in global memory
    c.Ag = k.b.Ag + (size_t)c.inst * k.b.Astride;
    if (k.b.L.vglob) c.Vg = c.Ag + k.b.Voff;
  }
  run_dc<V>(c, k.a);
}

int launch_dc(const KDc& k, const Launch& L, cudaStream_t s) {
  switch (L.vid) {
#define X(id, NC, R, G) case id: return do_launch(k_dc<Var<NC, R, G> >, k, L, s);
    AHK_REAL_VARIANTS(X)
#undef X
  }
  return (int)cudaErrorInvalidValue;
}

}  // namespace ahk
