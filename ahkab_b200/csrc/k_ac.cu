// k_ac.cu — AC kernels (ac.ac_analysis): unit = (instance, frequency chunk).  sm_100a.
#include "kernels.cuh"

// (NC = 16 at 4 blocks / 128 registers spills: measured 1.98 vs 1.89 ms on C1)

namespace ahk {

template <class V>
__global__ void __launch_bounds__(128, (V::NC > 0 && V::NC <= 16) ? 4 : (V::NC == 0 ? 3 : 2)) k_ac(const __grid_constant__ KAc k) {
  Ctx c; long long u;
  if (!make_ctx<V::G>(c, &k.p, &k.o, &k.AL.L, k.AL.stride, k.vary, k.B, k.B * k.nchunks, u)) return;
  c.inst = u / k.nchunks;
  if (V::NC == 0 && k.Ag) {   // large n: one complex matrix (+ value arrays) per (instance, chunk)
    c.Ag = k.Ag + (size_t)u * k.Astride;
    if (k.AL.L.vglob) c.Vg = c.Ag + k.Voff;
  }
  const int ch = (int)(u % k.nchunks);
  AcArgs a = k.a;
  const int per = (a.npts + k.nchunks - 1) / k.nchunks;
  a.f0 = ch * per;
  a.f1 = min(a.npts, a.f0 + per);
  run_ac<V>(c, k.AL, a);
}

int launch_ac(const KAc& k, const Launch& L, cudaStream_t s) {
  switch (L.vid) {
#define X(id, NC, R, G) case id: return do_launch(k_ac<Var<NC, R, G> >, k, L, s);
    AHK_AC_VARIANTS(X)
#undef X
  }
  return (int)cudaErrorInvalidValue;
}

}  // namespace ahk
