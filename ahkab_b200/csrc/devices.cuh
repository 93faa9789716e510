// devices.cuh — non-linear device evaluators and source waveforms, FP64.
//
// One call = one evaluation of a device at given port voltages, returning the
// current and the small-signal conductances the Newton Jacobian needs.  The
// formulas (including their quirks) are the reference's:
//   diode   ahkab/diode.py:352-424      EKV     ahkab/ekv.py:444-752
//   mosq    ahkab/mosq.py:552-803       switch  ahkab/switch.py:310-399
//   sources ahkab/time_functions.py:402-745
// Unlike the reference (which re-runs the whole EKV current computation four
// times per Newton iteration, ekv.py:676,687,698), each device is evaluated
// once and every output is derived from that single evaluation; the EKV inner
// solve is stateless (no warm-start state) and converged to ~1e-12.
#pragma once
#ifndef __CUDACC_RTC__
#include <math.h>
#else   // NVRTC: the math functions are built in, the macros of math.h are not
#define INFINITY __builtin_bit_cast(double, 0x7ff0000000000000ULL)
#endif

#ifdef __CUDACC__
#define AHK_DEV __device__ __forceinline__
#define AHK_MEM __device__ __forceinline__
// device models: inlined.  (-DAHK_NOINLINE_DEV makes them real functions, one copy per
// kernel: smaller code, but measured slower — C2 0.116 vs 0.109 ms, C4 213 vs 207 ms.)
#ifdef AHK_NOINLINE_DEV
#define AHK_DEVFN static __device__ __noinline__
#else
#define AHK_DEVFN static __device__ __forceinline__
#endif
#else
#define AHK_DEVFN static inline
#define AHK_DEV static inline
#define AHK_MEM inline
#endif

#include "fastmath.cuh"

namespace ahk {

// constants.py:30-53 — k*T/e at T = Tref = 300 K
#define AHK_VTH (1.3806503e-23 * 300 / 1.60217646e-19)
#define AHK_EPS 2.220446049250313e-16
#define AHK_ESI 104.5e-12

// ---- diode -----------------------------------------------------------------
AHK_DEV double safe_exp(double x) { return x < 70 ? fm::exp(x) : 2.5154386709191670e30 + 10 * x; }   // exp(70) = 2.515...e30

struct DiodeP { double IS, N, NBV, ISR, NR, RS, BV, IKF, VT, AREA; };

// Per-instance constants of one diode: the reciprocals of the thermal-voltage products are
// taken once when the instance is loaded (the reference divides at every evaluation).
enum { DK_IS, DK_INVT, DK_INBVVT, DK_IVT, DK_ISR, DK_INRVT, DK_RS, DK_BV, DK_IKF, DK_AREA, DK_NCONST };

AHK_DEVFN void diode_setup(const DiodeP& m, double* k) {
  k[DK_IS] = m.IS; k[DK_INVT] = 1.0 / (m.N * m.VT); k[DK_INBVVT] = 1.0 / (m.NBV * m.VT);
  k[DK_IVT] = 1.0 / m.VT; k[DK_ISR] = m.ISR; k[DK_INRVT] = 1.0 / (m.NR * m.VT);
  k[DK_RS] = m.RS; k[DK_BV] = m.BV; k[DK_IKF] = m.IKF; k[DK_AREA] = m.AREA;
}

// diode.py:387-395.  BV = inf (the default): exp(-inf) = 0 exactly, no call needed.
AHK_DEV double diode_i_raw(const double* k, double v, double efwd) {
  const double IS = k[DK_IS];
  double i_fwd = IS * (efwd - 1);
  double i_rec = k[DK_ISR] != 0.0 ? k[DK_ISR] * (safe_exp(v * k[DK_INRVT]) - 1) : 0.0;
  double erev = isinf(k[DK_BV]) ? (k[DK_BV] > 0 ? 0.0 : safe_exp(INFINITY)) : safe_exp(-(v + k[DK_BV]) * k[DK_INBVVT]);
  double i_rev = -IS * (erev - 1);
  double k_inj = 1;
  const double IKF = k[DK_IKF];
  if (!isinf(IKF) && IKF > 0 && i_fwd > 0) k_inj = sqrt(IKF / (IKF + i_fwd));
  return k_inj * i_fwd + i_rec + i_rev;
}

// diode.py:402-406 (not d i/dv: no NBV, sign of the breakdown term, no k_inj)
AHK_DEV double diode_gm_raw(const double* k, double v, double efwd) {
  const double IS = k[DK_IS];
  double erev = isinf(k[DK_BV]) ? (k[DK_BV] > 0 ? 0.0 : safe_exp(INFINITY)) : safe_exp(-(v + k[DK_BV]) * k[DK_IVT]);
  double g = IS * k[DK_INVT] * efwd + -IS * k[DK_IVT] * erev;
  if (k[DK_ISR] != 0.0) g += k[DK_ISR] * k[DK_INRVT] * safe_exp(v * k[DK_INRVT]);
  return g;
}

// out[0] = i, out[1] = gm (with the gm==0 -> 2*gmin rule of diode.py:178-179);
// *last_vd is the device state used by the RS != 0 inner Newton (diode.py:359-365).
AHK_DEVFN void diode_eval(const double* k, double v, double vea, double gmin,
                        double* last_vd, double* out) {
  const double RS = k[DK_RS], AREA = k[DK_AREA];
  const double efwd = safe_exp(v * k[DK_INVT]);     // shared by the current and the conductance
  double gm = diode_gm_raw(k, v, efwd);
  if (RS != 0.0) gm = 1. / (RS + 1. / (gm + 1e-3 * gmin));
  gm *= AREA;
  if (gm == 0) gm = gmin * 2;
  double i;
  if (!(RS != 0.0)) {
    i = diode_i_raw(k, v, efwd) * AREA;
    *last_vd = v;
  } else {
    // scipy.optimize.newton (Newton branch), tol = vea, x0 = last_vd
    double p0 = *last_vd, p = p0;
    for (int it = 0; it < 500; it++) {
      const double e0 = safe_exp((v - p0) * k[DK_INVT]);
      double fval = p0 / RS - diode_i_raw(k, v - p0, e0) * AREA;
      if (fval == 0) { p = p0; break; }
      double fder = 1. / RS + AREA * diode_gm_raw(k, v - p0, e0);
      if (fder == 0) { p = p0; break; }
      p = p0 - fval / fder;
      if (fabs(p - p0) <= vea) break;
      p0 = p;
    }
    *last_vd = p;
    i = diode_i_raw(k, v - p, safe_exp((v - p) * k[DK_INVT]));  // AREA dropped, as diode.py:363 does
  }
  out[0] = i;
  out[1] = gm;
}

// ---- EKV ---------------------------------------------------------------------
struct EkvP { double VTO, KP, GAMMA, PHI, COX, LAMBDA, XJ, UCRIT, W, L, M, N; int NP; };

#ifdef __CUDACC__
AHK_DEV double ahk_rsqrt(double x) { return rsqrt(x); }
#else
AHK_DEV double ahk_rsqrt(double x) { return 1.0 / sqrt(x); }
#endif

// Per-instance constants of one EKV device (model/geometry only), computed once
// when the instance is loaded instead of at every evaluation.
enum { EK_VTO, EK_GAMMA, EK_PHI, EK_SQPHI, EK_T, EK_KWL, EK_VC, EK_UTVC, EK_LVC, EK_LAMBDA,
       EK_LLC, EK_ILCU, EK_NL, EK_LMIN2, EK_IUCRIT, EK_PREF, EK_NCONST };

AHK_DEVFN void ekv_setup(const EkvP& m, double* k) {
  const double Ut = AHK_VTH;
  k[EK_VTO] = m.VTO; k[EK_GAMMA] = m.GAMMA; k[EK_PHI] = m.PHI;
  k[EK_SQPHI] = sqrt(m.PHI);
  k[EK_T] = k[EK_SQPHI] + m.GAMMA * 0.5;
  k[EK_KWL] = m.KP * m.W / m.L;
  const double Vc = m.UCRIT * m.N * m.L;                 // ekv.py:640
  k[EK_VC] = Vc; k[EK_UTVC] = Ut / Vc;
  k[EK_LVC] = Ut * (log(.5 * Vc / Ut) - .6);             // ekv.py:643
  k[EK_LAMBDA] = m.LAMBDA;
  const double Lc = sqrt(AHK_ESI * m.XJ / m.COX);        // ekv.py:650
  k[EK_LLC] = m.LAMBDA * Lc;
  k[EK_ILCU] = 1.0 / (Lc * m.UCRIT);
  k[EK_NL] = m.N * m.L;
  k[EK_LMIN2] = (m.N * m.L / 10.0) * (m.N * m.L / 10.0);
  k[EK_IUCRIT] = 1.0 / m.UCRIT;
  k[EK_PREF] = m.NP * m.L * m.M;                          // ekv.py:601
}

// ekv.py:703-752 "ismall": the normalised current i solves
//   ln(sqrt(1/4+i) - 1/2) + 2 sqrt(1/4+i) - 1 = v.
// With q = sqrt(1/4+i) - 1/2 (so i = q^2 + q) this is ln q + 2 q = v, and in
// y = ln q the function f(y) = y + 2 e^y - v is increasing and convex.  The
// reference runs scipy's Halley iteration on i from the previous solution until
// |di| <= 1.48e-8; here a closed-form start (within 0.23 of the root for every v)
// is followed by two Halley steps in y (cubic: 0.23 -> 1e-3 -> 2e-12 relative in
// q, measured over v in [-40, 590]); the last exponential is updated by its series.
// Stateless and branch-uniform: every lane does the same work.
AHK_DEV double ekv_q(double v) {
  double y, E;
  // below v = -100 the result (q = e^v < 4e-44) only ever meets the 10 EPS floor of its callers
  // (ekv.py:726): clamping v there changes nothing and bounds every exponential argument below,
  // so the range clamps of fm::exp are not needed (y <= ln(v / 2) bounds them above)
  v = fmax(v, -100.0);
  if (v > 2.0) {
    // y0 = ln(v/2) to single precision (enough for a start: Halley is cubic), and
    // E0 = v/2 taken directly instead of exp(y0): one fast log, no exponential
#ifdef __CUDACC__
    y = (double)__logf((float)(0.5 * v));
#else
    y = (double)logf((float)(0.5 * v));
#endif
    E = 0.5 * v;
  } else {
    y = v < -2.0 ? v : -0.85 + v * (0.555 - 0.065 * v);
    E = fm::exp_nc(y);
  }
  {
    double f = y + 2.0 * E - v, f1 = 1.0 + 2.0 * E;
    double d = fm::div(2.0 * f * f1, 2.0 * f1 * f1 - f * 2.0 * E);
    y -= d;
    E = fm::exp_nc(y);
  }
  {
    double f = y + 2.0 * E - v, f1 = 1.0 + 2.0 * E;
    double d = fm::div(2.0 * f * f1, 2.0 * f1 * f1 - f * 2.0 * E);
    E = E * (1.0 - d * (1.0 - 0.5 * d));
  }
  return E;
}

AHK_DEV double ekv_vp(const double* k, double VG) {   // ekv.py:524-540
  double VGeff = VG - k[EK_VTO] + k[EK_PHI] + k[EK_GAMMA] * k[EK_SQPHI];
  double t = k[EK_T];
  double arg = VG - k[EK_VTO] + t * t;
  double VP = -k[EK_PHI];
  if (VGeff > 0 && arg > 0) {
    VP = VG - k[EK_VTO] - k[EK_GAMMA] * (fm::sqrt(arg) - t);
    if (isnan(VP)) VP = 0;
  }
  return VP;
}

// out[0] = Ids, out[1..3] = gmd, gmg, gms.  k = the device's ekv_setup() block,
// NP = +1 (n) / -1 (p); (vd, vg, vs) are the drive-port voltages w.r.t. bulk.
AHK_DEVFN void ekv_eval(const double* k, int NP, double vd, double vg, double vs, double gmin,
                      double* out) {
  const double Ut = AHK_VTH, iUt = 1.0 / AHK_VTH;
  double VD = vd * NP, VG = vg * NP, VS = vs * NP;  // ekv.py:481-507
  int CS = +1;
  if (VS > VD) { double t = VD; VD = VS; VS = t; CS = -1; }
  const double VP = ekv_vp(k, VG);
  const double nq = 1.0 + .5 * k[EK_GAMMA] * fm::rsqrt(k[EK_PHI] + .5 * VP);
  const double Gs = 2 * nq * Ut * k[EK_KWL];        // ekv.py:509-522
  const double Is = Gs * Ut;
  double qf = ekv_q((VP - VS) * iUt);
  double ifn = qf * qf + qf;
  if (!(ifn > 10 * AHK_EPS)) { ifn = 10 * AHK_EPS; qf = 10 * AHK_EPS; }   // ekv.py:726
  // get_leq_virp, ekv.py:629-670 (Vds = half the drain-source voltage: quirk 11)
  const double Vc = k[EK_VC];
  const double sif = fm::sqrt(ifn);
  const double Vdss = Vc * (fm::sqrt(.25 + k[EK_UTVC] * sif) - .5);
  const double Vdssp = Vc * (fm::sqrt(.25 + k[EK_UTVC] * (sif - .75 * fm::log(ifn))) - .5) + k[EK_LVC];
  const double vser_1 = sif - Vdss * iUt;
  const double Vds = (VD - VS) * .5;
  const double dv2 = 16.0 * Ut * Ut * (k[EK_LAMBDA] * vser_1 + 1.0 / 64);
  const double Vip = fm::sqrt(Vdss * Vdss + dv2) - fm::sqrt((Vds - Vdss) * (Vds - Vdss) + dv2);
  const double delta_l = k[EK_LLC] * fm::log(1 + (Vds - Vip) * k[EK_ILCU]);
  const double Lp = k[EK_NL] - delta_l + (Vds + Vip) * k[EK_IUCRIT];
  const double Leq = .5 * (Lp + fm::sqrt(Lp * Lp + k[EK_LMIN2]));
  const double v_irn = (VP - Vds - VS - fm::sqrt(Vdssp * Vdssp + dv2) +
                        fm::sqrt((Vds - Vdssp) * (Vds - Vdssp) + dv2)) * iUt;
  double qr = ekv_q(v_irn);
  double irn = qr * qr + qr;
  if (!(irn > 10 * AHK_EPS)) { irn = 10 * AHK_EPS; qr = 10 * AHK_EPS; }
  out[0] = fm::div(CS * k[EK_PREF], Leq) * Is * (ifn - irn);       // ekv.py:601-602
  double gmd = CS == 1 ? Gs * qr : Gs * qf;                 // ekv.py:672-692
  double gms = CS == 1 ? -1.0 * Gs * qf : -Gs * qr;
  // get_gmg: nv from the UN-flipped vg (ekv.py:697, quirk 11)
  const double VPu = NP == 1 ? VP : ekv_vp(k, vg);
  const double nv = 1.0 + .5 * k[EK_GAMMA] * fm::rsqrt(k[EK_PHI] + VPu + 1e-12);
  double gmg = fm::div(CS * Gs * (qf - qr), nv);
  if (gmd == 0) gmd = gmin * 2.0;  // ekv.py:331-336
  if (gmg == 0) gmg = gmin * 2.0;
  if (gms == 0) gms = -gmin * 2.0;
  out[1] = gmd; out[2] = gmg; out[3] = gms;
}

// K devices at once, statement by statement ("vertical" form of ekv_eval): the K evaluations
// are independent, so writing each step for all of them back to back hands the instruction
// scheduler K interleaved dependency chains — the transcendental sequences of one evaluation
// (log, sqrt, exp, reciprocals: long serial FP64 chains) hide behind those of the others.
// One lane per instance has 1 warp per scheduler at the BASELINE batch sizes; without this
// the FP64 pipe idles most of the time.  Same arithmetic, same order per device as ekv_eval.
// kk[j] = constant block, NP[j] = +-1, (vd, vg, vs)[j] = port voltages w.r.t. bulk.
template <int K>
AHK_DEVFN void ekv_eval_v(const double* const (&kk)[K], const int (&NP)[K], const double (&vd)[K],
                          const double (&vg)[K], const double (&vs)[K], double gmin,
                          double (&o0)[K], double (&o1)[K], double (&o2)[K], double (&o3)[K]) {
  const double Ut = AHK_VTH, iUt = 1.0 / AHK_VTH;
  double VD[K], VG[K], VS[K], VP[K], VPu[K], nq[K], Gs[K], qf[K], ifn[K], sif[K], Vdss[K], Vdssp[K],
      Vds[K], dv2[K], Vip[K], Leq[K], qr[K], irn[K];
  int CS[K];
#define AHK_V for (int j = 0; j < K; j++)
#pragma unroll
  AHK_V {
    VD[j] = vd[j] * NP[j]; VG[j] = vg[j] * NP[j]; VS[j] = vs[j] * NP[j];
    CS[j] = +1;
    if (VS[j] > VD[j]) { double t = VD[j]; VD[j] = VS[j]; VS[j] = t; CS[j] = -1; }
  }
#pragma unroll
  AHK_V VP[j] = ekv_vp(kk[j], VG[j]);
#pragma unroll
  AHK_V VPu[j] = NP[j] == 1 ? VP[j] : ekv_vp(kk[j], vg[j]);   // get_gmg: nv from the UN-flipped vg
#pragma unroll
  AHK_V nq[j] = 1.0 + .5 * kk[j][EK_GAMMA] * fm::rsqrt(kk[j][EK_PHI] + .5 * VP[j]);
#pragma unroll
  AHK_V Gs[j] = 2 * nq[j] * Ut * kk[j][EK_KWL];
#pragma unroll
  AHK_V qf[j] = ekv_q((VP[j] - VS[j]) * iUt);
#pragma unroll
  AHK_V {
    ifn[j] = qf[j] * qf[j] + qf[j];
    if (!(ifn[j] > 10 * AHK_EPS)) { ifn[j] = 10 * AHK_EPS; qf[j] = 10 * AHK_EPS; }
  }
#pragma unroll
  AHK_V sif[j] = fm::sqrt(ifn[j]);
#pragma unroll
  AHK_V Vdss[j] = kk[j][EK_VC] * (fm::sqrt(.25 + kk[j][EK_UTVC] * sif[j]) - .5);
#pragma unroll
  AHK_V Vdssp[j] = kk[j][EK_VC] * (fm::sqrt(.25 + kk[j][EK_UTVC] * (sif[j] - .75 * fm::log(ifn[j]))) - .5) + kk[j][EK_LVC];
#pragma unroll
  AHK_V {
    const double vser_1 = sif[j] - Vdss[j] * iUt;
    Vds[j] = (VD[j] - VS[j]) * .5;
    dv2[j] = 16.0 * Ut * Ut * (kk[j][EK_LAMBDA] * vser_1 + 1.0 / 64);
  }
#pragma unroll
  AHK_V Vip[j] = fm::sqrt(Vdss[j] * Vdss[j] + dv2[j]) - fm::sqrt((Vds[j] - Vdss[j]) * (Vds[j] - Vdss[j]) + dv2[j]);
#pragma unroll
  AHK_V {
    const double delta_l = kk[j][EK_LLC] * fm::log(1 + (Vds[j] - Vip[j]) * kk[j][EK_ILCU]);
    const double Lp = kk[j][EK_NL] - delta_l + (Vds[j] + Vip[j]) * kk[j][EK_IUCRIT];
    Leq[j] = .5 * (Lp + fm::sqrt(Lp * Lp + kk[j][EK_LMIN2]));
  }
#pragma unroll
  AHK_V {
    const double v_irn = (VP[j] - Vds[j] - VS[j] - fm::sqrt(Vdssp[j] * Vdssp[j] + dv2[j]) +
                          fm::sqrt((Vds[j] - Vdssp[j]) * (Vds[j] - Vdssp[j]) + dv2[j])) * iUt;
    qr[j] = ekv_q(v_irn);
  }
#pragma unroll
  AHK_V {
    irn[j] = qr[j] * qr[j] + qr[j];
    if (!(irn[j] > 10 * AHK_EPS)) { irn[j] = 10 * AHK_EPS; qr[j] = 10 * AHK_EPS; }
  }
#pragma unroll
  AHK_V {
    const double Is = Gs[j] * Ut;
    o0[j] = fm::div(CS[j] * kk[j][EK_PREF], Leq[j]) * Is * (ifn[j] - irn[j]);
    double gmd = CS[j] == 1 ? Gs[j] * qr[j] : Gs[j] * qf[j];
    double gms = CS[j] == 1 ? -1.0 * Gs[j] * qf[j] : -Gs[j] * qr[j];
    const double nv = 1.0 + .5 * kk[j][EK_GAMMA] * fm::rsqrt(kk[j][EK_PHI] + VPu[j] + 1e-12);
    double gmg = fm::div(CS[j] * Gs[j] * (qf[j] - qr[j]), nv);
    if (gmd == 0) gmd = gmin * 2.0;
    if (gmg == 0) gmg = gmin * 2.0;
    if (gms == 0) gms = -gmin * 2.0;
    o1[j] = gmd; o2[j] = gmg; o3[j] = gms;
  }
#undef AHK_V
}

// ---- square-law MOS ----------------------------------------------------------
struct MosqP { double VTO, KP, GAMMA, PHI, LAMBDA, AVT, AKP, W, L, M, N, ZVT, ZKP; int NP; };

// out[0] = CS*Ids; out[1..4] = stamp row "d" (cols d,g,s,b); out[5..8] = row "s".
// Swapped case keeps only [d][s] and [s][d] (element-wise T1*stamp*T2, mosq.py:380-381).
AHK_DEVFN void mosq_eval(const MosqP& m, double vds, double vgs, double vbs, double iea,
                       double gmin, double* out) {
  vds *= m.NP; vgs *= m.NP; vbs *= m.NP;  // mosq.py:552-582
  int CS = +1;
  if (vds < 0) { vgs = vgs - vds; vbs = vbs - vds; vds = -vds; CS = -1; }
  double svt = 0, skp = 0;
  if (m.ZVT != 0 || m.ZKP != 0) {  // mosq.py:584-594
    double r = sqrt(2 * m.W * m.L);
    svt = m.ZVT * m.AVT / r;
    skp = m.ZKP * m.AKP / r;
  }
  double vsqrt1 = fmax(-vbs + 2 * m.PHI, 0.), vsqrt2 = fmax(2 * m.PHI, 0.);
  double s1 = sqrt(vsqrt1), s2 = sqrt(vsqrt2);
  double VT = m.VTO + svt + m.GAMMA * (s1 - s2);
  double KWL = m.KP * m.W / m.L;
  double ids, gmd, gm, gmb = 0;
  if (vgs < VT) {  // mosq.py:642-643
    ids = iea * (vgs / VT + vds / VT) / 100;
    gmd = iea / VT / 100;
    gm = iea / VT / 100;
  } else {
    double ov = vgs - VT;
    if (vds < ov - 0.5 * m.LAMBDA * ov * ov) {  // mosq.py:647-654
      ids = (skp + 1) * KWL * (ov * vds - .5 * vds * vds);
      gmd = KWL * (vgs - vds - VT);
    } else {
      ids = (skp + 1) * .5 * KWL * ov * ov *
            (1 + m.LAMBDA * (vds - vgs + VT + 0.25 * m.LAMBDA * ov * ov));
      gmd = 0.5 * m.KP * m.LAMBDA * m.W / m.L * ov * ov;
    }
    if (vds < ov) {  // gm / gmb region test has no LAMBDA term (mosq.py:700,789)
      gm = KWL * vds;
      if (vsqrt1 > 0) gmb = m.KP * m.GAMMA * vds * m.W / (2 * m.L * s1);
    } else {
      double a = -m.GAMMA * (-s2 + s1) + vgs - m.VTO;
      gm = -0.5 * m.KP * m.LAMBDA * m.W / m.L * a * a +
           0.5 * m.KP * m.W / m.L *
               (m.LAMBDA * (m.GAMMA * (-s2 + s1) + vds - vgs + m.VTO) + 1.0) *
               (-2 * m.GAMMA * (-s2 + s1) + 2 * vgs - 2 * m.VTO);
      if (vsqrt1 > 0) {
        gmb += -0.25 * m.KP * m.GAMMA * m.LAMBDA * m.W * a * a / (m.L * s1);
        gmb += +0.5 * m.KP * m.GAMMA * m.W *
               (m.LAMBDA * (m.GAMMA * (s2 + s1) + vds - vgs + m.VTO) + 1.0) *
               (-m.GAMMA * (s2 + s1) + vgs - m.VTO) / (m.L * s1);
      }
    }
  }
  double MN = m.M / m.N;
  double Ids = m.NP * MN * ids;
  gmd = (1 + skp) * gmd * MN;
  gm = (1 + skp) * gm * MN;
  gmb = m.NP * (1 + skp) * gmb * MN;
  if (gmd == 0) gmd = gmin * 2;  // mosq.py:370-375
  if (gm == 0) gm = gmin * 2;
  if (gmb == 0) gmb = -2 * gmin;
  out[0] = CS * Ids;
  if (CS == 1) {
    out[1] = gmd; out[2] = gm; out[3] = -gmd - gmb - gm; out[4] = gmb;
    out[5] = -gmd; out[6] = -gm; out[7] = gmd + gm + gmb; out[8] = -gmb;
  } else {
    out[1] = 0; out[2] = 0; out[3] = -gmd - gmb - gm; out[4] = 0;
    out[5] = -gmd; out[6] = 0; out[7] = 0; out[8] = 0;
  }
}

// ---- voltage-controlled switch -------------------------------------------------
struct SwP { double VT, VH, RON, ROFF; };

// out[0] = i, out[1] = go, out[2] = gm; *is_on is the hysteresis state, updated on
// every evaluation like the reference's _update_status (switch.py:342-360; the
// three calls per iteration see the same vin, so one update is equivalent).
AHK_DEVFN void switch_eval(const SwP& m, double vout, double vin, double gmin,
                         double* is_on, double* out) {
  double A = (m.RON - m.ROFF) / 2, B = (m.RON + m.ROFF) / 2.;
  for (int rep = 0; rep < 2; rep++) {  // 2nd pass: status after a flip (idempotent)
    int on = *is_on != 0;
    double V = m.VT + m.VH * 2 * (!on) - m.VH;
    double V2 = m.VT + m.VH * 2 * (on) - m.VH;
    double R1 = A * tanh((vin - V) * 1e2) + B;
    double R2 = A * tanh((vin - V2) * 1e2) + B;
    if (vin > V && !on && R1 - R2 == 0.0) { on = 1; V = m.VT + m.VH * 2 * (!on) - m.VH; }
    if (vin < V && on && R1 - R2 == 0.0) { on = 0; }
    *is_on = on;
  }
  int on = *is_on != 0;
  double V = m.VT + m.VH * 2 * (!on) - m.VH;
  double R = A * tanh((vin - V) * 1e2) + B;
  out[0] = vout / R;
  out[1] = 1. / R;
  double th = tanh(1e2 * (V - vin));
  double den = A * th - B;
  out[2] = A * 1e2 * (th * th - 1) / (den * den) + gmin;  // switch.py:393-399
}

// ---- source waveforms -------------------------------------------------------------
AHK_DEVFN double tf_eval(int kind, const double* q, double time) {
  const double PI = 3.141592653589793;
  switch (kind) {
    case 1: {  // pulse
      double v1 = q[0], v2 = q[1], td = q[2], tr = q[3], pw = q[4], tf = q[5], per = q[6];
      time = time - per * (double)(long long)(time / per);
      if (time < td) return v1;
      if (time < td + tr) return v1 + ((v2 - v1) / tr) * (time - td);
      if (time < td + tr + pw) return v2;
      if (time < td + tr + pw + tf) return v2 + ((v1 - v2) / tf) * (time - (td + tr + pw));
      return v1;
    }
    case 2: {  // sin
      double vo = q[0], va = q[1], fr = q[2], td = q[3], th = q[4], phi = q[5];
      if (time < td) return vo + va * sin(PI * phi / 180.);
      return vo + va * exp((td - time) * th) * sin(2 * PI * fr * (time - td) + PI * phi / 180.);
    }
    case 3: {  // exp
      double v1 = q[0], v2 = q[1], td1 = q[2], tau1 = q[3], td2 = q[4], tau2 = q[5];
      if (time < td1) return v1;
      if (time < td2) return v1 + (v2 - v1) * (1 - exp(-1 * (time - td1) / tau1));
      return v1 + (v2 - v1) * (1 - exp(-1 * (time - td1) / tau1)) +
             (v1 - v2) * (1 - exp(-1 * (time - td2) / tau2));
    }
    case 4: {  // sffm
      double vo = q[0], va = q[1], fc = q[2], mdi = q[3], fs = q[4], td = q[5];
      if (time <= td) return vo;
      return vo + va * sin(2 * PI * fc * (time - td) + mdi * sin(2 * PI * fs * (time - td)));
    }
    case 5: {  // am
      double sa = q[0], fc = q[1], fm = q[2], oc = q[3], td = q[4];
      if (time <= td) return 0.;
      return sa * (oc + sin(2 * PI * fm * (time - td))) * sin(2 * PI * fc * (time - td));
    }
    case 6: {  // pwl: q = npts td repeat repeat_time x0 y0 x1 y1 ... (k = 1 spline, extrapolating)
      const int np_ = (int)q[0];
      const double td = q[1], rpt = q[3];
      const double* xy = q + 4;
      const double xmax = xy[2 * (np_ - 1)];
      if (time <= td) time = 0;
      else {
        time = time - td;
        if (q[2] != 0.0 && time > xmax) time = fmod(time - xmax, xmax - rpt) + rpt;
      }
      int i = 0;   // segment [i, i+1] that holds (or, outside the table, is nearest to) time
      while (i < np_ - 2 && time > xy[2 * (i + 1)]) i++;
      const double x0 = xy[2 * i], y0 = xy[2 * i + 1], x1 = xy[2 * i + 2], y1 = xy[2 * i + 3];
      return y0 + (y1 - y0) / (x1 - x0) * (time - x0);
    }
  }
  return 0;
}

}  // namespace ahk
