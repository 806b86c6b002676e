// thermo2.cuh — packed (two-lane) dry thermodynamics + hydrostatic reference state for the k5_* kernels.
// Same formulas and operation order as `thermo` in common.cuh (precomputed_quantities.jl:733-815 dry branch;
// refstate_thermodynamics.jl:22-168), evaluated on f32x2 pairs: FFMA2/FMUL2 algebra, packed log/exp (pair.cuh), MUFU-based
// reciprocals.  Float64 instantiates the same code on the two-member struct with log/exp/division from libm.
#pragma once
#include "common.cuh"
#include "pair.cuh"

namespace b200 {

template <class FT> __device__ __forceinline__ P2<FT> rcpn2(P2<FT> a) { return P2<FT>(rcpn_(a.lo()), rcpn_(a.hi())); }
__device__ __forceinline__ float mx_(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double mx_(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float mn_(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double mn_(double a, double b) { return fmin(a, b); }
template <class FT> __device__ __forceinline__ P2<FT> max2(FT s, P2<FT> a) { return P2<FT>(mx_(s, a.lo()), mx_(s, a.hi())); }

template <class FT>
struct Pt2 {
  P2<FT> T, p, h, Pi, thp /*θ_v-θ_vr*/, thv, phir, sdr, lnPi;
};
template <class FT>
__device__ __forceinline__ Pt2<FT> thermo2(const Par<FT>& P, P2<FT> rho, P2<FT> rhoe, P2<FT> K, FT Phi) {
  using V = P2<FT>;
  Pt2<FT> o;
  const V etot = rhoe * rcpn2(rho);
  const V eint = (etot - K) - Phi;
  // e_int = cv_d (T − T_0) − R_d T_0  (docs/src/thermodynamics.md:103-111)
  o.T = max2(P.T_min_sgs, fma2(eint + P.RT0, V(P.icv), V(P.T_0)));
  o.h = fma2(o.T, V(P.R_d), etot);
  o.p = (rho * P.R_d) * o.T;
  o.lnPi = logp(o.p * P.ip0) * P.kappa;
  o.Pi = expp(o.lnPi);
  const V rPi = rcpn2(o.Pi);
  const V x2 = o.Pi * o.Pi, x4 = x2 * x2;
  const V Pi7 = (x4 * x2) * o.Pi;
  const V Tr = fma2(Pi7, V(P.Ts_ref - P.Tmin_ref), V(P.Tmin_ref));
  o.thv = o.T * rPi;
  o.thp = (o.T - Tr) * rPi;
  o.phir = fma2(o.lnPi, V(P.Tmin_ref), (Pi7 - FT(1)) * P.dTs7) * (-P.cp_d);
  o.sdr = fma2(Tr - P.T_0, V(P.cp_d), o.phir);
  return o;
}

// two scalar thermodynamic states → one packed state
template <class FT>
__device__ __forceinline__ Pt2<FT> pack_pt(const Pt<FT>& a, const Pt<FT>& b) {
  Pt2<FT> o;
  o.T = P2<FT>(a.T, b.T); o.p = P2<FT>(a.p, b.p); o.h = P2<FT>(a.h, b.h); o.Pi = P2<FT>(a.Pi, b.Pi); o.thp = P2<FT>(a.thp, b.thp);
  o.thv = P2<FT>(a.thv, b.thv); o.phir = P2<FT>(a.phir, b.phir); o.sdr = P2<FT>(a.sdr, b.sdr); o.lnPi = P2<FT>(a.lnPi, b.lnPi);
  return o;
}
// Packed form of pgf_aux / pgf_diff (common.cuh): Float32 evaluates ΔΠ and ΔΦ_r between adjacent levels in difference form from
// κ·log(p_hi/p_lo); Float64 keeps the literal differences.  expm1 by its Taylor series (|Δ| ≲ 0.35 for any grid with Δz ≤ 8 km:
// the x¹⁰/10! remainder is < 10⁻¹¹).
template <class FT> __device__ __forceinline__ P2<FT> pgf_aux2(const Pt2<FT>& t);
template <> __device__ __forceinline__ P2<double> pgf_aux2<double>(const Pt2<double>& t) { return t.phir; }
template <> __device__ __forceinline__ P2<float> pgf_aux2<float>(const Pt2<float>& t) { return t.p; }
__device__ __forceinline__ void pgf_diff2(const Par<double>&, P2<double> Pilo, P2<double> Pihi, P2<double> qlo, P2<double> qhi,
                                          P2<double>& dPi, P2<double>& dphr) {
  dPi = Pihi - Pilo; dphr = qhi - qlo;
}
__device__ __forceinline__ void pgf_diff2(const Par<float>& P, P2<float> Pilo, P2<float> /*Pihi*/, P2<float> plo, P2<float> phi,
                                          P2<float>& dPi, P2<float>& dphr) {
  using V = P2<float>;
  const V dl = logp(phi * rcpn2(plo)) * P.kappa;
  V e = fma2(dl, V(1.0f / 362880.0f), V(1.0f / 40320.0f));  // expm1(x) = x(1 + x/2 + x²/6 + … + x⁸/9!)
  e = fma2(e, dl, V(1.0f / 5040.0f));
  e = fma2(e, dl, V(1.0f / 720.0f));
  e = fma2(e, dl, V(1.0f / 120.0f));
  e = fma2(e, dl, V(1.0f / 24.0f));
  e = fma2(e, dl, V(1.0f / 6.0f));
  e = fma2(e, dl, V(0.5f));
  e = fma2(e, dl, V(1.0f));
  e = e * dl;
  dPi = Pilo * e;
  V q = e + 7.0f;  // (1 + e)⁷ − 1 = e(7 + 21e + 35e² + 35e³ + 21e⁴ + 7e⁵ + e⁶)
  q = fma2(q, e, V(21.0f));
  q = fma2(q, e, V(35.0f));
  q = fma2(q, e, V(35.0f));
  q = fma2(q, e, V(21.0f));
  q = fma2(q, e, V(7.0f));
  const V x2 = Pilo * Pilo, x4 = x2 * x2;
  const V d7 = ((x4 * x2) * Pilo) * (q * e);
  dphr = fma2(dl, V(P.Tmin_ref), d7 * P.dTs7) * (-P.cp_d);
}

}  // namespace b200
