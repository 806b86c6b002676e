// moist2.cuh — packed (two-lane) moist thermodynamic state for the k5_* kernels (microphysics_model 0M).
// The common path — unsaturated air — is evaluated on f32x2 pairs like thermo2 (FFMA2 algebra, packed log/exp, MUFU reciprocals):
// the all-vapour temperature, q_vs(T, ρ) for the saturation test, then p, Π, θ_v and the reference-state quantities.  Pairs with a
// saturated lane run the Newton iteration of the saturation adjustment, also packed (saturation_adjustment2; in the baroclinic-wave
// setups every column is ice-saturated near its very cold top, so about half of the warps take it).  Same formulas as thermo_m / the oracle; the
// Rankine–Kirchhoff exponents A = Δcp/R_v and B = (L_0 − Δcp T_0)/R_v are linear in the liquid fraction and come precomputed
// (MPar: A_liq, A_ice, B_liq, B_ice).  Float64 instantiates the same code on the two-member struct (libm log/exp, IEEE division).
#pragma once
#include "moist.cuh"
#include "thermo2.cuh"

namespace b200 {

template <class FT>
struct Mst2 {
  P2<FT> T, qt /* q_tot_nonneg */, ql, qi, Rm, cvm;
};

template <class FT> __device__ __forceinline__ P2<FT> div2(P2<FT> a, P2<FT> b) { return a * rcpn2(b); }

// ln p_vs(T) on pairs for a given liquid fraction pair
template <class FT>
__device__ __forceinline__ P2<FT> ln_pvs2(const Par<FT>& P, P2<FT> T, P2<FT> lam) {
  using V = P2<FT>;
  const V A = fma2(lam, V(P.M.A_liq - P.M.A_ice), V(P.M.A_ice)), B = fma2(lam, V(P.M.B_liq - P.M.B_ice), V(P.M.B_ice));
  return fma2(B, V(P.M.iT_tr) - rcpn2(T), fma2(A, logp(T * P.M.iT_tr), V(P.M.ln_ptr)));
}
template <class FT>
__device__ __forceinline__ P2<FT> liquid_fraction2(const Par<FT>& P, P2<FT> T) {
  FT d0, d1;
  return P2<FT>(liquid_fraction(P, T.lo(), d0), liquid_fraction(P, T.hi(), d1));
}

// Saturation adjustment on a pair: the Newton iteration of moist.cuh (same residual, derivative, safeguard and stopping rule) in packed
// arithmetic, both lanes at once; lanes that are unsaturated (sat = false) or finished keep their value.  d/dT ln p_vs at fixed λ is
// A/T + B/T² with the precomputed exponents.  Returns T; ql, qi are the equilibrium partition at the returned T.
template <class FT>
__device__ __forceinline__ P2<FT> saturation_adjustment2(const Par<FT>& P, P2<FT> rho, P2<FT> eint, P2<FT> qt, P2<FT> cvu, P2<FT> T1, bool s0, bool s1,
                                                         P2<FT>& ql, P2<FT>& qi) {
  using V = P2<FT>;
  V T = T1, Tlo = T1;
  const FT tol = FT(8) * eps_<FT>();
  const V irvr = rcpn2(rho * P.M.R_v);
  const V base = (qt * P.M.e_v0 - (V(FT(1)) - qt) * P.RT0) - eint;  // the T-independent part of the residual
  bool a0 = s0, a1 = s1;
  for (int it = 0; it < 40 && (a0 || a1); ++it) {
    FT d0, d1;
    const V lam(liquid_fraction(P, T.lo(), d0), liquid_fraction(P, T.hi(), d1)), dlam(d0, d1);
    const V A = fma2(lam, V(P.M.A_liq - P.M.A_ice), V(P.M.A_ice)), B = fma2(lam, V(P.M.B_liq - P.M.B_ice), V(P.M.B_ice));
    const V rT = rcpn2(T), lt = logp(T * P.M.iT_tr), dr = V(P.M.iT_tr) - rT;
    const V qvs = expp(fma2(B, dr, fma2(A, lt, V(P.M.ln_ptr)))) * (irvr * rT);
    const V dlnp = fma2(dlam, fma2(dr, V(P.M.B_liq - P.M.B_ice), lt * (P.M.A_liq - P.M.A_ice)), fma2(A, T, B) * (rT * rT));
    const bool on0 = qt.lo() > qvs.lo(), on1 = qt.hi() > qvs.hi();
    const V d = qt - qvs, qc(on0 ? d.lo() : FT(0), on1 ? d.hi() : FT(0));
    const V dqc = -(qvs * (dlnp - rT));
    const V l = lam * qc, i = qc - l;
    const V dl = fma2(lam, dqc, dlam * qc), di = dqc - dl;
    const V cvm = fma2(i, V(P.M.cp_i - P.M.cv_v), fma2(l, V(P.M.cp_l - P.M.cv_v), cvu));
    const V dT = T - P.T_0;
    const V f = fma2(cvm, dT, base) - fma2(i, V(P.M.e_i0), qc * P.M.e_v0);
    const V df = fma2(dT, fma2(di, V(P.M.cp_i - P.M.cv_v), dl * (P.M.cp_l - P.M.cv_v)), cvm) - fma2(di, V(P.M.e_i0), dqc * P.M.e_v0);
    const V nw = T - f * rcpn2(df), bs = (Tlo + T) * FT(0.5);
    FT t0 = T.lo(), t1 = T.hi(), lo0 = Tlo.lo(), lo1 = Tlo.hi();
    if (a0) {
      if (on0 && f.lo() < FT(0)) lo0 = t0;
      const FT tn = on0 ? nw.lo() : bs.lo();
      a0 = !(on0 && abs_(tn - t0) <= tol * t0);
      t0 = tn;
    }
    if (a1) {
      if (on1 && f.hi() < FT(0)) lo1 = t1;
      const FT tn = on1 ? nw.hi() : bs.hi();
      a1 = !(on1 && abs_(tn - t1) <= tol * t1);
      t1 = tn;
    }
    T = V(t0, t1); Tlo = V(lo0, lo1);
  }
  // equilibrium partition at the converged temperature (saturated lanes only)
  FT d0, d1;
  const V lam(liquid_fraction(P, T.lo(), d0), liquid_fraction(P, T.hi(), d1));
  const V qvs = expp(ln_pvs2(P, T, lam)) * (irvr * rcpn2(T));
  const V d = qt - qvs, qc(s0 ? fmax_(FT(0), d.lo()) : FT(0), s1 ? fmax_(FT(0), d.hi()) : FT(0));
  ql = lam * qc; qi = qc - ql;
  return T;
}

// thermodynamic state of a pair of points from (ρ, ρe_tot, ρq_tot, K, Φ): Pt2 as thermo2, the moist extras in m
template <class FT>
__device__ __forceinline__ Pt2<FT> thermo2m(const Par<FT>& P, P2<FT> rho, P2<FT> rhoe, P2<FT> rhoq, P2<FT> K, FT Phi, Mst2<FT>& m) {
  using V = P2<FT>;
  Pt2<FT> o;
  const V irho = rcpn2(rho);
  const V etot = rhoe * irho;
  const V eint = (etot - K) - Phi;
  const V qt = max2(FT(0), rhoq * irho);
  // all-vapour temperature T_1 = T_0 + (e_int − q_t e_v0 + (1 − q_t) R_d T_0)/cv_m(q_t)
  const V cvu = fma2(qt, V(P.M.cv_v - P.cv_d), V(P.cv_d));
  const V num = (eint - qt * P.M.e_v0) + (V(FT(1)) - qt) * P.RT0;
  V T = fma2(num, rcpn2(cvu), V(P.T_0));
  V ql(FT(0)), qi(FT(0));
  {  // saturated?  q_t > q_vs(T_1, ρ) with the liquid fraction of T_1
    const V lam1 = liquid_fraction2(P, T);
    const V qvs = expp(ln_pvs2(P, T, lam1)) * rcpn2((rho * P.M.R_v) * T);
    const bool s0 = qt.lo() > qvs.lo(), s1 = qt.hi() > qvs.hi();
    if (s0 || s1) {
      // Newton only where the latent heating can move T in this precision: q_t L_s/cv_d ≥ ¼ ulp(T).  Below that (the ice-saturated,
      // very cold and dry top of the baroclinic-wave columns: q_t = 1e-12) T_1 IS the adjusted temperature and the condensate is
      // the excess at T_1.  In Float64 the threshold is q_t ≈ 1e-20: never taken.
      const bool g0 = s0 && qt.lo() >= P.M.q_neg * T.lo(), g1 = s1 && qt.hi() >= P.M.q_neg * T.hi();
      const V d = qt - qvs, qc1((s0 && !g0) ? d.lo() : FT(0), (s1 && !g1) ? d.hi() : FT(0));
      V l(FT(0)), i(FT(0));
      if (g0 || g1) T = saturation_adjustment2(P, rho, eint, qt, cvu, T, g0, g1, l, i);
      ql = fma2(lam1, qc1, l); qi = i + (qc1 - lam1 * qc1);
    }
  }
  const V qc = ql + qi;
  m.T = T; m.qt = qt; m.ql = ql; m.qi = qi;
  m.Rm = V(P.R_d) + (qt * (P.M.R_v - P.R_d) - qc * P.M.R_v);
  m.cvm = fma2(qi, V(P.M.cp_i - P.M.cv_v), fma2(ql, V(P.M.cp_l - P.M.cv_v), cvu));
  o.T = T;
  o.h = fma2(m.Rm, T, etot);  // TD.total_enthalpy = e_tot + R_m T
  const V Tv = fma2(T, qt * (P.M.epsv - FT(1)) - qc * P.M.epsv, T);  // T R_m/R_d
  o.p = (rho * P.R_d) * Tv;
  o.lnPi = logp(o.p * P.ip0) * P.kappa;
  o.Pi = expp(o.lnPi);
  const V rPi = rcpn2(o.Pi);
  const V x2 = o.Pi * o.Pi, x4 = x2 * x2;
  const V Pi7 = (x4 * x2) * o.Pi;
  const V Tr = fma2(Pi7, V(P.Ts_ref - P.Tmin_ref), V(P.Tmin_ref));
  o.thv = Tv * rPi;
  o.thp = (Tv - Tr) * rPi;
  o.phir = fma2(o.lnPi, V(P.Tmin_ref), (Pi7 - FT(1)) * P.dTs7) * (-P.cp_d);
  o.sdr = fma2(Tr - P.T_0, V(P.cp_d), o.phir);
  return o;
}
// T_r(p) of a state returned by thermo2m
template <class FT> __device__ __forceinline__ P2<FT> t_ref2(const Par<FT>& P, const Pt2<FT>& t) {
  const P2<FT> x2 = t.Pi * t.Pi, x4 = x2 * x2;
  return fma2((x4 * x2) * t.Pi, P2<FT>(P.Ts_ref - P.Tmin_ref), P2<FT>(P.Tmin_ref));
}
// q_tot_r(p) = ½ q_sat(T_r, ρ_r = p/(R_d T_r)) over liquid = ½ p_vs,liq(T_r) R_d/(R_v p), zero below 250 hPa
template <class FT> __device__ __forceinline__ P2<FT> q_tot_r2(const Par<FT>& P, const Pt2<FT>& t) {
  using V = P2<FT>;
  const V q = (expp(ln_pvs2(P, t_ref2(P, t), V(FT(1)))) * (FT(0.5) / P.M.epsv)) * rcpn2(t.p);
  return V(t.p.lo() < FT(25000) ? FT(0) : q.lo(), t.p.hi() < FT(25000) ? FT(0) : q.hi());
}
// ᶜh_eff_plus_Φ!
template <class FT> __device__ __forceinline__ P2<FT> h_eff_plus_phi2(const Par<FT>& P, const Mst2<FT>& m, FT Phi) {
  using V = P2<FT>;
  const V qv = max2(FT(0), m.qt - m.ql - m.qi), ql = max2(FT(0), m.ql), qi = max2(FT(0), m.qi);
  const V dT = m.T - P.T_0;
  const V num = fma2(fma2(dT, V(P.M.cp_i), V(-P.M.e_i0)), qi, fma2(dT * P.M.cp_l, ql, fma2(dT, V(P.M.cp_v), V(P.M.LH_v0)) * qv));
  return num * rcpn2(max2(eps_<FT>(), (qv + ql) + qi)) + Phi;
}
// ∂p/∂ρq_tot at constant ρ, ρe_tot with κ_m = R_m/cv_m (also returned)
template <class FT> __device__ __forceinline__ P2<FT> dp_drhoq2(const Par<FT>& P, const Mst2<FT>& m, P2<FT>& kap) {
  using V = P2<FT>;
  kap = m.Rm * rcpn2(m.cvm);
  return fma2(kap, V(-P.M.e_v0 - P.RT0) - (m.T - P.T_0) * (P.M.cv_v - P.cv_d), m.T * (P.M.R_v - P.R_d));
}

}  // namespace b200
