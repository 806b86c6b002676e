// moist.cuh — moist thermodynamics of the EquilibriumMicrophysics0M configuration (ρq_tot = component 4 of Y.c, thermodynamically
// active; condensate diagnosed by saturation adjustment).  Reference call sites: set_implicit_precomputed_quantities!
// (src/cache/precomputed_quantities.jl:735-815), theta_v / q_tot_r (src/utils/refstate_thermodynamics.jl:56-61,140-150),
// ᶜkappa_m_field! / ᶜ∂p∂ρq_tot_field! (manual_sparse_jacobian.jl:653-690), ᶜh_eff_plus_Φ! (eddy_diffusion_closures.jl:970-983).
// Thermodynamics.jl 1.3.0 is not vendored: the formulation is the published one of docs/src/thermodynamics.md:60-150 (calorically
// perfect constituents, Romps 2008 energy references, Rankine–Kirchhoff saturation vapour pressure with the Pressel 2015
// liquid-fraction-weighted latent heat, Kaul 2015 supercooled-liquid ramp), restated literally in oracle/dycore_oracle.py.
// Scalar code (one point at a time): the packed kernels call it per lane — the saturated branch is rare and data-dependent.
#pragma once
#include "common.cuh"

namespace b200 {

template <class FT> __device__ __forceinline__ FT eps_();
template <> __device__ __forceinline__ float eps_<float>() { return 1.1920929e-07f; }
template <> __device__ __forceinline__ double eps_<double>() { return 2.220446049250313e-16; }

// thermodynamic state at one point
template <class FT>
struct Mst {
  FT T, qt /* q_tot_nonneg */, ql, qi, Rm, cvm;
};

template <class FT> __device__ __forceinline__ FT gas_constant_air(const Par<FT>& P, FT qt, FT ql, FT qi) {
  // = R_d (1 − q_t) + R_v q_v, summed around the exact constant R_d (small terms first: one final rounding in Float32)
  return P.R_d + ((P.M.R_v - P.R_d) * qt - P.M.R_v * (ql + qi));
}
template <class FT> __device__ __forceinline__ FT cv_m(const Par<FT>& P, FT qt, FT ql, FT qi) {
  return P.cv_d + (P.M.cv_v - P.cv_d) * qt + (P.M.cp_l - P.M.cv_v) * ql + (P.M.cp_i - P.M.cv_v) * qi;
}
template <class FT> __device__ __forceinline__ FT internal_energy(const Par<FT>& P, FT T, FT qt, FT ql, FT qi) {
  return cv_m(P, qt, ql, qi) * (T - P.T_0) + (qt - ql - qi) * P.M.e_v0 - qi * P.M.e_i0 - (FT(1) - qt) * P.RT0;
}
// supercooled-liquid ramp λ(T) and dλ/dT
template <class FT> __device__ __forceinline__ FT liquid_fraction(const Par<FT>& P, FT T, FT& dlam) {
  dlam = FT(0);
  if (T >= P.M.T_frz) return FT(1);
  if (T <= P.M.T_icn) return FT(0);
  const FT w = P.M.T_frz - P.M.T_icn, x = (T - P.M.T_icn) / w, n = P.M.pow_icn;
  if (n == FT(1)) { dlam = FT(1) / w; return x; }
  dlam = n * pow_(x, n - FT(1)) / w;
  return pow_(x, n);
}
// ln p_vs(T) with the λ-weighted latent heat; also returns Δcp(λ) and L_0(λ)
template <class FT> __device__ __forceinline__ FT ln_pvs(const Par<FT>& P, FT T, FT lam, FT& dcp, FT& L0) {
  dcp = lam * (P.M.cp_v - P.M.cp_l) + (FT(1) - lam) * (P.M.cp_v - P.M.cp_i);
  L0 = lam * P.M.LH_v0 + (FT(1) - lam) * P.M.LH_s0;
  return P.M.ln_ptr + dcp / P.M.R_v * log_(T / P.M.T_tr) + (L0 - dcp * P.T_0) / P.M.R_v * (FT(1) / P.M.T_tr - FT(1) / T);
}
template <class FT> __device__ __forceinline__ FT q_vap_saturation(const Par<FT>& P, FT T, FT rho, FT lam) {
  FT dcp, L0;
  return exp_(ln_pvs(P, T, lam, dcp, L0)) / (rho * P.M.R_v * T);
}

// TD.saturation_adjustment(thermo_params, TD.ρe(), ρ, e_int, q_tot): Newton's method with the analytic derivative from the all-vapour
// temperature, iterated to round-off; an iterate beyond the dew point is pulled back half-way to the last iterate with a negative
// residual (strongly supersaturated input only).  Operation order as in Oracle.saturation_adjustment.
template <class FT>
__device__ __forceinline__ Mst<FT> saturation_adjustment(const Par<FT>& P, FT rho, FT e_int, FT qt) {
  Mst<FT> o;
  o.qt = qt; o.ql = o.qi = FT(0);
  const FT cvu = cv_m(P, qt, FT(0), FT(0));
  FT T = P.T_0 + (e_int - qt * P.M.e_v0 + (FT(1) - qt) * P.RT0) / cvu;
  FT dlam, lam = liquid_fraction(P, T, dlam);
  if (qt > q_vap_saturation(P, T, rho, lam)) {
    FT Tlo = T;
    const FT tol = FT(8) * eps_<FT>();
    const FT dcpl = P.M.cp_i - P.M.cp_l;
    for (int it = 0; it < 40; ++it) {
      lam = liquid_fraction(P, T, dlam);
      FT dcp, L0;
      const FT lnp = ln_pvs(P, T, lam, dcp, L0);
      const FT qvs = exp_(lnp) / (rho * P.M.R_v * T);
      const FT dlnp = (L0 + dcp * (T - P.T_0)) / (P.M.R_v * T * T) +
                      dlam * (dcpl / P.M.R_v * log_(T / P.M.T_tr) + (-P.M.e_i0 - dcpl * P.T_0) / P.M.R_v * (FT(1) / P.M.T_tr - FT(1) / T));
      const bool on = qt > qvs;
      const FT qc = on ? qt - qvs : FT(0);
      const FT dqc = -qvs * (dlnp - FT(1) / T);
      const FT ql = lam * qc, qi = (FT(1) - lam) * qc;
      const FT dql = dlam * qc + lam * dqc, dqi = -dlam * qc + (FT(1) - lam) * dqc;
      const FT f = internal_energy(P, T, qt, ql, qi) - e_int;
      const FT df = cv_m(P, qt, ql, qi) + (T - P.T_0) * ((P.M.cp_l - P.M.cv_v) * dql + (P.M.cp_i - P.M.cv_v) * dqi) - dqc * P.M.e_v0 - dqi * P.M.e_i0;
      if (on && f < FT(0)) Tlo = T;
      const FT Tn = on ? T - f / df : FT(0.5) * (Tlo + T);
      const bool done = on && abs_(Tn - T) <= tol * T;
      T = Tn;
      if (done) break;
    }
    lam = liquid_fraction(P, T, dlam);
    const FT qc = fmax_(FT(0), qt - q_vap_saturation(P, T, rho, lam));
    o.ql = lam * qc; o.qi = (FT(1) - lam) * qc;
  }
  o.T = T;
  o.Rm = gas_constant_air(P, qt, o.ql, o.qi);
  o.cvm = cv_m(P, qt, o.ql, o.qi);
  return o;
}

// ᶜh_eff_plus_Φ!: aggregate specific enthalpy of the suspended water (vapour + cloud liquid + cloud ice) + Φ
template <class FT> __device__ __forceinline__ FT h_eff_plus_phi(const Par<FT>& P, const Mst<FT>& m, FT Phi) {
  const FT qv = fmax_(FT(0), m.qt - m.ql - m.qi), ql = fmax_(FT(0), m.ql), qi = fmax_(FT(0), m.qi);
  const FT dT = m.T - P.T_0;
  const FT num = (P.M.cp_v * dT + P.M.LH_v0) * qv + (P.M.cp_l * dT) * ql + (P.M.cp_i * dT - P.M.e_i0) * qi;
  return num / fmax_(qv + ql + qi, eps_<FT>()) + Phi;
}
// q_tot_r(p) = RH_ref·q_sat(T_r(p), ρ_r(p)) over liquid, zero above 250 hPa (refstate_thermodynamics.jl:140-150); Tr = T_r(p)
template <class FT> __device__ __forceinline__ FT q_tot_r(const Par<FT>& P, FT p, FT Tr) {
  if (p < FT(25000)) return FT(0);
  return FT(0.5) * q_vap_saturation(P, Tr, p / (P.R_d * Tr), FT(1));
}

// Moist counterpart of `thermo` (common.cuh): the thermodynamic state of one point from (ρ, ρe_tot, ρq_tot, K, Φ) and the hydrostatic
// reference-state quantities; no T_min_sgs floor in this branch (precomputed_quantities.jl:735-747).
template <class FT>
__device__ __forceinline__ Pt<FT> thermo_m(const Par<FT>& P, FT rho, FT rhoe, FT rhoq, FT K, FT Phi, Mst<FT>& m) {
  Pt<FT> o;
  const FT etot = rhoe / rho;
  const FT eint = etot - K - Phi;
  m = saturation_adjustment(P, rho, eint, fmax_(FT(0), rhoq / rho));
  o.T = m.T;
  o.h = etot + m.Rm * o.T;  // TD.total_enthalpy = e_tot + R_m T
  // virtual temperature T_v = T R_m/R_d = T + T·((R_v/R_d − 1) q_t − (R_v/R_d) q_c): p = ρ R_d T_v and θ_v = T_v/Π share it
  const FT epsv = P.M.R_v / P.R_d;
  const FT Tv = fma_(o.T, (epsv - FT(1)) * m.qt - epsv * (m.ql + m.qi), o.T);
  o.p = (rho * P.R_d) * Tv;
  const FT lnPi = P.kappa * log_(o.p * P.ip0);  // TD.exner_given_pressure: the dry exponent R_d/cp_d
  o.Pi = exp_(lnPi);
  o.lnPi = lnPi;
  const FT Pi7 = pow7(o.Pi);
  const FT Tr = P.Tmin_ref + (P.Ts_ref - P.Tmin_ref) * Pi7;
  o.thv = Tv / o.Pi;
  o.thp = (Tv - Tr) / o.Pi;
  o.phir = -P.cp_d * (P.Tmin_ref * lnPi + P.dTs7 * (Pi7 - FT(1)));
  o.sdr = P.cp_d * (Tr - P.T_0) + o.phir;
  return o;
}
// ∂p/∂ρq_tot at constant ρ, ρe_tot (manual_sparse_jacobian.jl:670-690) and ∂p/∂ρ (:816-818) with κ_m = R_m/cv_m
template <class FT> __device__ __forceinline__ FT dp_drhoq(const Par<FT>& P, const Mst<FT>& m) {
  const FT kap = m.Rm / m.cvm;
  return kap * (-P.M.e_v0 - P.RT0 - (P.M.cv_v - P.cv_d) * (m.T - P.T_0)) + (P.M.R_v - P.R_d) * m.T;
}

}  // namespace b200
