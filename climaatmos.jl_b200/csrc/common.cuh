// common.cuh — shared device-side definitions for the B200 dycore kernels.
//
// Data layout (ClimaCore VIJFH, C order [h][f][j][i][v]): for a fixed element h and component f
// the (j,i,v) slab is one contiguous block of 16*Nv (centres) or 16*(Nv+1) (faces) values with the
// level index fastest.  Every element-kernel below assigns ONE ELEMENT PER CTA: consecutive lanes
// own consecutive levels (fully coalesced global loads of the contiguous slab), the 16 GLL nodes
// are spread over the warps, and all horizontal (4x4 derivative-matrix) and vertical (k±1)
// neighbours are read from a shared-memory copy of the slab with node stride LVP = 65 words
// (odd ⇒ conflict-free both for level-major and for column-sequential access).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

constexpr int NQ = 4;
constexpr int NN = 16;    // nodes per element
constexpr int LV = 64;    // max faces per column (Nv + 1 <= 64)
constexpr int LVP = 65;   // padded shared-memory node stride
constexpr int NT = 256;   // threads per element CTA
constexpr int NIT = (NN * LV) / NT;  // 4 point-iterations per thread
constexpr int SLAB = NN * LVP;       // shared-memory words per field

// horizontal geometry components, stored [h][HG_N][16]
enum {
  HG_J2 = 0, HG_RJ2, HG_GI11, HG_GI12, HG_GI22, HG_GC11, HG_GC12, HG_GC22,
  HG_COR1, HG_COR2, HG_COR3, HG_SIN2, HG_COS2, HG_DSSW, HG_A00, HG_A01, HG_A10, HG_A11,
  HG_AI00, HG_AI01, HG_AI10, HG_AI11, HG_WJ /* quadrature weight × J2 (limiter) */, HG_N
};
constexpr int HG_ELEM = 13;  // components the element kernels stage (J2..COS2)

// per-level constants (host-precomputed in double, stored in FT)
template <class FT>
struct VLev {
  FT sc2i[LV];   // 1 / s_c^2, s = (R+z)/R (deep) or 1
  FT sf2i[LV];   // 1 / s_f^2
  FT sf[LV];     // s_f
  FT dzc[LV];    // vertical J at centres
  FT dzf[LV];    // vertical J at faces
  FT mc[LV];     // s_c^2 * dz_c   (J_c = J2 * mc)
  FT rmc[LV];    // 1 / mc
  FT g33f[LV];   // 1 / dz_f^2
  FT phic[LV];   // grav * z_c
  FT dphif[LV];  // ᶠgradᵥ(Φ): phic[f]-phic[f-1], 0 on boundary faces
  FT brw[LV];    // β_rayleigh_u₃(z_f)
  FT bruh[LV];   // β_rayleigh_uₕ(z_c)
  FT bvc[LV];    // β_viscous(z_c)
  FT bvf[LV];    // β_viscous(z_f)
  FT D[16];      // strong derivative matrix  D[i*4+k]
  FT Dw[16];     // weak derivative matrix   Dw[i*4+k] = -D[k][i] w_k / w_i
};

// moist thermodynamics parameters (EquilibriumMicrophysics0M, moist.cuh); all zero for dry contexts
template <class FT>
struct MPar {
  FT R_v, cv_v, cp_v, cp_l, cp_i, LH_v0, LH_s0, e_v0 /* L_v0 − R_v T_0 */, e_i0 /* L_f0 */, T_tr, ln_ptr /* ln p_triple */, T_frz, T_icn, pow_icn;
  // derived (host, double precision): Rankine–Kirchhoff exponents A = Δcp/R_v, B = (L_0 − Δcp T_0)/R_v over liquid / ice, 1/T_triple, R_v/R_d
  FT A_liq, A_ice, B_liq, B_ice, iT_tr, epsv;
  FT q_neg;  // ¼ ulp · cv_d / L_s0: below q_t = q_neg·T condensation cannot change T in this precision (moist2.cuh)
};

template <class FT>
struct Par {
  FT R_d, cp_d, cv_d, T_0, p0, kappa, Ts_ref, Tmin_ref, T_min_sgs, dt;
  FT icv, ip0, dTs7, RT0;  // 1/cv_d, 1/p0, (Ts_ref − Tmin_ref)/7, R_d·T_0
  FT nu4v, nu4s, ddf;
  int nh, nv;
  int ncf;   // components of Y.c: 4 + number of passive tracers (ρ, uₕ₁, uₕ₂, ρe_tot, ρχ…)
  int tupw;  // tracer_upwinding: 0 none, 1 first_order, 3 vanleer_limiter
  int hyperdiff, rayleigh, viscous, upwinding;
  int hs;  // Held–Suarez forcing
  FT hs_ka, hs_ks, hs_kf, hs_sigb, hs_isig, hs_dTy, hs_Teq, hs_dthz, hs_Tmin, hs_iMSLP, hs_ikap;
  int moist;  // microphysics_model 0M: component 4 of Y.c is the thermodynamically active ρq_tot
  MPar<FT> M;
};

// Programmatic dependent launch (capi.cu: launchx).  pdl_launch lets the next kernel of the stream start its CTAs as soon as every
// CTA of this grid has been scheduled; pdl_wait blocks until all earlier grids have completed and their writes are visible.  A kernel
// only reads context constants (geometry, level tables, DSS records) before pdl_wait.  Both are no-ops in a plain launch.
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// pdl_wait + pointer laundering.  Loads through `const T* __restrict__` kernel parameters are "invariant" for the compiler and may be
// hoisted above any barrier, including griddepcontrol.wait (observed: k5_exp_a read the state before the previous kernel had written
// it).  Passing the pointers that earlier kernels write through an opaque asm AFTER the wait makes every later load depend on it.
template <class P> __device__ __forceinline__ void pdl_launder(P& p) {
  unsigned long long u = reinterpret_cast<unsigned long long>(p);
  asm volatile("" : "+l"(u) : : "memory");
  p = reinterpret_cast<P>(u);
}
template <class... P> __device__ __forceinline__ void pdl_wait(P&... p) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  (pdl_launder(p), ...);
}

// ---------------------------------------------------------------------------------------------
template <class FT> __device__ __forceinline__ FT fmax_(FT a, FT b) { return a > b ? a : b; }
template <class FT> __device__ __forceinline__ FT fmin_(FT a, FT b) { return a < b ? a : b; }
__device__ __forceinline__ float  pow_(float a, float b)  { return powf(a, b); }
__device__ __forceinline__ double pow_(double a, double b) { return pow(a, b); }
__device__ __forceinline__ float  log_(float a)  { return logf(a); }
__device__ __forceinline__ double log_(double a) { return log(a); }
__device__ __forceinline__ float  exp_(float a)  { return expf(a); }
__device__ __forceinline__ double exp_(double a) { return exp(a); }
// reciprocal (IEEE division; a MUFU.RCP + Newton variant measured slower in k2_exp_a: 171 vs 163 µs)
__device__ __forceinline__ float  rcp_(float x)  { return 1.0f / x; }
__device__ __forceinline__ double rcp_(double x) { return 1.0 / x; }
__device__ __forceinline__ float  fma_(float a, float b, float c)  { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float  abs_(float a)  { return fabsf(a); }
__device__ __forceinline__ double abs_(double a) { return fabs(a); }

// reciprocal of a NORMAL number: MUFU.RCP + one Newton step (≤ 1 ulp) for Float32 — the IEEE division expands to ≈12
// instructions with a slow-path branch — plain division for Float64.  Used by the packed (k5_*) kernels.
__device__ __forceinline__ float rcpn_(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return fmaf(fmaf(-x, r, 1.0f), r, r);
}
__device__ __forceinline__ double rcpn_(double x) { return 1.0 / x; }

template <class FT> __device__ __forceinline__ FT pow7(FT x) {
  FT x2 = x * x, x4 = x2 * x2;
  return x4 * x2 * x;
}

// Dry thermodynamics + hydrostatic reference state at one point
// (precomputed_quantities.jl:733-815 dry branch; refstate_thermodynamics.jl:22-168).
template <class FT>
struct Pt {
  FT T, p, h, Pi, thp /*θ_v-θ_vr*/, thv, phir, sdr, lnPi;
};
template <class FT, bool FASTRCP = false>
__device__ __forceinline__ Pt<FT> thermo(const Par<FT>& P, FT rho, FT rhoe, FT K, FT Phi) {
  Pt<FT> o;
  // one reciprocal for ρ and Π, one log shared by Π = exp(κ ln(p/p0)) and ln Π (the first version used
  // powf + logf + 5 divisions per point; the XU pipe showed up at 12–16 % in ncu)
  FT etot = rhoe * (FASTRCP ? rcpn_(rho) : rcp_(rho));
  FT eint = etot - K - Phi;
  // e_int = cv_d (T − T_0) − R_d T_0  (docs/src/thermodynamics.md:103-111)
  o.T = fmax_(P.T_min_sgs, P.T_0 + (eint + P.RT0) * P.icv);
  o.h = etot + P.R_d * o.T;
  o.p = rho * P.R_d * o.T;
  FT lnPi = P.kappa * log_(o.p * P.ip0);
  o.Pi = exp_(lnPi);
  o.lnPi = lnPi;
  FT rPi = FASTRCP ? rcpn_(o.Pi) : rcp_(o.Pi);
  FT Pi7 = pow7(o.Pi);
  FT Tr = P.Tmin_ref + (P.Ts_ref - P.Tmin_ref) * Pi7;
  o.thv = o.T * rPi;
  o.thp = (o.T - Tr) * rPi;
  o.phir = -P.cp_d * (P.Tmin_ref * lnPi + P.dTs7 * (Pi7 - FT(1)));
  o.sdr = P.cp_d * (Tr - P.T_0) + o.phir;
  return o;
}


// ---- vertical pressure-gradient differences of the u₃ equation (implicit_tendency.jl:292-293):
//     ᶠgradᵥΦ − ᶠgradᵥΦ_r(p) + cp_d ᶠinterp(θ_v − θ_vr) ᶠgradᵥΠ
// The reference subtracts the level values (Φ_r up to 6·10⁵ J/kg at 60 km, Π = O(1)); in Float32 that rounding alone is 1.5·10⁻⁵ of
// u₃ after one step (tests/test_oracle_identities.py::test_float32_floor_of_the_reference_formulation).  The Float32 kernels therefore
// evaluate the two differences between adjacent levels in DIFFERENCE FORM from one log of the pressure ratio:
//     Δ = κ·log(p_hi/p_lo),  ΔΠ = Π_lo·expm1(Δ),  Δ(Π⁷) = Π_lo⁷·((1 + e)⁷ − 1) with e = expm1(Δ),
//     ΔΦ_r = −cp_d (T_min_ref·Δ + (T_surf_ref − T_min_ref)/7 · Δ(Π⁷))
// — the same mathematical quantities, 2–3× closer to the Float64 result.  The per-level slab that carried Φ_r carries p instead.
// Float64 keeps the reference's literal differences (bit-compatible with the oracle).
template <class FT> __device__ __forceinline__ FT pgf_aux(const Pt<FT>& t);
template <> __device__ __forceinline__ double pgf_aux<double>(const Pt<double>& t) { return t.phir; }
template <> __device__ __forceinline__ float pgf_aux<float>(const Pt<float>& t) { return t.p; }
// (1 + e)⁷ − 1 = e·(7 + 21e + 35e² + 35e³ + 21e⁴ + 7e⁵ + e⁶)
template <class FT> __device__ __forceinline__ FT pow7m1(FT e) {
  return e * (FT(7) + e * (FT(21) + e * (FT(35) + e * (FT(35) + e * (FT(21) + e * (FT(7) + e))))));
}
__device__ __forceinline__ void pgf_diff(const Par<double>&, double Pilo, double Pihi, double qlo, double qhi, double& dPi, double& dphr) {
  dPi = Pihi - Pilo; dphr = qhi - qlo;  // q = Φ_r
}
__device__ __forceinline__ void pgf_diff(const Par<float>& P, float Pilo, float /*Pihi*/, float plo, float phi, float& dPi, float& dphr) {
  const float dl = P.kappa * logf(phi / plo);  // q = p
  const float e = expm1f(dl);
  dPi = Pilo * e;
  dphr = -P.cp_d * (P.Tmin_ref * dl + P.dTs7 * (pow7(Pilo) * pow7m1(e)));
}

// Stage one component slab (nlev levels per node) from global into shared memory.
template <class FT>
__device__ __forceinline__ void load_slab(FT* s, const FT* __restrict__ g, int nlev) {
  for (int idx = threadIdx.x; idx < NN * LV; idx += NT) {
    int n = idx >> 6, v = idx & 63;
    if (v < nlev) s[n * LVP + v] = g[n * nlev + v];
  }
}
template <class FT>
__device__ __forceinline__ void load_hgeo(FT* s, const FT* __restrict__ hgeo, int h) {
  for (int idx = threadIdx.x; idx < HG_ELEM * 16; idx += NT) s[idx] = hgeo[(size_t)h * HG_N * 16 + idx];
}
template <class FT>
__device__ __forceinline__ void load_vlev(VLev<FT>* s, const VLev<FT>* __restrict__ g) {
  const FT* gp = reinterpret_cast<const FT*>(g);
  FT* sp = reinterpret_cast<FT*>(s);
  for (int idx = threadIdx.x; idx < (int)(sizeof(VLev<FT>) / sizeof(FT)); idx += NT) sp[idx] = gp[idx];
}

}  // namespace b200
