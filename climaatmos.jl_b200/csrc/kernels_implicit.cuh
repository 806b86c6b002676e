// kernels_implicit.cuh — vertically implicit part of the step, one element (16 columns) per CTA.
//
//   k_cache_imp   set_implicit_precomputed_quantities!  (src/cache/precomputed_quantities.jl:698-831)
//   k_t_imp       implicit_tendency!                    (src/prognostic_equations/implicit/implicit_tendency.jl:36-98,185-298)
//   k_wfact       update_jacobian! (advection blocks)   (.../implicit/manual_sparse_jacobian.jl:713-870)
//   k_ldiv        invert_jacobian! (BlockArrowheadSolve → Schur onto u₃ → Thomas) (:504-585,1897)
//   k_t_post_imp  correct_implicit_advection_tendency!  (implicit_tendency.jl:322-339)
//   k_imp_stage   all of the above fused into one pass over the element (native stepper)
//
// The band-matrix blocks of the reference are never materialised as fields: k_wfact stores the
// 15 per-level coefficients the solve needs (Schur tridiagonal + couplings); the fused kernel
// keeps them in shared memory only.
#pragma once
#include "common.cuh"
#include "moist.cuh"

namespace b200 {

// jacobian coefficient planes written by k_wfact, each [16][Nv+1] per element
enum { JC_L = 0, JC_D, JC_U, JC_UR_LO, JC_UR_HI, JC_UE_LO, JC_UE_HI, JC_U1_LO, JC_U1_HI, JC_U2_LO,
       JC_U2_HI, JC_RU_LO, JC_RU_HI, JC_EU_LO, JC_EU_HI, JC_N };

template <class FT>
__device__ __forceinline__ FT kinetic(const FT* hg, const VLev<FT>& V, FT u1, FT u2, FT u3lo, FT u3hi,
                                      int n, int v) {
  FT c1 = hg[HG_GI11 * 16 + n] * u1 + hg[HG_GI12 * 16 + n] * u2;
  FT c2 = hg[HG_GI12 * 16 + n] * u1 + hg[HG_GI22 * 16 + n] * u2;
  return FT(0.5) * ((u1 * c1 + u2 * c2) * V.sc2i[v] +
                    FT(0.5) * (u3lo * (V.g33f[v] * u3lo) + u3hi * (V.g33f[v + 1] * u3hi)));
}

// shared-memory carve-up helper
template <class FT>
struct Smem {
  FT* base;
  __device__ Smem(void* p) : base(reinterpret_cast<FT*>(p)) {}
  __device__ FT* take(int n) { FT* r = base; base += n; return r; }
};

// ---------------------------------------------------------------------------------------------
template <class FT>
__global__ void __launch_bounds__(NT) k_cache_imp(Par<FT> P, const FT* __restrict__ hgeo,
                                                  const VLev<FT>* __restrict__ vlev, const FT* __restrict__ Yc,
                                                  FT* Yf, FT* uc, FT* u3f, FT* Kc, FT* Tc, FT* pc, FT* hc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<FT> sm(smem_raw);
  VLev<FT>& V = *reinterpret_cast<VLev<FT>*>(sm.take(sizeof(VLev<FT>) / sizeof(FT)));
  FT* hg = sm.take(HG_ELEM * 16);
  FT* s_u3 = sm.take(SLAB);
  const int h = blockIdx.x, nv = P.nv, nf = nv + 1;
  load_vlev(&V, vlev);
  load_hgeo(hg, hgeo, h);
  FT* gYf = Yf + (size_t)h * 16 * nf;
  load_slab(s_u3, gYf, nf);
  __syncthreads();
  // set_velocity_at_surface!/top! (:486-554): flat grid ⇒ ᶠuₕ³ = 0 ⇒ u₃ = -0/g³³ = 0 on both boundaries
  if (threadIdx.x < 16) {
    int n = threadIdx.x;
    s_u3[n * LVP] = FT(0);
    s_u3[n * LVP + nv] = FT(0);
    gYf[n * nf] = FT(0);
    gYf[n * nf + nv] = FT(0);
  }
  __syncthreads();
  const FT* gYc = Yc + (size_t)h * P.ncf * 16 * nv;
  for (int idx = threadIdx.x; idx < NN * LV; idx += NT) {
    int n = idx >> 6, v = idx & 63;
    if (v < nv) {
      FT rho = gYc[(0 * 16 + n) * nv + v], u1 = gYc[(1 * 16 + n) * nv + v];
      FT u2 = gYc[(2 * 16 + n) * nv + v], re = gYc[(3 * 16 + n) * nv + v];
      FT lo = s_u3[n * LVP + v], hi = s_u3[n * LVP + v + 1];
      FT K = kinetic(hg, V, u1, u2, lo, hi, n, v);
      Pt<FT> t;
      if (P.moist) { Mst<FT> m; t = thermo_m(P, rho, re, gYc[(4 * 16 + n) * nv + v], K, V.phic[v], m); }  // :735-815 (0M branch)
      else t = thermo(P, rho, re, K, V.phic[v]);
      size_t o = ((size_t)h * 16 + n) * nv + v;
      if (Kc) Kc[o] = K;
      if (Tc) Tc[o] = t.T;
      if (pc) pc[o] = t.p;
      if (hc) hc[o] = t.h;
      if (uc) {
        size_t o3 = ((size_t)h * 3 * 16 + n) * nv + v;
        uc[o3] = u1;
        uc[o3 + (size_t)16 * nv] = u2;
        uc[o3 + (size_t)32 * nv] = FT(0.5) * (lo + hi);
      }
    }
    if (v < nf && u3f) u3f[((size_t)h * 16 + n) * nf + v] = V.g33f[v] * s_u3[n * LVP + v];
  }
}

// ---------------------------------------------------------------------------------------------
// Shared preparation: stage the state, compute per-centre thermodynamics into shared slabs.
template <class FT>
struct ImpSlabs {
  FT *rho, *u1, *u2, *re, *u3;  // state
  FT *K, *h, *Pi, *thv, *thp, *phr, *T;  // centre diagnostics
};

template <class FT>
__device__ __forceinline__ void imp_carve(Smem<FT>& sm, ImpSlabs<FT>& S) {
  S.rho = sm.take(SLAB); S.u1 = sm.take(SLAB); S.u2 = sm.take(SLAB); S.re = sm.take(SLAB); S.u3 = sm.take(SLAB);
  S.K = sm.take(SLAB); S.h = sm.take(SLAB); S.Pi = sm.take(SLAB); S.thv = sm.take(SLAB); S.thp = sm.take(SLAB);
  S.phr = sm.take(SLAB); S.T = sm.take(SLAB);
}

template <class FT>
__device__ __forceinline__ void imp_load_state(ImpSlabs<FT>& S, const FT* __restrict__ Yc,
                                               const FT* __restrict__ Yf, int h, int nv, int ncf) {
  const FT* gYc = Yc + (size_t)h * ncf * 16 * nv;
  load_slab(S.rho, gYc, nv);
  load_slab(S.u1, gYc + 16 * nv, nv);
  load_slab(S.u2, gYc + 32 * nv, nv);
  load_slab(S.re, gYc + 48 * nv, nv);
  load_slab(S.u3, Yf + (size_t)h * 16 * (nv + 1), nv + 1);
}

template <class FT>
__device__ __forceinline__ void imp_thermo(const Par<FT>& P, const FT* hg, const VLev<FT>& V, ImpSlabs<FT>& S) {
  const int nv = P.nv;
  for (int idx = threadIdx.x; idx < NN * LV; idx += NT) {
    int n = idx >> 6, v = idx & 63;
    if (v < nv) {
      int o = n * LVP + v;
      FT K = kinetic(hg, V, S.u1[o], S.u2[o], S.u3[o], S.u3[o + 1], n, v);
      Pt<FT> t = thermo(P, S.rho[o], S.re[o], K, V.phic[v]);
      S.K[o] = K; S.h[o] = t.h; S.Pi[o] = t.Pi; S.thv[o] = t.thv; S.thp[o] = t.thp; S.phr[o] = pgf_aux(t); S.T[o] = t.T;
    }
  }
}

// mass flux through interior face f (zero on the boundary faces: ᶜadvdivᵥ SetValue(0)):
// (ᶠinterp(ρJ)·u³)/J2 = ½(ρ[f-1] m_c[f-1] + ρ[f] m_c[f]) · g³³_f u₃[f]
template <class FT>
__device__ __forceinline__ FT rho_mface(const VLev<FT>& V, const FT* rho, int o, int f) {
  return FT(0.5) * (rho[o - 1] * V.mc[f - 1] + rho[o] * V.mc[f]);
}

// T_imp at one centre / face given the prepared slabs
template <class FT>
__device__ __forceinline__ void timp_center(const VLev<FT>& V, const ImpSlabs<FT>& S, int n, int v, int nv,
                                            FT& rt, FT& et) {
  int o = n * LVP + v;
  FT mlo = FT(0), mhi = FT(0), hlo = FT(0), hhi = FT(0);
  if (v > 0) { mlo = rho_mface(V, S.rho, o, v) * (V.g33f[v] * S.u3[o]); hlo = FT(0.5) * (S.h[o - 1] + S.h[o]); }
  if (v < nv - 1) { mhi = rho_mface(V, S.rho, o + 1, v + 1) * (V.g33f[v + 1] * S.u3[o + 1]); hhi = FT(0.5) * (S.h[o] + S.h[o + 1]); }
  rt = -(mhi - mlo) / V.mc[v];
  et = -(mhi * hhi - mlo * hlo) / V.mc[v];
}
template <class FT>
__device__ __forceinline__ FT timp_face(const Par<FT>& P, const VLev<FT>& V, const ImpSlabs<FT>& S, int n, int f, int nv) {
  int o = n * LVP + f;
  FT r = FT(0);
  if (f > 0 && f < nv) {
    FT dPi, dphr;  // S.phr carries Φ_r (Float64) or p (Float32): common.cuh pgf_diff
    pgf_diff(P, S.Pi[o - 1], S.Pi[o], S.phr[o - 1], S.phr[o], dPi, dphr);
    r = -(V.dphif[f] - dphr + P.cp_d * (FT(0.5) * (S.thp[o - 1] + S.thp[o])) * dPi);
  }
  if (P.rayleigh) r += -V.brw[f] * S.u3[o];
  return r;
}


// ---------------------------------------------------------------------------------------------
// Jacobian coefficients at face f / centre v (manual_sparse_jacobian.jl:746-868, dry flat grid).
template <class FT>
struct FaceCoef { FT l, d, u, ur_lo, ur_hi, ue_lo, ue_hi, u1_lo, u1_hi, u2_lo, u2_hi; };

template <class FT>
__device__ __forceinline__ void center_coef(const VLev<FT>& V, const ImpSlabs<FT>& S, FT dtg, int n, int v, int nv,
                                            FT& ru_lo, FT& ru_hi, FT& eu_lo, FT& eu_hi) {
  int o = n * LVP + v;
  ru_lo = ru_hi = eu_lo = eu_hi = FT(0);
  if (v > 0) {  // face v (lower)
    FT a = dtg * (rho_mface(V, S.rho, o, v) / V.mc[v]) * V.g33f[v];
    ru_lo = a; eu_lo = a * (FT(0.5) * (S.h[o - 1] + S.h[o]));
  }
  if (v < nv - 1) {  // face v+1 (upper)
    FT a = -dtg * (rho_mface(V, S.rho, o + 1, v + 1) / V.mc[v]) * V.g33f[v + 1];
    ru_hi = a; eu_hi = a * (FT(0.5) * (S.h[o] + S.h[o + 1]));
  }
}

template <class FT>
__device__ __forceinline__ FaceCoef<FT> face_coef(const Par<FT>& P, const FT* hg, const VLev<FT>& V,
                                                  const ImpSlabs<FT>& S, FT dtg, int n, int f, int nv) {
  FaceCoef<FT> c;
  FT beta = P.rayleigh ? V.brw[f] : FT(0);
  c.l = c.u = c.ur_lo = c.ur_hi = c.ue_lo = c.ue_hi = c.u1_lo = c.u1_hi = c.u2_lo = c.u2_hi = FT(0);
  c.d = dtg * (-beta) - FT(1);
  if (f == 0 || f == nv) return c;  // ᶠgradᵥ rows vanish on the boundary faces
  int o = n * LVP + f;  // centre f ("hi"); centre f-1 is o-1 ("lo")
  const FT kap = P.R_d / P.cv_d;
  FT rlo = S.rho[o - 1], rhi = S.rho[o];
  FT rf = FT(0.5) * (rlo + rhi);
  FT pg_lo = FT(1) / rf, pg_hi = -FT(1) / rf;
  FT dp_lo = kap * (P.T_0 * P.cp_d - S.K[o - 1] - V.phic[f - 1]) + (P.R_d - kap * P.cv_d) * S.T[o - 1];
  FT dp_hi = kap * (P.T_0 * P.cp_d - S.K[o] - V.phic[f]) + (P.R_d - kap * P.cv_d) * S.T[o];
  FT dPi, dphr_;
  pgf_diff(P, S.Pi[o - 1], S.Pi[o], S.phr[o - 1], S.phr[o], dPi, dphr_);
  FT buoy = P.cp_d * (FT(0.5) * (S.thv[o - 1] + S.thv[o])) * dPi / rf;
  c.ur_lo = dtg * (pg_lo * dp_lo + buoy * FT(0.5));
  c.ur_hi = dtg * (pg_hi * dp_hi + buoy * FT(0.5));
  c.ue_lo = dtg * pg_lo * kap;
  c.ue_hi = dtg * pg_hi * kap;
  FT x_lo = pg_lo * (-kap * rlo), x_hi = pg_hi * (-kap * rhi);
  {  // ∂K/∂uₕ = CT12(uₕ)
    FT g11 = hg[HG_GI11 * 16 + n], g12 = hg[HG_GI12 * 16 + n], g22 = hg[HG_GI22 * 16 + n];
    FT a1 = (g11 * S.u1[o - 1] + g12 * S.u2[o - 1]) * V.sc2i[f - 1];
    FT a2 = (g12 * S.u1[o - 1] + g22 * S.u2[o - 1]) * V.sc2i[f - 1];
    FT b1 = (g11 * S.u1[o] + g12 * S.u2[o]) * V.sc2i[f];
    FT b2 = (g12 * S.u1[o] + g22 * S.u2[o]) * V.sc2i[f];
    c.u1_lo = dtg * x_lo * a1; c.u1_hi = dtg * x_hi * b1;
    c.u2_lo = dtg * x_lo * a2; c.u2_hi = dtg * x_hi * b2;
  }
  // ∂K/∂u₃ rows: centre k has ½g³³u₃ at faces k and k+1
  FT km = FT(0.5) * V.g33f[f - 1] * S.u3[o - 1], k0 = FT(0.5) * V.g33f[f] * S.u3[o], kp = FT(0.5) * V.g33f[f + 1] * S.u3[o + 1];
  FT l = dtg * (x_lo * km), d = dtg * ((x_lo * k0 + x_hi * k0) - beta) - FT(1), u = dtg * (x_hi * kp);
  // Schur complement with the scalar blocks (A11 = -I): centres f-1 and f
  FT ru_lo_a, ru_hi_a, eu_lo_a, eu_hi_a, ru_lo_b, ru_hi_b, eu_lo_b, eu_hi_b;
  center_coef(V, S, dtg, n, f - 1, nv, ru_lo_a, ru_hi_a, eu_lo_a, eu_hi_a);
  center_coef(V, S, dtg, n, f, nv, ru_lo_b, ru_hi_b, eu_lo_b, eu_hi_b);
  l += c.ur_lo * ru_lo_a + c.ue_lo * eu_lo_a;
  d += c.ur_lo * ru_hi_a + c.ur_hi * ru_lo_b + c.ue_lo * eu_hi_a + c.ue_hi * eu_lo_b;
  u += c.ur_hi * ru_hi_b + c.ue_hi * eu_hi_b;
  c.l = l; c.d = d; c.u = u;
  return c;
}




// ---------------------------------------------------------------------------------------------
// van Leer (Lin 1994, MonotoneLocalExtrema) face value of χ upwinded with u³ (abbreviations.jl:250-256)
template <class FT>
__device__ __forceinline__ FT vl_slope(FT am, FT a0, FT ap) {
  FT d = ((a0 - am) + (ap - a0)) / FT(2);
  FT mn = fmin_(fmin_(am, a0), ap), mx = fmax_(fmax_(am, a0), ap);
  FT lim = fmin_(abs_(d), fmin_(FT(2) * (a0 - mn), FT(2) * (mx - a0)));
  return d > FT(0) ? lim : (d < FT(0) ? -lim : FT(0));
}
// ᶠupwind3 face value (abbreviations.jl:229-240; Upwind3rdOrderBiasedProductC2F with ThirdOrderOneSided closures [UPSTREAM-RECALL]):
// interior faces 2..nv−2 the upwind-biased (−a⁻⁻ + 5a⁻ + 2a⁺)/6 (mirror image for w < 0); the first interior face the right-biased
// (4a⁻ + 10a⁺ − 2a⁺⁺)/12 and the last one the left-biased (−2a⁻⁻ + 10a⁻ + 4a⁺)/12, whatever the sign of w.  Needs nv ≥ 3.
template <class FT>
__device__ __forceinline__ FT upwind3_face(FT amm, FT am, FT ap, FT app, int f, int nv, FT w) {
  if (f == 1) return (FT(4) * am + FT(10) * ap - FT(2) * app) / FT(12);
  if (f == nv - 1) return (-FT(2) * amm + FT(10) * am + FT(4) * ap) / FT(12);
  return w >= FT(0) ? (-amm + FT(5) * am + FT(2) * ap) / FT(6) : (FT(2) * am + FT(5) * ap - app) / FT(6);
}
// returns (upwinded χ - centred χ) at interior face f for contravariant velocity w
template <class FT>
__device__ __forceinline__ FT upwind_minus_central(const Par<FT>& P, const FT* chi, int o, int f, int nv, FT w) {
  FT am = chi[o - 1], ap = chi[o];
  FT cen = FT(0.5) * (am + ap);
  FT up;
  if (P.upwinding == 3 && f >= 2 && f <= nv - 2) {
    if (w >= FT(0)) up = am + vl_slope(chi[o - 2], am, ap) / FT(2) * (FT(1) - w * P.dt);
    else up = ap - vl_slope(am, ap, chi[o + 1]) / FT(2) * (FT(1) + w * P.dt);
  } else if (P.upwinding == 2 && nv >= 3) {
    up = upwind3_face(f >= 2 ? chi[o - 2] : am, am, ap, f <= nv - 2 ? chi[o + 1] : ap, f, nv, w);
  } else {
    up = w >= FT(0) ? am : ap;  // first order (also the FirstOrderOneSided boundary closure)
  }
  return up - cen;
}


// ---------------------------------------------------------------------------------------------

}  // namespace b200
