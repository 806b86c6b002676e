// kernels_explicit.cuh — explicit ("remaining") tendency, one element per CTA.
//
//   k_texp_a  everything of remaining_tendency! that does not need DSSed data
//             (src/prognostic_equations/remaining_tendency.jl:48-58):
//               horizontal_dynamics_tendency!          advection.jl:36-91
//               prep_hyperdiffusion_tendency!          hyperdiffusion.jl:116-147   → H = (∇²u, ∇²s_d)
//               explicit_vertical_advection_tendency!  advection.jl:205-290
//               Rayleigh/viscous sponges               remaining_tendency.jl:166-167, viscous_sponge.jl:138-175
//   k_texp_c  apply_hyperdiffusion_tendency!           hyperdiffusion.jl:247-307 (after the DSS of H)
//
// Precomputed quantities (ᶜK, ᶜT, ᶜp, ᶜh_tot, ᶠu³, ᶜu) are recomputed in-kernel from Y instead
// of being re-read from p.precomputed (pointwise, cheaper than the HBM traffic).
// Metric terms use the factored flat-shell geometry: a 2-D per-node part (J2, G^{ab}, G_ab) and
// per-level scale factors, so no 3-D LocalGeometry is ever streamed.
//
// Horizontal operators on a level reduce to 4x4 contractions with D (strong) or Dw (weak,
// Dw[i][k] = -D[k][i] w_k/w_i): weak(op) = strong(op) with D → Dw.
#pragma once
#include "common.cuh"
#include "kernels_implicit.cuh"

namespace b200 {

template <class FT>
__global__ void __launch_bounds__(NT) k_texp_a(Par<FT> P, const FT* __restrict__ hgeo, const VLev<FT>* __restrict__ vlev,
                                               const FT* __restrict__ Yc, const FT* __restrict__ Yf, FT* Ytc, FT* Ytf,
                                               FT* H) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<FT> sm(smem_raw);
  VLev<FT>& V = *reinterpret_cast<VLev<FT>*>(sm.take(sizeof(VLev<FT>) / sizeof(FT)));
  FT* hg = sm.take(HG_ELEM * 16);
  FT *rho = sm.take(SLAB), *u1 = sm.take(SLAB), *u2 = sm.take(SLAB), *re = sm.take(SLAB), *u3 = sm.take(SLAB);
  FT *sK = sm.take(SLAB), *hh = sm.take(SLAB), *Pi = sm.take(SLAB), *thp = sm.take(SLAB), *sE = sm.take(SLAB);
  FT *ss = sm.take(SLAB), *sd = sm.take(SLAB), *U1 = sm.take(SLAB), *U2 = sm.take(SLAB);
  FT *D2 = sm.take(SLAB), *ze = sm.take(SLAB), *P1 = sm.take(SLAB), *P2 = sm.take(SLAB), *Q1 = sm.take(SLAB), *Q2 = sm.take(SLAB);
  FT *X1 = sm.take(SLAB), *X2 = sm.take(SLAB);
  // viscous-sponge second-pass inputs alias slabs that are dead after pass 2
  FT *R1 = Pi, *R2 = thp, *S1 = sE, *S2 = hh;
  const int h = blockIdx.x, nv = P.nv, nf = nv + 1;
  const FT* D = V.D;
  const FT* Dw = V.Dw;
  load_vlev(&V, vlev);
  load_hgeo(hg, hgeo, h);
  {
    const FT* gYc = Yc + (size_t)h * 4 * 16 * nv;
    load_slab(rho, gYc, nv); load_slab(u1, gYc + 16 * nv, nv); load_slab(u2, gYc + 32 * nv, nv);
    load_slab(re, gYc + 48 * nv, nv); load_slab(u3, Yf + (size_t)h * 16 * nf, nf);
  }
  __syncthreads();
  // ---- pass 1: pointwise centre quantities
  for (int idx = threadIdx.x; idx < NN * LV; idx += NT) {
    int n = idx >> 6, v = idx & 63;
    if (v >= nv) continue;
    int o = n * LVP + v;
    FT a1 = u1[o], a2 = u2[o];
    FT c1 = hg[HG_GI11 * 16 + n] * a1 + hg[HG_GI12 * 16 + n] * a2;
    FT c2 = hg[HG_GI12 * 16 + n] * a1 + hg[HG_GI22 * 16 + n] * a2;
    FT lo = u3[o], hi = u3[o + 1];
    FT K = FT(0.5) * ((a1 * c1 + a2 * c2) * V.sc2i[v] + FT(0.5) * (lo * (V.g33f[v] * lo) + hi * (V.g33f[v + 1] * hi)));
    Pt<FT> t = thermo(P, rho[o], re[o], K, V.phic[v]);
    sK[o] = K; hh[o] = t.h; Pi[o] = t.Pi; thp[o] = t.thp;
    sE[o] = (K + V.phic[v]) - t.phir;
    FT sdv = P.cp_d * (t.T - P.T_0) + V.phic[v];
    sd[o] = sdv; ss[o] = sdv - t.sdr;
    U1[o] = hg[HG_J2 * 16 + n] * c1; U2[o] = hg[HG_J2 * 16 + n] * c2;
  }
  __syncthreads();
  // ---- pass 2: first derivatives; tendencies accumulate in registers
  FT rt[NIT], et[NIT], t1[NIT], t2[NIT], t3[NIT];
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int idx = threadIdx.x + it * NT;
    const int n = idx >> 6, v = idx & 63, i = n & 3, j = n >> 2;
    const int o = n * LVP + v;
    rt[it] = et[it] = t1[it] = t2[it] = t3[it] = FT(0);
    const FT rJ2 = hg[HG_RJ2 * 16 + n], J2 = hg[HG_J2 * 16 + n];
    const FT gi11 = hg[HG_GI11 * 16 + n], gi12 = hg[HG_GI12 * 16 + n], gi22 = hg[HG_GI22 * 16 + n];
    if (v < nv) {
      const FT sc = V.sc2i[v];
      // horizontal_dynamics_tendency! (advection.jl:48,59,82-88)
      FT wdivF = (d1p(Dw, rho, U1, i, j, v) + d2p(Dw, rho, U2, i, j, v)) * rJ2 * sc;
      FT wdivFh = (d1p3(Dw, rho, U1, hh, i, j, v) + d2p3(Dw, rho, U2, hh, i, j, v)) * rJ2 * sc;
      FT gh1 = d1(D, hh, i, j, v), gh2 = d2(D, hh, i, j, v);
      FT F1 = rho[o] * U1[o], F2 = rho[o] * U2[o];
      rt[it] = -wdivF;
      et[it] = -(FT(0.5) * wdivFh + FT(0.5) * (hh[o] * wdivF + (F1 * gh1 + F2 * gh2) * rJ2 * sc));
      FT th = thp[o], pi = Pi[o];
      t1[it] = -(d1(D, sE, i, j, v) +
                 P.cp_d * (th * d1(D, Pi, i, j, v) + d1p(D, thp, Pi, i, j, v) - pi * d1(D, thp, i, j, v)) / FT(2));
      t2[it] = -(d2(D, sE, i, j, v) +
                 P.cp_d * (th * d2(D, Pi, i, j, v) + d2p(D, thp, Pi, i, j, v) - pi * d2(D, thp, i, j, v)) / FT(2));
      // prep_hyperdiffusion_tendency! first level (hyperdiffusion.jl:141-147) — also feeds the viscous sponge
      D2[o] = (d1(D, U1, i, j, v) + d2(D, U2, i, j, v)) * rJ2;
      ze[o] = (d1(D, u2, i, j, v) - d2(D, u1, i, j, v)) * rJ2;
      FT g31 = FT(0.5) * (d1(D, u3, i, j, v) + d1(D, u3, i, j, v + 1));
      FT g32 = FT(0.5) * (d2(D, u3, i, j, v) + d2(D, u3, i, j, v + 1));
      P1[o] = J2 * (gi11 * g31 + gi12 * g32); P2[o] = J2 * (gi12 * g31 + gi22 * g32);
      FT gs1 = d1(D, ss, i, j, v), gs2 = d2(D, ss, i, j, v);
      Q1[o] = J2 * (gi11 * gs1 + gi12 * gs2); Q2[o] = J2 * (gi12 * gs1 + gi22 * gs2);
      // explicit_vertical_advection_tendency!: (ᶜf³ + ᶜω³) × CT12(ᶜu)  (advection.jl:228,275-277)
      FT wz = sc * (d1(Dw, u2, i, j, v) - d2(Dw, u1, i, j, v)) * rJ2;
      FT tot = hg[HG_COR3 * 16 + n] + wz;
      t1[it] += tot * U2[o];
      t2[it] -= tot * U1[o];
      if (P.rayleigh) { t1[it] -= V.bruh[v] * u1[o]; t2[it] -= V.bruh[v] * u2[o]; }
    }
    if (v < nf) {
      // ᶠω¹² = ᶠcurlᵥ(uₕ) + CT12(wcurlₕ(u₃)) (advection.jl:233,237), times ᶠJ
      const bool interior = (v > 0 && v < nv);
      FT jt1 = J2 * V.sf[v] * V.dzf[v] * hg[HG_COR1 * 16 + n] + d2(Dw, u3, i, j, v);
      FT jt2 = J2 * V.sf[v] * V.dzf[v] * hg[HG_COR2 * 16 + n] - d1(Dw, u3, i, j, v);
      FT Vn, ub1, ub2, dK = FT(0);
      if (interior) {
        jt1 -= (u2[o] - u2[o - 1]);
        jt2 += (u1[o] - u1[o - 1]);
        Vn = rho_mface(V, rho, o, v);
        ub1 = FT(0.5) * (U1[o - 1] * V.sc2i[v - 1] + U1[o] * V.sc2i[v]) * rJ2;
        ub2 = FT(0.5) * (U2[o - 1] * V.sc2i[v - 1] + U2[o] * V.sc2i[v]) * rJ2;
        dK = sK[o] - sK[o - 1];
      } else if (v == 0) {
        Vn = rho[o] * V.mc[0];
        ub1 = U1[o] * V.sc2i[0] * rJ2; ub2 = U2[o] * V.sc2i[0] * rJ2;
      } else {
        Vn = rho[o - 1] * V.mc[nv - 1];
        ub1 = U1[o - 1] * V.sc2i[nv - 1] * rJ2; ub2 = U2[o - 1] * V.sc2i[nv - 1] * rJ2;
      }
      Vn *= V.g33f[v] * u3[o];
      X1[o] = jt2 * Vn; X2[o] = -jt1 * Vn;
      t3[it] = -(jt1 * ub2 - jt2 * ub1) - dK;
    }
  }
  __syncthreads();
  // ---- pass 2b: viscous-sponge fluxes (viscous_sponge.jl:64,79)
  if (P.viscous) {
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int idx = threadIdx.x + it * NT;
      const int n = idx >> 6, v = idx & 63, i = n & 3, j = n >> 2;
      const int o = n * LVP + v;
      const FT J2 = hg[HG_J2 * 16 + n];
      const FT gi11 = hg[HG_GI11 * 16 + n], gi12 = hg[HG_GI12 * 16 + n], gi22 = hg[HG_GI22 * 16 + n];
      FT r1 = FT(0), r2 = FT(0), s1 = FT(0), s2 = FT(0);
      if (v < nf) {
        FT g1 = d1(D, u3, i, j, v), g2 = d2(D, u3, i, j, v);
        r1 = J2 * (gi11 * g1 + gi12 * g2); r2 = J2 * (gi12 * g1 + gi22 * g2);
      }
      if (v < nv) {
        FT g1 = d1(D, sd, i, j, v), g2 = d2(D, sd, i, j, v);
        s1 = rho[o] * J2 * (gi11 * g1 + gi12 * g2); s2 = rho[o] * J2 * (gi12 * g1 + gi22 * g2);
      }
      if (v < nf) { R1[o] = r1; R2[o] = r2; }
      if (v < nv) { S1[o] = s1; S2[o] = s2; }
    }
    __syncthreads();
  }
  // ---- pass 3: second derivatives, ∇² outputs, finish tendencies
  FT* gT = Ytc + (size_t)h * 4 * 16 * nv;
  FT* gF = Ytf + (size_t)h * 16 * nf;
  FT* gH = H ? H + (size_t)h * 4 * 16 * nv : nullptr;
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int idx = threadIdx.x + it * NT;
    const int n = idx >> 6, v = idx & 63, i = n & 3, j = n >> 2;
    const int o = n * LVP + v;
    const FT rJ2 = hg[HG_RJ2 * 16 + n];
    if (v < nv) {
      const FT sc = V.sc2i[v];
      FT dD1 = d1(Dw, D2, i, j, v), dD2 = d2(Dw, D2, i, j, v);
      FT dz1 = d1(Dw, ze, i, j, v), dz2 = d2(Dw, ze, i, j, v);
      FT gc11 = hg[HG_GC11 * 16 + n], gc12 = hg[HG_GC12 * 16 + n], gc22 = hg[HG_GC22 * 16 + n];
      FT L1 = sc * (dD1 - (gc11 * dz2 - gc12 * dz1) * rJ2);
      FT L2 = sc * (dD2 - (gc12 * dz2 - gc22 * dz1) * rJ2);
      if (gH) {
        FT L3 = sc * (d1(Dw, P1, i, j, v) + d2(Dw, P2, i, j, v)) * rJ2;
        FT Ls = sc * (d1(Dw, Q1, i, j, v) + d2(Dw, Q2, i, j, v)) * rJ2;
        gH[(0 * 16 + n) * nv + v] = L1; gH[(1 * 16 + n) * nv + v] = L2;
        gH[(2 * 16 + n) * nv + v] = L3; gH[(3 * 16 + n) * nv + v] = Ls;
      }
      // ᶜinterp((ᶠf¹²+ᶠω¹²) × (ᶠinterp(ρJ) ᶠu³)) / (ρJ)  (advection.jl:273-276)
      FT rm = rho[o] * V.mc[v];
      t1[it] -= FT(0.5) * (X1[o] + X1[o + 1]) / rm;
      t2[it] -= FT(0.5) * (X2[o] + X2[o + 1]) / rm;
      if (P.viscous) {
        FT b = V.bvc[v];
        t1[it] += b * L1; t2[it] += b * L2;
        et[it] += b * (sc * (d1(Dw, S1, i, j, v) + d2(Dw, S2, i, j, v)) * rJ2);
      }
      gT[(0 * 16 + n) * nv + v] = rt[it]; gT[(1 * 16 + n) * nv + v] = t1[it];
      gT[(2 * 16 + n) * nv + v] = t2[it]; gT[(3 * 16 + n) * nv + v] = et[it];
    }
    if (v < nf) {
      if (P.viscous) t3[it] += V.bvf[v] * (V.sf2i[v] * (d1(Dw, R1, i, j, v) + d2(Dw, R2, i, j, v)) * rJ2);
      gF[n * nf + v] = t3[it];
    }
  }
}

// ---------------------------------------------------------------------------------------------
template <class FT>
__global__ void __launch_bounds__(NT) k_texp_c(Par<FT> P, const FT* __restrict__ hgeo, const VLev<FT>* __restrict__ vlev,
                                               const FT* __restrict__ Yc, const FT* __restrict__ H, FT* Ytc, FT* Ytf) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<FT> sm(smem_raw);
  VLev<FT>& V = *reinterpret_cast<VLev<FT>*>(sm.take(sizeof(VLev<FT>) / sizeof(FT)));
  FT* hg = sm.take(HG_ELEM * 16);
  FT *rho = sm.take(SLAB), *L1 = sm.take(SLAB), *L2 = sm.take(SLAB), *L3 = sm.take(SLAB), *Ls = sm.take(SLAB);
  FT *U1 = sm.take(SLAB), *U2 = sm.take(SLAB);
  FT *D2 = sm.take(SLAB), *ze = sm.take(SLAB), *P1 = sm.take(SLAB), *P2 = sm.take(SLAB), *Q1 = sm.take(SLAB), *Q2 = sm.take(SLAB);
  FT* q3 = U1;  // reused after pass 2
  const int h = blockIdx.x, nv = P.nv, nf = nv + 1;
  const FT* D = V.D;
  const FT* Dw = V.Dw;
  load_vlev(&V, vlev);
  load_hgeo(hg, hgeo, h);
  {
    const FT* gH = H + (size_t)h * 4 * 16 * nv;
    load_slab(rho, Yc + (size_t)h * 4 * 16 * nv, nv);
    load_slab(L1, gH, nv); load_slab(L2, gH + 16 * nv, nv); load_slab(L3, gH + 32 * nv, nv); load_slab(Ls, gH + 48 * nv, nv);
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < NN * LV; idx += NT) {
    int n = idx >> 6, v = idx & 63;
    if (v >= nv) continue;
    int o = n * LVP + v;
    FT J2 = hg[HG_J2 * 16 + n];
    U1[o] = J2 * (hg[HG_GI11 * 16 + n] * L1[o] + hg[HG_GI12 * 16 + n] * L2[o]);
    U2[o] = J2 * (hg[HG_GI12 * 16 + n] * L1[o] + hg[HG_GI22 * 16 + n] * L2[o]);
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int idx = threadIdx.x + it * NT;
    const int n = idx >> 6, v = idx & 63, i = n & 3, j = n >> 2;
    const int o = n * LVP + v;
    if (v >= nv) continue;
    const FT rJ2 = hg[HG_RJ2 * 16 + n], J2 = hg[HG_J2 * 16 + n];
    const FT gi11 = hg[HG_GI11 * 16 + n], gi12 = hg[HG_GI12 * 16 + n], gi22 = hg[HG_GI22 * 16 + n];
    D2[o] = (d1(D, U1, i, j, v) + d2(D, U2, i, j, v)) * rJ2;
    ze[o] = (d1(D, L2, i, j, v) - d2(D, L1, i, j, v)) * rJ2;
    FT g1 = d1(D, L3, i, j, v), g2 = d2(D, L3, i, j, v);
    P1[o] = J2 * (gi11 * g1 + gi12 * g2); P2[o] = J2 * (gi12 * g1 + gi22 * g2);
    g1 = d1(D, Ls, i, j, v); g2 = d2(D, Ls, i, j, v);
    Q1[o] = rho[o] * J2 * (gi11 * g1 + gi12 * g2); Q2[o] = rho[o] * J2 * (gi12 * g1 + gi22 * g2);
  }
  __syncthreads();
  FT* gT = Ytc + (size_t)h * 4 * 16 * nv;
  FT* gF = Ytf + (size_t)h * 16 * nf;
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int idx = threadIdx.x + it * NT;
    const int n = idx >> 6, v = idx & 63, i = n & 3, j = n >> 2;
    const int o = n * LVP + v;
    if (v >= nv) continue;
    const FT rJ2 = hg[HG_RJ2 * 16 + n], sc = V.sc2i[v];
    FT dD1 = d1(Dw, D2, i, j, v), dD2 = d2(Dw, D2, i, j, v);
    FT dz1 = d1(Dw, ze, i, j, v), dz2 = d2(Dw, ze, i, j, v);
    FT gc11 = hg[HG_GC11 * 16 + n], gc12 = hg[HG_GC12 * 16 + n], gc22 = hg[HG_GC22 * 16 + n];
    // ∇⁴u = δ_div·wgradₕ(divₕ(∇²u)) − wcurlₕ(curlₕ(∇²u))  (hyperdiffusion.jl:273-277)
    FT Qa = sc * (P.ddf * dD1 - (gc11 * dz2 - gc12 * dz1) * rJ2);
    FT Qb = sc * (P.ddf * dD2 - (gc12 * dz2 - gc22 * dz1) * rJ2);
    FT Qc = sc * (d1(Dw, P1, i, j, v) + d2(Dw, P2, i, j, v)) * rJ2;
    FT Le = sc * (d1(Dw, Q1, i, j, v) + d2(Dw, Q2, i, j, v)) * rJ2;
    gT[(1 * 16 + n) * nv + v] -= P.nu4v * Qa;
    gT[(2 * 16 + n) * nv + v] -= P.nu4v * Qb;
    gT[(3 * 16 + n) * nv + v] -= P.nu4s * Le;
    q3[o] = Qc;
  }
  __syncthreads();
  // Yₜ.f.u₃ −= ν₄ᵥ ᶠwinterp(ᶜJ ρ, C3(∇⁴u))  (hyperdiffusion.jl:277)
  for (int idx = threadIdx.x; idx < NN * LV; idx += NT) {
    int n = idx >> 6, f = idx & 63;
    if (f >= nf) continue;
    int o = n * LVP + f;
    FT val;
    if (f == 0) val = q3[o];
    else if (f == nv) val = q3[o - 1];
    else {
      FT wl = V.mc[f - 1] * rho[o - 1], wh = V.mc[f] * rho[o];
      val = (wl * q3[o - 1] + wh * q3[o]) / (wl + wh);
    }
    gF[n * nf + f] -= P.nu4v * val;
  }
}

}  // namespace b200
