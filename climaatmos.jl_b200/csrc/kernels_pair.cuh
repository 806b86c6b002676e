// kernels_pair.cuh — explicit-tendency kernels in the row layout of kernels_row.cuh with PACKED two-lane
// arithmetic (pair.cuh): the four GLL nodes of a thread's row are held as two pairs (nodes 0|1 and 2|3), so every
// 4×4 contraction costs 8 FFMA2/FMUL2 instead of 16 FFMA/FMUL and the pointwise metric algebra is halved too.
// Same arithmetic, operation order and memory traffic as k2_exp_a / k2_exp_c (which stay selectable with
// B200_EXP_KERNEL=2 for A/B runs); Float64 instantiates the same code on a scalar two-member struct.
#pragma once
#include "common.cuh"
#include "kernels_row.cuh"
#include "pair.cuh"
#include "thermo2.cuh"
#include "moist2.cuh"

namespace b200 {

__constant__ float2 c_Pf[16];   // [(w*4 + k)*2 + p] = (M_w[2p][k], M_w[2p+1][k]),  w = 0: D, 1: Dw
__constant__ double2 c_Pd[16];
template <class FT> __device__ __forceinline__ P2<FT> cP(int idx);
template <> __device__ __forceinline__ P2<float> cP<float>(int idx) {
  P2<float> r;
  r.v = *reinterpret_cast<const unsigned long long*>(&c_Pf[idx]);
  return r;
}
template <> __device__ __forceinline__ P2<double> cP<double>(int idx) { return P2<double>(c_Pd[idx].x, c_Pd[idx].y); }

template <class FT, int W>
__device__ __forceinline__ void dxi4p(const P2<FT> (&a)[2], P2<FT> (&o)[2]) {
  const FT a0 = a[0].lo(), a1 = a[0].hi(), a2 = a[1].lo(), a3 = a[1].hi();
#pragma unroll
  for (int p = 0; p < 2; ++p)
    o[p] = fma2(cP<FT>((W * 4 + 3) * 2 + p), a3, fma2(cP<FT>((W * 4 + 2) * 2 + p), a2, fma2(cP<FT>((W * 4 + 1) * 2 + p), a1, cP<FT>((W * 4 + 0) * 2 + p) * a0)));
}
// ξ²-contraction o_j = Σ_k M[j][k]·a_k over the four lanes j = lane>>3 of a level, as a REDUCE-SCATTER: every lane multiplies its
// own row by the matrix COLUMN it owns and the partial sums travel in two butterfly steps (lanes ^16, then ^8) — 3 shuffles per
// value instead of the 4 of a gather (ncu: the LSU pipe, i.e. the shuffles, is the busiest pipe of k5_exp_a).
// `m` is the column in butterfly order (ROW_COLUMNS below): m[k] = M[j ^ k][j].
template <class FT>
__device__ __forceinline__ P2<FT> shflxp(const P2<FT>& a, int mask) {
  return P2<FT>(__shfl_xor_sync(FULLM, a.lo(), mask), __shfl_xor_sync(FULLM, a.hi(), mask));
}
template <class FT>
__device__ __forceinline__ void deta4p(const P2<FT> (&a)[2], const FT (&m)[4], int /*vl*/, P2<FT> (&o)[2]) {
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const P2<FT> rC = shflxp(a[p] * m[2], 16), rD = shflxp(a[p] * m[3], 16);  // partner j^2 collects rows j^2 and j^3
    const P2<FT> qa = fma2(a[p], m[0], rC), qb = fma2(a[p], m[1], rD);          // rows j and j^1 over lanes {j, j^2}
    o[p] = qa + shflxp(qb, 8);                                                  // partner j^1 collects row j^1
  }
}
// after B200_ROW_PROLOGUE*: replace the matrix rows by the columns in butterfly order
#define ROW_COLUMNS                                                                               \
  _Pragma("unroll") for (int k = 0; k < 4; ++k) { md[k] = cM<FT>((j ^ k) * 4 + j); mw[k] = cM<FT>(16 + (j ^ k) * 4 + j); }
template <class FT, int W>
__device__ __forceinline__ void div4p(const P2<FT> (&a1)[2], const P2<FT> (&a2)[2], const FT (&m)[4], int vl, P2<FT> (&o)[2]) {
  P2<FT> t[2];
  deta4p(a2, m, vl, o);
  dxi4p<FT, W>(a1, t);
  o[0] = o[0] + t[0]; o[1] = o[1] + t[1];
}
template <class FT>
__device__ __forceinline__ void ld4p(P2<FT> (&a)[2], const FT* __restrict__ g, int nlev, int j, int v, bool ok, FT dflt) {
  FT t[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) t[i] = ok ? g[(j * 4 + i) * nlev + v] : dflt;
  a[0] = P2<FT>(t[0], t[1]); a[1] = P2<FT>(t[2], t[3]);
}
template <class FT>
__device__ __forceinline__ void st4p(const P2<FT> (&a)[2], FT* __restrict__ g, int nlev, int j, int v) {
  g[(j * 4 + 0) * nlev + v] = a[0].lo(); g[(j * 4 + 1) * nlev + v] = a[0].hi();
  g[(j * 4 + 2) * nlev + v] = a[1].lo(); g[(j * 4 + 3) * nlev + v] = a[1].hi();
}
// thread-pointer forms: g already points at (row j, level v); the four nodes of the row are nlev apart.  With a
// compile-time nlev every offset is an immediate of the load/store (no per-access address arithmetic).
template <class FT>
__device__ __forceinline__ void ld4q(P2<FT> (&a)[2], const FT* __restrict__ g, int nlev, bool ok, FT dflt) {
  FT t[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) t[i] = ok ? g[i * nlev] : dflt;
  a[0] = P2<FT>(t[0], t[1]); a[1] = P2<FT>(t[2], t[3]);
}
template <class FT>
__device__ __forceinline__ void st4q(const P2<FT> (&a)[2], FT* __restrict__ g, int nlev) {
  g[0] = a[0].lo(); g[nlev] = a[0].hi(); g[2 * nlev] = a[1].lo(); g[3 * nlev] = a[1].hi();
}
template <class FT>
__device__ __forceinline__ void sputp(FT* s, const P2<FT> (&a)[2], int j, int v) {
  s[(j * 4 + 0) * LVP + v] = a[0].lo(); s[(j * 4 + 1) * LVP + v] = a[0].hi();
  s[(j * 4 + 2) * LVP + v] = a[1].lo(); s[(j * 4 + 3) * LVP + v] = a[1].hi();
}
// Pair-layout exchange slabs for k5_exp_a: s[(2j + p)·XLV + v] holds the pair p of row j at level v as ONE 64-bit word
// (STS.64 / LDS.64: half the LSU instructions of the scalar slabs).  XLV = 68: the row stride 2·XLV pairs = 272 words ≡ 16 (mod 32)
// puts the two rows of a half-warp on disjoint banks (the scalar slabs with stride 65 were 2-way conflicted).
constexpr int XLV = 68;
constexpr int XSLAB = 8 * XLV * 2;  // FT words per pair slab
template <class FT>
__device__ __forceinline__ void sputq(FT* s, const P2<FT> (&a)[2], int j, int v) {
  P2<FT>* q = reinterpret_cast<P2<FT>*>(s);
  q[(2 * j) * XLV + v] = a[0]; q[(2 * j + 1) * XLV + v] = a[1];
}
template <class FT>
__device__ __forceinline__ void sgetq(const FT* s, P2<FT> (&a)[2], int j, int v) {
  const P2<FT>* q = reinterpret_cast<const P2<FT>*>(s);
  a[0] = q[(2 * j) * XLV + v]; a[1] = q[(2 * j + 1) * XLV + v];
}
// metric pair of component c for nodes (2p, 2p+1) of this thread's row
#define HGP(c, p) ldpair(&hg[(c) * 16 + n0 + 2 * (p)])
// J2·G^{ab}·(g1, g2): contravariant flux components scaled by J2 (a pointwise 2×2 metric product)
#define METRIC_FLUX(o1, o2, g1, g2, pre)                                                        \
  _Pragma("unroll") for (int p = 0; p < 2; ++p) {                                               \
    P2<FT> w_ = pre;                                                                            \
    o1[p] = w_ * fma2(HGP(HG_GI12, p), g2[p], HGP(HG_GI11, p) * g1[p]);                         \
    o2[p] = w_ * fma2(HGP(HG_GI22, p), g2[p], HGP(HG_GI12, p) * g1[p]);                         \
  }

// ---------------------------------------------------------------------------------------------
// NVC: compile-time number of levels (63 in every production configuration), 0 = run-time P.nv
// MOIST (microphysics_model 0M; component 4 of Y.c is the active ρq_tot): moist thermodynamic state (moist.cuh); ∇²q_tot_eff =
// wdivₕ(gradₕ(q_tot − q_tot_r(p))) → H[4] (hyperdiffusion.jl:148-165); ρ(h_eff + Φ) → Hw for the water enthalpy flux of the apply
// kernel (:293-306); the whole ρq_tot tendency of this phase — horizontal advection −split_divₕ(ρu, q_tot) into Yₜ_lim (advection.jl:121-124;
// ρu and wdivₕ(ρu) are already in registers) and the viscous sponge on the total water, whose aggregate tendency also enters ρ and whose
// enthalpy flux enters ρe_tot (viscous_sponge.jl:158-199).  k5_tracer_a then only serves the passive tracers.  The dry instantiations are
// unchanged.
template <class FT, int NVC, bool MOIST = false>
__global__ void __launch_bounds__(CT, (sizeof(FT) == 4 ? 2 : 1))
k5_exp_a(Par<FT> P, const FT* __restrict__ hgeo, const VLev<FT>* __restrict__ vlev, const FT* __restrict__ Yc,
         const FT* __restrict__ Yf, FT* __restrict__ Ytc, FT* __restrict__ Ytf, FT* __restrict__ H, FT* __restrict__ Hw = nullptr,
         FT* __restrict__ Ylc = nullptr) {
  using V = P2<FT>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FT* hg = reinterpret_cast<FT*>(smem_raw);
  FT* sx = hg + HG_ELEM * 16;
  FT *s_u3 = sx, *s_r = sx + XSLAB, *s_u1 = sx + 2 * XSLAB, *s_u2 = sx + 3 * XSLAB, *s_U1 = sx + 4 * XSLAB, *s_U2 = sx + 5 * XSLAB,
     *s_K = sx + 6 * XSLAB, *s_X1 = sx + 7 * XSLAB, *s_X2 = sx + 8 * XSLAB;
  pdl_launch();
  B200_ROW_PROLOGUE_NV(NVC)
  ROW_COLUMNS
  pdl_wait(Yc, Yf, Ytc, Ytf, H);
  const bool interior = v > 0 && v < nv;
  const size_t offc = (size_t)e * P.ncf * 16 * nv + (n0 * nv + v), offf = (size_t)e * 16 * nf + (n0 * nf + v);  // (row j, level v)
  const FT* gY = Yc + offc;
  V rho[2], u1[2], u2[2], re[2], u3[2], U1[2], U2[2];
  ld4q(rho, gY, nv, cv, FT(1)); ld4q(u1, gY + 16 * nv, nv, cv, FT(0)); ld4q(u2, gY + 32 * nv, nv, cv, FT(0));
  ld4q(re, gY + 48 * nv, nv, cv, FT(0)); ld4q(u3, Yf + offf, nf, fv, FT(0));
  V rq[2], qs[2], qe[2], hw[2];  // MOIST: ρq_tot, q_tot, q_tot − q_tot_r(p), h_eff + Φ
  if (MOIST) ld4q(rq, gY + 64 * nv, nv, cv, FT(0));
  sputq(s_u3, u3, j, v); sputq(s_r, rho, j, v); sputq(s_u1, u1, j, v); sputq(s_u2, u2, j, v);
  __syncthreads();  // hg + first exchange slabs
  V c1[2], c2[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    c1[p] = fma2(HGP(HG_GI12, p), u2[p], HGP(HG_GI11, p) * u1[p]);
    c2[p] = fma2(HGP(HG_GI22, p), u2[p], HGP(HG_GI12, p) * u1[p]);
    U1[p] = HGP(HG_J2, p) * c1[p]; U2[p] = HGP(HG_J2, p) * c2[p];
  }
  sputq(s_U1, U1, j, v); sputq(s_U2, U2, j, v);
  V K[2], hh[2], ss[2], sd[2], Pi[2], th[2], sE[2], u3c[2], hs_e[2], hs_d[2];
  {
    V u3h[2];
    sgetq(s_u3, u3h, j, v < nv ? v + 1 : v);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      V kh = fma2(u2[p], c2[p], u1[p] * c1[p]) * L.sc;
      V kv = fma2(u3h[p], u3h[p] * L.g33hi, u3[p] * (u3[p] * L.g33lo)) * FT(0.5);
      K[p] = (kh + kv) * FT(0.5);
      u3c[p] = (u3[p] + u3h[p]) * FT(0.5);
      Pt2<FT> t;
      if constexpr (MOIST) {
        Mst2<FT> m;
        t = thermo2m(P, rho[p], re[p], rq[p], K[p], L.phi, m);
        qs[p] = div2(rq[p], rho[p]);
        qe[p] = qs[p] - q_tot_r2(P, t);
        hw[p] = h_eff_plus_phi2(P, m, L.phi);
      } else {
        t = thermo2(P, rho[p], re[p], K[p], L.phi);
      }
      hh[p] = t.h; Pi[p] = t.Pi; th[p] = t.thp;
      sE[p] = (K[p] + L.phi) - t.phir;
      sd[p] = fma2(t.T - P.T_0, P.cp_d, V(L.phi));
      ss[p] = sd[p] - t.sdr;
      hs_e[p] = V(FT(0)); hs_d[p] = V(FT(0));
      if (P.hs) {  // Held–Suarez forcing (held_suarez.jl:111-296), scalar per node
        FT he[2], hd[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const FT s2 = hg[HG_SIN2 * 16 + n0 + 2 * p + q], cc2 = hg[HG_COS2 * 16 + n0 + 2 * p + q];
          const FT r = q ? rho[p].hi() : rho[p].lo(), tp = q ? t.p.hi() : t.p.lo(), tl = q ? t.lnPi.hi() : t.lnPi.lo(),
                   tP = q ? t.Pi.hi() : t.Pi.lo();
          FT hf = fmax_(FT(0), (tp * P.hs_iMSLP - P.hs_sigb) * P.hs_isig);
          FT Teq = fmax_(P.hs_Tmin, (P.hs_Teq - P.hs_dTy * s2 - P.hs_dthz * (tl * P.hs_ikap) * cc2) * tP);
          FT dRT = (P.hs_ka + (P.hs_ks - P.hs_ka) * hf * cc2 * cc2) * r * (tp / (r * P.R_d) - Teq);
          he[q] = -dRT * P.cv_d; hd[q] = P.hs_kf * hf;
        }
        hs_e[p] = V(he[0], he[1]); hs_d[p] = V(hd[0], hd[1]);
      }
    }
  }
  sputq(s_K, K, j, v);
  FT* gT = Ytc + offc;
  FT* gH = H ? H + offc : nullptr;
  const bool any_visc = P.viscous && __any_sync(FULLM, L.bvc != FT(0));
  V rjs[2];  // sc / J2
#pragma unroll
  for (int p = 0; p < 2; ++p) rjs[p] = HGP(HG_RJ2, p) * L.sc;
  // ---- scalars: split-form flux divergences (advection.jl:48,59), viscous sponge on ρe_tot, ∇²s_d
  {
    V F1[2], F2[2], wd[2], t[2], g1[2], g2[2], G1[2], G2[2], et[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) { F1[p] = rho[p] * U1[p]; F2[p] = rho[p] * U2[p]; }
    div4p<FT, 1>(F1, F2, mw, vl, wd);
#pragma unroll
    for (int p = 0; p < 2; ++p) { wd[p] = wd[p] * rjs[p]; G1[p] = F1[p] * hh[p]; G2[p] = F2[p] * hh[p]; }
    V rt[2] = {-wd[0], -wd[1]};
    if (!MOIST && cv) st4q(rt, gT, nv);
    V gq1[2], gq2[2], limq[2], outq[2];  // MOIST: gradₕ q_tot, −split_divₕ(ρu, q_tot), sponge part of the ρq_tot tendency
    if constexpr (MOIST) {
      V Q1[2], Q2[2], tq[2];
#pragma unroll
      for (int p = 0; p < 2; ++p) { Q1[p] = F1[p] * qs[p]; Q2[p] = F2[p] * qs[p]; outq[p] = V(FT(0)); }
      div4p<FT, 1>(Q1, Q2, mw, vl, tq);
      deta4p(qs, md, vl, gq2);
      dxi4p<FT, 0>(qs, gq1);
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        V adv = fma2(F2[p], gq2[p], F1[p] * gq1[p]) * rjs[p];
        limq[p] = -((tq[p] * rjs[p]) * FT(0.5) + fma2(qs[p], wd[p], adv) * FT(0.5));
      }
    }
    div4p<FT, 1>(G1, G2, mw, vl, t);
    deta4p(hh, md, vl, g2);
    dxi4p<FT, 0>(hh, g1);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      V adv = fma2(F2[p], g2[p], F1[p] * g1[p]) * rjs[p];
      et[p] = hs_e[p] - ((t[p] * rjs[p]) * FT(0.5) + fma2(hh[p], wd[p], adv) * FT(0.5));
    }
    if (any_visc) {  // β wdivₕ(ρ gradₕ s_d)  (viscous_sponge.jl:79)
      V S1[2], S2[2];
      deta4p(sd, md, vl, g2);
      dxi4p<FT, 0>(sd, g1);
      METRIC_FLUX(S1, S2, g1, g2, rho[p] * HGP(HG_J2, p))
      div4p<FT, 1>(S1, S2, mw, vl, t);
#pragma unroll
      for (int p = 0; p < 2; ++p) et[p] = fma2(t[p] * rjs[p], L.bvc, et[p]);
      if constexpr (MOIST) {  // β wdivₕ(ρ gradₕ q_tot) → ρq_tot and ρ ; β wdivₕ(ρ (h_eff + Φ) gradₕ q_tot) → ρe_tot
        METRIC_FLUX(S1, S2, gq1, gq2, rho[p] * HGP(HG_J2, p))
        div4p<FT, 1>(S1, S2, mw, vl, t);
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          outq[p] = (t[p] * rjs[p]) * L.bvc;
          rt[p] = fma2(t[p] * rjs[p], L.bvc, rt[p]); S1[p] = S1[p] * hw[p]; S2[p] = S2[p] * hw[p];
        }
        div4p<FT, 1>(S1, S2, mw, vl, t);
#pragma unroll
        for (int p = 0; p < 2; ++p) et[p] = fma2(t[p] * rjs[p], L.bvc, et[p]);
      }
    }
    if (MOIST && cv) {
      st4q(rt, gT, nv);
      if (Ylc) { st4q(outq, gT + 64 * nv, nv); st4q(limq, Ylc + offc + 64 * nv, nv); }
      else { outq[0] = outq[0] + limq[0]; outq[1] = outq[1] + limq[1]; st4q(outq, gT + 64 * nv, nv); }
    }
    if (cv) st4q(et, gT + 48 * nv, nv);
    if (gH) {  // ∇²(s_d − s_d,r)  (hyperdiffusion.jl:142-147)
      V Q1[2], Q2[2];
      deta4p(ss, md, vl, g2);
      dxi4p<FT, 0>(ss, g1);
      METRIC_FLUX(Q1, Q2, g1, g2, HGP(HG_J2, p))
      div4p<FT, 1>(Q1, Q2, mw, vl, t);
#pragma unroll
      for (int p = 0; p < 2; ++p) t[p] = t[p] * rjs[p];
      if (cv) st4q(t, gH + 48 * nv, nv);
      if constexpr (MOIST) {  // ∇²q_tot_eff and ρ(h_eff + Φ)
        deta4p(qe, md, vl, g2);
        dxi4p<FT, 0>(qe, g1);
        METRIC_FLUX(Q1, Q2, g1, g2, HGP(HG_J2, p))
        div4p<FT, 1>(Q1, Q2, mw, vl, t);
#pragma unroll
        for (int p = 0; p < 2; ++p) { t[p] = t[p] * rjs[p]; hw[p] = hw[p] * rho[p]; }
        if (cv) { st4q(t, gH + 64 * nv, nv); st4q(hw, Hw + ((size_t)e * 16 * nv + (n0 * nv + v)), nv); }
      }
    }
  }
  // ---- momentum: split-form PGF (advection.jl:82-88)
  V t1[2], t2[2];
  {
    V tp[2], a[2], b[2], c[2], d[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) tp[p] = th[p] * Pi[p];
    const FT hcp = P.cp_d * FT(0.5);
    dxi4p<FT, 0>(sE, a); dxi4p<FT, 0>(Pi, b); dxi4p<FT, 0>(tp, c); dxi4p<FT, 0>(th, d);
#pragma unroll
    for (int p = 0; p < 2; ++p) t1[p] = -(fma2((fma2(th[p], b[p], c[p]) - Pi[p] * d[p]), hcp, a[p]));
    deta4p(sE, md, vl, a); deta4p(Pi, md, vl, b); deta4p(tp, md, vl, c); deta4p(th, md, vl, d);
#pragma unroll
    for (int p = 0; p < 2; ++p) t2[p] = -(fma2((fma2(th[p], b[p], c[p]) - Pi[p] * d[p]), hcp, a[p]));
  }
  // ---- ∇²u (hyperdiffusion.jl:141) and viscous sponge on uₕ
  {
    V D2[2], ze[2], a[2], b[2], g1[2];
    div4p<FT, 0>(U1, U2, md, vl, D2);
    deta4p(u1, md, vl, a);
    dxi4p<FT, 0>(u2, g1);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      D2[p] = D2[p] * HGP(HG_RJ2, p);
      ze[p] = (g1[p] - a[p]) * HGP(HG_RJ2, p);
    }
    V dD1[2], dz1[2];
    deta4p(D2, mw, vl, a); deta4p(ze, mw, vl, b);  // a = ∂̃₂D2, b = ∂̃₂ζ
    dxi4p<FT, 1>(D2, dD1); dxi4p<FT, 1>(ze, dz1);
    V L1[2], L2[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      V rJ2 = HGP(HG_RJ2, p);
      L1[p] = (dD1[p] - (HGP(HG_GC11, p) * b[p] - HGP(HG_GC12, p) * dz1[p]) * rJ2) * L.sc;
      L2[p] = (a[p] - (HGP(HG_GC12, p) * b[p] - HGP(HG_GC22, p) * dz1[p]) * rJ2) * L.sc;
      if (P.viscous) { t1[p] = fma2(L1[p], L.bvc, t1[p]); t2[p] = fma2(L2[p], L.bvc, t2[p]); }
    }
    if (gH && cv) { st4q(L1, gH, nv); st4q(L2, gH + 16 * nv, nv); }
    if (gH) {  // ∇²u₃ = wdivₕ(gradₕ(ᶜinterp(u₃))) on the flat shell
      V P1[2], P2_[2];
      deta4p(u3c, md, vl, a);
      dxi4p<FT, 0>(u3c, g1);
      METRIC_FLUX(P1, P2_, g1, a, HGP(HG_J2, p))
      div4p<FT, 1>(P1, P2_, mw, vl, b);
#pragma unroll
      for (int p = 0; p < 2; ++p) b[p] = b[p] * rjs[p];
      if (cv) st4q(b, gH + 32 * nv, nv);
    }
    // (ᶜf³ + ᶜω³) × CT12(ᶜu), Rayleigh sponge, Held–Suarez drag
    deta4p(u1, mw, vl, a);
    dxi4p<FT, 1>(u2, g1);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      V tot = fma2(g1[p] - a[p], rjs[p], HGP(HG_COR3, p));
      t1[p] = fma2(tot, U2[p], t1[p]); t2[p] = t2[p] - tot * U1[p];
      if (P.rayleigh) { t1[p] = t1[p] - u1[p] * L.bruh; t2[p] = t2[p] - u2[p] * L.bruh; }
      if (P.hs) { t1[p] = t1[p] - hs_d[p] * u1[p]; t2[p] = t2[p] - hs_d[p] * u2[p]; }
    }
  }
  __syncthreads();  // s_U1, s_U2, s_K complete
  // ---- face level v: ᶠω¹², mass flux, u₃ tendency (advection.jl:233-237,273-278)
  V X1[2], X2[2];
  {
    V d3[2], d3x[2], rl[2], a1[2], a2[2], b1[2], b2[2], kl[2], lap[2];
    deta4p(u3, mw, vl, d3);
    dxi4p<FT, 1>(u3, d3x);
    const int vm = v > 0 ? v - 1 : 0;
    sgetq(s_r, rl, j, vm); sgetq(s_u1, a1, j, vm); sgetq(s_u2, a2, j, vm); sgetq(s_U1, b1, j, vm); sgetq(s_U2, b2, j, vm); sgetq(s_K, kl, j, vm);
    const bool any_v3 = P.viscous && __any_sync(FULLM, L.bvf != FT(0));
    if (any_v3) {  // β wdivₕ(gradₕ u₃) on faces (viscous_sponge.jl:64)
      V R1[2], R2[2], g1[2], g2[2];
      deta4p(u3, md, vl, g2);
      dxi4p<FT, 0>(u3, g1);
      METRIC_FLUX(R1, R2, g1, g2, HGP(HG_J2, p))
      div4p<FT, 1>(R1, R2, mw, vl, lap);
    }
    V t3[2];
    const FT cf = L.sf * L.dzf;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const V J2 = HGP(HG_J2, p), rJ2 = HGP(HG_RJ2, p);
      V jt1 = fma2(J2 * cf, HGP(HG_COR1, p), d3[p]);
      V jt2 = fma2(J2 * cf, HGP(HG_COR2, p), -d3x[p]);
      V Vn, ub1, ub2, dk = V(FT(0));
      if (interior) {
        jt1 = jt1 - (u2[p] - a2[p]); jt2 = jt2 + (u1[p] - a1[p]);
        Vn = fma2(rho[p], L.mc, rl[p] * L.mclo) * FT(0.5);
        ub1 = (fma2(U1[p], L.sc, b1[p] * L.sclo) * FT(0.5)) * rJ2;
        ub2 = (fma2(U2[p], L.sc, b2[p] * L.sclo) * FT(0.5)) * rJ2;
        dk = K[p] - kl[p];
      } else if (v == 0) {
        Vn = rho[p] * L.mc; ub1 = (U1[p] * L.sc) * rJ2; ub2 = (U2[p] * L.sc) * rJ2;
      } else {
        Vn = rl[p] * L.mclo; ub1 = (b1[p] * L.sclo) * rJ2; ub2 = (b2[p] * L.sclo) * rJ2;
      }
      Vn = Vn * (u3[p] * L.g33lo);
      X1[p] = jt2 * Vn; X2[p] = -(jt1 * Vn);
      t3[p] = -(jt1 * ub2 - jt2 * ub1) - dk;
      if (any_v3) t3[p] = fma2((lap[p] * L.sf2i) * rJ2, L.bvf, t3[p]);
    }
    if (fv) st4q(t3, Ytf + offf, nf);
  }
  sputq(s_X1, X1, j, v); sputq(s_X2, X2, j, v);
  __syncthreads();
  if (cv) {
    V h1[2], h2[2];
    sgetq(s_X1, h1, j, v + 1); sgetq(s_X2, h2, j, v + 1);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      V irm = V(FT(0.5) * L.rmc * rcpn_(rho[p].lo()), FT(0.5) * L.rmc * rcpn_(rho[p].hi()));
      t1[p] = t1[p] - (X1[p] + h1[p]) * irm;
      t2[p] = t2[p] - (X2[p] + h2[p]) * irm;
    }
    st4q(t1, gT + 16 * nv, nv);
    st4q(t2, gT + 32 * nv, nv);
  }
}

}  // namespace b200

namespace b200 {

// ---------------------------------------------------------------------------------------------
// Passive grid-scale tracers ρχ (components 4.. of Y.c, e.g. the chemistry tracer ρq_gas_A), one (element, tracer)
// per CTA in the same row layout:
//   k5_tracer_a  horizontal_tracer_advection_tendency!  ρχₜ_lim −= split_divₕ(ρu, χ)          (advection.jl:113-143)
//                prep_tracer_hyperdiffusion_tendency!   ∇²χ = wdivₕ(gradₕ χ) → H[4+q]          (hyperdiffusion.jl:420-432)
//                explicit vertical transport            ρχₜ += −ᶜadvdivᵥ(ᶠinterp(ρJ)/ᶠJ · U(ᶠu³, χ)) with tracer_upwinding
//                                                       (advection.jl:249-255; implicit_tendency.jl:120-143)
//                viscous sponge                          ρχₜ += β wdivₕ(ρ gradₕ χ)              (viscous_sponge.jl:226-231)
//   (apply_tracer_hyperdiffusion_tendency!, ρχₜ_lim −= ν₄ₛ wdivₕ(ρ gradₕ ∇²χ), hyperdiffusion.jl:524-532: parts 3.. of k7_exp_c, kernels_lvl.cuh)
// With Ylc == nullptr (native stepper: lim! is a no-op) the limited part is accumulated into Yₜ as well.
template <class FT>
__global__ void __launch_bounds__(CT, (sizeof(FT) == 4 ? 3 : 1))
k5_tracer_a(Par<FT> P, const FT* __restrict__ hgeo, const VLev<FT>* __restrict__ vlev, const FT* __restrict__ Yc,
            const FT* __restrict__ Yf, FT* __restrict__ Ytc, FT* __restrict__ Ylc, FT* __restrict__ H) {
  using V = P2<FT>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FT* hg = reinterpret_cast<FT*>(smem_raw);
  FT* sx = hg + HG_ELEM * 16;
  FT *s_chi = sx, *s_r = sx + SLAB, *s_fx = sx + 2 * SLAB;
  B200_ROW_PROLOGUE
  ROW_COLUMNS
  // passive tracers only: in a moist (0M) context component 4 is the active ρq_tot, served by k5_exp_a<…, MOIST> / part 1 of k7_exp_c
  const int q = 4 + (P.moist ? 1 : 0) + blockIdx.y;
  const FT* gY = Yc + (size_t)e * P.ncf * 16 * nv;
  V rho[2], u1[2], u2[2], rq[2], u3[2], chi[2];
  ld4p(rho, gY, nv, j, v, cv, FT(1)); ld4p(u1, gY + 16 * nv, nv, j, v, cv, FT(0)); ld4p(u2, gY + 32 * nv, nv, j, v, cv, FT(0));
  ld4p(rq, gY + (size_t)q * 16 * nv, nv, j, v, cv, FT(0)); ld4p(u3, Yf + (size_t)e * 16 * nf, nf, j, v, fv, FT(0));
#pragma unroll
  for (int p = 0; p < 2; ++p) chi[p] = V(rq[p].lo() / rho[p].lo(), rq[p].hi() / rho[p].hi());
  sputp(s_chi, chi, j, v); sputp(s_r, rho, j, v);
  __syncthreads();
  V U1[2], U2[2], F1[2], F2[2], rjs[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    V c1 = fma2(HGP(HG_GI12, p), u2[p], HGP(HG_GI11, p) * u1[p]);
    V c2 = fma2(HGP(HG_GI22, p), u2[p], HGP(HG_GI12, p) * u1[p]);
    U1[p] = HGP(HG_J2, p) * c1; U2[p] = HGP(HG_J2, p) * c2;
    F1[p] = rho[p] * U1[p]; F2[p] = rho[p] * U2[p];
    rjs[p] = HGP(HG_RJ2, p) * L.sc;
  }
  V wd[2], t[2], g1[2], g2[2], G1[2], G2[2], lim[2], out[2];
  div4p<FT, 1>(F1, F2, mw, vl, wd);
#pragma unroll
  for (int p = 0; p < 2; ++p) { wd[p] = wd[p] * rjs[p]; G1[p] = F1[p] * chi[p]; G2[p] = F2[p] * chi[p]; }
  div4p<FT, 1>(G1, G2, mw, vl, t);
  deta4p(chi, md, vl, g2);
  dxi4p<FT, 0>(chi, g1);
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    V adv = fma2(F2[p], g2[p], F1[p] * g1[p]) * rjs[p];
    lim[p] = -((t[p] * rjs[p]) * FT(0.5) + fma2(chi[p], wd[p], adv) * FT(0.5));
    out[p] = V(FT(0));
  }
  if (H) {  // ∇²χ
    V Q1[2], Q2[2];
    METRIC_FLUX(Q1, Q2, g1, g2, HGP(HG_J2, p))
    div4p<FT, 1>(Q1, Q2, mw, vl, t);
#pragma unroll
    for (int p = 0; p < 2; ++p) t[p] = t[p] * rjs[p];
    if (cv) st4p(t, H + (size_t)e * P.ncf * 16 * nv + (size_t)q * 16 * nv, nv, j, v);
  }
  if (P.viscous && __any_sync(FULLM, L.bvc != FT(0))) {
    V S1[2], S2[2];
    METRIC_FLUX(S1, S2, g1, g2, rho[p] * HGP(HG_J2, p))
    div4p<FT, 1>(S1, S2, mw, vl, t);
#pragma unroll
    for (int p = 0; p < 2; ++p) out[p] = (t[p] * rjs[p]) * L.bvc;
  }
  // vertical transport: flux through face v of (ᶠinterp(ρJ)/J2)·u³·χ_face, zero on the boundary faces
  {
    const bool interior = v > 0 && v < nv;
    FT fx[4];
    const int vm = v > 0 ? v - 1 : 0, vm2 = v > 1 ? v - 2 : 0, vp = v < nv - 1 ? v + 1 : v;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int o = (n0 + i) * LVP;
      FT r = FT(0);
      if (interior) {
        const FT u3i = (i < 2) ? (i == 0 ? u3[0].lo() : u3[0].hi()) : (i == 2 ? u3[1].lo() : u3[1].hi());
        const FT w = L.g33lo * u3i;
        const FT am = s_chi[o + vm], ap = s_chi[o + v];
        FT face;
        if (P.tupw == 3 && v >= 2 && v <= nv - 2) {
          if (w >= FT(0)) face = am + vl_slope(s_chi[o + vm2], am, ap) / FT(2) * (FT(1) - w * P.dt);
          else face = ap - vl_slope(am, ap, s_chi[o + vp]) / FT(2) * (FT(1) + w * P.dt);
        } else if (P.tupw == 2 && nv >= 3) {  // ᶠupwind3 (tracer_upwinding: third_order)
          face = upwind3_face(s_chi[o + vm2], am, ap, s_chi[o + vp], v, nv, w);
        } else if (P.tupw == 0) {
          face = FT(0.5) * (am + ap);
        } else {
          face = w >= FT(0) ? am : ap;
        }
        r = (FT(0.5) * (s_r[o + vm] * L.mclo + s_r[o + v] * L.mc)) * (w * face);
      }
      fx[i] = r;
      if (fv) s_fx[o + v] = r;
    }
    __syncthreads();
    if (cv) {
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int o0 = (n0 + 2 * p) * LVP + v + 1, o1 = (n0 + 2 * p + 1) * LVP + v + 1;
        V up(s_fx[o0], s_fx[o1]), dn(fx[2 * p], fx[2 * p + 1]);
        out[p] = out[p] - (up - dn) * L.rmc;
      }
    }
  }
  if (cv) {
    FT* gT = Ytc + (size_t)e * P.ncf * 16 * nv + (size_t)q * 16 * nv;
    if (Ylc) {
      st4p(out, gT, nv, j, v);
      st4p(lim, Ylc + (size_t)e * P.ncf * 16 * nv + (size_t)q * 16 * nv, nv, j, v);
    } else {
      out[0] = out[0] + lim[0]; out[1] = out[1] + lim[1];
      st4p(out, gT, nv, j, v);
    }
  }
}

}  // namespace b200
