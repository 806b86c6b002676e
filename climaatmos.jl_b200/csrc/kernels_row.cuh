// kernels_row.cuh — explicit-tendency kernels, "one GLL row per thread" layout (round-1, 2nd generation).
//
// One element per CTA of 256 threads.  Thread (v, j) owns the four nodes i = 0..3 of row j at level v in
// registers; a warp holds 8 consecutive levels × 4 rows (lane = vl + 8·j).  On B200 this gives
//   * ξ¹-derivatives: thread-local 4×4 contractions with D/Dw entries from the constant bank;
//   * ξ²-derivatives: 4 warp shuffles (lanes vl, vl+8, vl+16, vl+24) per output with per-thread
//     matrix rows D[j][·], Dw[j][·] — no shared-memory slabs, no block barriers for horizontal work;
//   * vertical neighbours (k±1): a shared-memory column exchange, 3 block barriers per launch;
//   * global accesses: 32-byte segments of 8 consecutive levels (full sector efficiency).
// Why: ncu showed the first generation (whole slabs in shared memory, 500 LDS/point, 2 CTAs/SM) LSU- and
// latency-bound and the one-thread-per-level variant (16 nodes/thread, 255 registers, ≈150 KB of
// straight-line code) instruction-fetch-bound (profiles/r1_ncu_summary.md).  Per point this layout
// needs ≈100 shuffles, ≈12 LDS and ≈80 registers per thread, with the 4-node body reused by all rows.
//
//   k2_exp_a  everything of remaining_tendency! before the DSS (see kernels_explicit.cuh for the list)
//   k2_exp_c  apply_hyperdiffusion_tendency! after the DSS
#pragma once
#include "common.cuh"
#include "kernels_reg.cuh"

namespace b200 {

constexpr int CT = 256;

template <class FT, int W>
__device__ __forceinline__ FT dxi4(const FT (&a)[4], int i) {
  return cM<FT>(W + i * 4 + 0) * a[0] + cM<FT>(W + i * 4 + 1) * a[1] + cM<FT>(W + i * 4 + 2) * a[2] + cM<FT>(W + i * 4 + 3) * a[3];
}
// ξ²-derivative of the rows held by lanes vl + 8k with this thread's matrix row m[k] = M[j][k]
template <class FT>
__device__ __forceinline__ void deta4(const FT (&a)[4], const FT (&m)[4], int vl, FT (&o)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    FT s = m[0] * __shfl_sync(FULLM, a[i], vl);
    s += m[1] * __shfl_sync(FULLM, a[i], vl + 8);
    s += m[2] * __shfl_sync(FULLM, a[i], vl + 16);
    s += m[3] * __shfl_sync(FULLM, a[i], vl + 24);
    o[i] = s;
  }
}
// o[i] = (∂₁ a1 + ∂₂ a2)[i]  (divergence-like) with matrix set W for ξ¹ and row m for ξ²
template <class FT, int W>
__device__ __forceinline__ void div4(const FT (&a1)[4], const FT (&a2)[4], const FT (&m)[4], int vl, FT (&o)[4]) {
  deta4(a2, m, vl, o);
#pragma unroll
  for (int i = 0; i < 4; ++i) o[i] += dxi4<FT, W>(a1, i);
}

template <class FT>
__device__ __forceinline__ void ld4(FT (&a)[4], const FT* __restrict__ g, int nlev, int j, int v, bool ok, FT dflt) {
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = ok ? g[(j * 4 + i) * nlev + v] : dflt;
}
template <class FT>
__device__ __forceinline__ void sput(FT* s, const FT (&a)[4], int j, int v) {
#pragma unroll
  for (int i = 0; i < 4; ++i) s[(j * 4 + i) * LVP + v] = a[i];
}
template <class FT>
__device__ __forceinline__ void sget(const FT* s, FT (&a)[4], int j, int v) {
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = s[(j * 4 + i) * LVP + v];
}

#define B200_ROW_PROLOGUE                                                                             \
  const int e = blockIdx.x, lane = threadIdx.x & 31, vl = lane & 7, j = lane >> 3;                    \
  const int v = (threadIdx.x >> 5) * 8 + vl, nv = P.nv, nf = nv + 1;                                  \
  const bool cv = v < nv, fv = v < nf;                                                                \
  for (int k = threadIdx.x; k < HG_ELEM * 16; k += CT) hg[k] = hgeo[(size_t)e * HG_N * 16 + k];       \
  FT md[4], mw[4];                                                                                    \
  _Pragma("unroll") for (int k = 0; k < 4; ++k) { md[k] = cM<FT>(j * 4 + k); mw[k] = cM<FT>(16 + j * 4 + k); } \
  const Lev<FT> L = load_lev(vlev, v, nv);                                                            \
  const int n0 = j * 4;

#define B200_ROW_PROLOGUE_NV(NVC_)                                                                            \
  const int e = blockIdx.x, lane = threadIdx.x & 31, vl = lane & 7, j = lane >> 3;                    \
  const int v = (threadIdx.x >> 5) * 8 + vl, nv = (NVC_) ? (NVC_) : P.nv, nf = nv + 1;                                  \
  const bool cv = v < nv, fv = v < nf;                                                                \
  for (int k = threadIdx.x; k < HG_ELEM * 16; k += CT) hg[k] = hgeo[(size_t)e * HG_N * 16 + k];       \
  FT md[4], mw[4];                                                                                    \
  _Pragma("unroll") for (int k = 0; k < 4; ++k) { md[k] = cM<FT>(j * 4 + k); mw[k] = cM<FT>(16 + j * 4 + k); } \
  const Lev<FT> L = load_lev(vlev, v, nv);                                                            \
  const int n0 = j * 4;

// ---------------------------------------------------------------------------------------------
template <class FT>
__global__ void __launch_bounds__(CT, (sizeof(FT) == 4 ? 2 : 1))
k2_exp_a(Par<FT> P, const FT* __restrict__ hgeo, const VLev<FT>* __restrict__ vlev, const FT* __restrict__ Yc,
         const FT* __restrict__ Yf, FT* __restrict__ Ytc, FT* __restrict__ Ytf, FT* __restrict__ H) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FT* hg = reinterpret_cast<FT*>(smem_raw);
  FT* sx = hg + HG_ELEM * 16;  // 9 exchange slabs
  FT *s_u3 = sx, *s_r = sx + SLAB, *s_u1 = sx + 2 * SLAB, *s_u2 = sx + 3 * SLAB, *s_U1 = sx + 4 * SLAB, *s_U2 = sx + 5 * SLAB,
     *s_K = sx + 6 * SLAB, *s_X1 = sx + 7 * SLAB, *s_X2 = sx + 8 * SLAB;
  B200_ROW_PROLOGUE
  const bool interior = v > 0 && v < nv;
  const FT* gY = Yc + (size_t)e * P.ncf * 16 * nv;
  FT rho[4], u1[4], u2[4], re[4], u3[4], U1[4], U2[4];
  ld4(rho, gY, nv, j, v, cv, FT(1)); ld4(u1, gY + 16 * nv, nv, j, v, cv, FT(0)); ld4(u2, gY + 32 * nv, nv, j, v, cv, FT(0));
  ld4(re, gY + 48 * nv, nv, j, v, cv, FT(0)); ld4(u3, Yf + (size_t)e * 16 * nf, nf, j, v, fv, FT(0));
  sput(s_u3, u3, j, v); sput(s_r, rho, j, v); sput(s_u1, u1, j, v); sput(s_u2, u2, j, v);
  __syncthreads();  // hg + first exchange slabs
  FT c1[4], c2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    c1[i] = hg[HG_GI11 * 16 + n0 + i] * u1[i] + hg[HG_GI12 * 16 + n0 + i] * u2[i];
    c2[i] = hg[HG_GI12 * 16 + n0 + i] * u1[i] + hg[HG_GI22 * 16 + n0 + i] * u2[i];
    U1[i] = hg[HG_J2 * 16 + n0 + i] * c1[i]; U2[i] = hg[HG_J2 * 16 + n0 + i] * c2[i];
  }
  sput(s_U1, U1, j, v); sput(s_U2, U2, j, v);
  FT K[4], hh[4], ss[4], sd[4], Pi[4], th[4], sE[4], u3c[4];
  FT hs_e[4], hs_d[4];  // Held–Suarez: ρe_tot relaxation and uₕ drag coefficient (held_suarez.jl:111-296)
  {
    FT u3h[4];
    sget(s_u3, u3h, j, v < nv ? v + 1 : v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      K[i] = FT(0.5) * ((u1[i] * c1[i] + u2[i] * c2[i]) * L.sc + FT(0.5) * (u3[i] * (L.g33lo * u3[i]) + u3h[i] * (L.g33hi * u3h[i])));
      Pt<FT> t = thermo(P, rho[i], re[i], K[i], L.phi);
      hh[i] = t.h; Pi[i] = t.Pi; th[i] = t.thp; sE[i] = (K[i] + L.phi) - t.phir;
      sd[i] = P.cp_d * (t.T - P.T_0) + L.phi; ss[i] = sd[i] - t.sdr;
      u3c[i] = FT(0.5) * (u3[i] + u3h[i]);
      hs_e[i] = FT(0); hs_d[i] = FT(0);
      if (P.hs) {
        const FT s2 = hg[HG_SIN2 * 16 + n0 + i], c2 = hg[HG_COS2 * 16 + n0 + i];
        FT hf = fmax_(FT(0), (t.p * P.hs_iMSLP - P.hs_sigb) * P.hs_isig);
        FT Teq = fmax_(P.hs_Tmin, (P.hs_Teq - P.hs_dTy * s2 - P.hs_dthz * (t.lnPi * P.hs_ikap) * c2) * t.Pi);
        FT dRT = (P.hs_ka + (P.hs_ks - P.hs_ka) * hf * c2 * c2) * rho[i] * (t.p / (rho[i] * P.R_d) - Teq);
        hs_e[i] = -dRT * P.cv_d; hs_d[i] = P.hs_kf * hf;
      }
    }
  }
  sput(s_K, K, j, v);
  FT* gT = Ytc + (size_t)e * P.ncf * 16 * nv;
  FT* gH = H ? H + (size_t)e * P.ncf * 16 * nv : nullptr;
  const bool any_visc = P.viscous && __any_sync(FULLM, L.bvc != FT(0));
  // ---- scalars: split-form flux divergences (advection.jl:48,59), viscous sponge on ρe_tot, ∇²s_d
  {
    FT F1[4], F2[4], wd[4], t[4], g2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { F1[i] = rho[i] * U1[i]; F2[i] = rho[i] * U2[i]; }
    div4<FT, 16>(F1, F2, mw, vl, wd);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      wd[i] *= hg[HG_RJ2 * 16 + n0 + i] * L.sc;
      if (cv) gT[(n0 + i) * nv + v] = -wd[i];
    }
    FT G1[4], G2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { G1[i] = F1[i] * hh[i]; G2[i] = F2[i] * hh[i]; }
    div4<FT, 16>(G1, G2, mw, vl, t);
    deta4(hh, md, vl, g2);
    FT et[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      FT rjs = hg[HG_RJ2 * 16 + n0 + i] * L.sc;
      FT g1 = dxi4<FT, 0>(hh, i);
      et[i] = -(FT(0.5) * (t[i] * rjs) + FT(0.5) * (hh[i] * wd[i] + (F1[i] * g1 + F2[i] * g2[i]) * rjs)) + hs_e[i];
    }
    if (any_visc) {  // β wdivₕ(ρ gradₕ s_d)  (viscous_sponge.jl:79)
      FT S1[4], S2[4];
      deta4(sd, md, vl, g2);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        FT g1 = dxi4<FT, 0>(sd, i), rj = rho[i] * hg[HG_J2 * 16 + n0 + i];
        S1[i] = rj * (hg[HG_GI11 * 16 + n0 + i] * g1 + hg[HG_GI12 * 16 + n0 + i] * g2[i]);
        S2[i] = rj * (hg[HG_GI12 * 16 + n0 + i] * g1 + hg[HG_GI22 * 16 + n0 + i] * g2[i]);
      }
      div4<FT, 16>(S1, S2, mw, vl, t);
#pragma unroll
      for (int i = 0; i < 4; ++i) et[i] += L.bvc * (L.sc * t[i] * hg[HG_RJ2 * 16 + n0 + i]);
    }
    if (cv) {
#pragma unroll
      for (int i = 0; i < 4; ++i) gT[(48 + n0 + i) * nv + v] = et[i];
    }
    if (gH) {  // ∇²(s_d − s_d,r)  (hyperdiffusion.jl:142-147)
      FT Q1[4], Q2[4];
      deta4(ss, md, vl, g2);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        FT g1 = dxi4<FT, 0>(ss, i), J2 = hg[HG_J2 * 16 + n0 + i];
        Q1[i] = J2 * (hg[HG_GI11 * 16 + n0 + i] * g1 + hg[HG_GI12 * 16 + n0 + i] * g2[i]);
        Q2[i] = J2 * (hg[HG_GI12 * 16 + n0 + i] * g1 + hg[HG_GI22 * 16 + n0 + i] * g2[i]);
      }
      div4<FT, 16>(Q1, Q2, mw, vl, t);
      if (cv) {
#pragma unroll
        for (int i = 0; i < 4; ++i) gH[(48 + n0 + i) * nv + v] = L.sc * t[i] * hg[HG_RJ2 * 16 + n0 + i];
      }
    }
  }
  // ---- momentum: split-form PGF (advection.jl:82-88)
  FT t1[4], t2[4];
  {
    FT tp[4], gE[4], gP[4], gT2[4], gTP[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) tp[i] = th[i] * Pi[i];
    deta4(sE, md, vl, gE); deta4(Pi, md, vl, gP); deta4(th, md, vl, gT2); deta4(tp, md, vl, gTP);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      t1[i] = -(dxi4<FT, 0>(sE, i) + P.cp_d * (th[i] * dxi4<FT, 0>(Pi, i) + dxi4<FT, 0>(tp, i) - Pi[i] * dxi4<FT, 0>(th, i)) / FT(2));
      t2[i] = -(gE[i] + P.cp_d * (th[i] * gP[i] + gTP[i] - Pi[i] * gT2[i]) / FT(2));
    }
  }
  // ---- ∇²u (hyperdiffusion.jl:141) and viscous sponge on uₕ
  {
    FT D2[4], ze[4], a[4], b[4];
    div4<FT, 0>(U1, U2, md, vl, D2);
    deta4(u1, md, vl, a);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      D2[i] *= hg[HG_RJ2 * 16 + n0 + i];
      ze[i] = (dxi4<FT, 0>(u2, i) - a[i]) * hg[HG_RJ2 * 16 + n0 + i];
    }
    deta4(D2, mw, vl, a); deta4(ze, mw, vl, b);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      FT rJ2 = hg[HG_RJ2 * 16 + n0 + i];
      FT dz1 = dxi4<FT, 16>(ze, i), dz2 = b[i];
      FT L1 = L.sc * (dxi4<FT, 16>(D2, i) - (hg[HG_GC11 * 16 + n0 + i] * dz2 - hg[HG_GC12 * 16 + n0 + i] * dz1) * rJ2);
      FT L2 = L.sc * (a[i] - (hg[HG_GC12 * 16 + n0 + i] * dz2 - hg[HG_GC22 * 16 + n0 + i] * dz1) * rJ2);
      if (gH && cv) { gH[(n0 + i) * nv + v] = L1; gH[(16 + n0 + i) * nv + v] = L2; }
      if (P.viscous) { t1[i] += L.bvc * L1; t2[i] += L.bvc * L2; }
    }
    if (gH) {  // ∇²u₃ = wdivₕ(gradₕ(ᶜinterp(u₃))) on the flat shell
      FT P1[4], P2[4];
      deta4(u3c, md, vl, a);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        FT g1 = dxi4<FT, 0>(u3c, i), J2 = hg[HG_J2 * 16 + n0 + i];
        P1[i] = J2 * (hg[HG_GI11 * 16 + n0 + i] * g1 + hg[HG_GI12 * 16 + n0 + i] * a[i]);
        P2[i] = J2 * (hg[HG_GI12 * 16 + n0 + i] * g1 + hg[HG_GI22 * 16 + n0 + i] * a[i]);
      }
      div4<FT, 16>(P1, P2, mw, vl, b);
      if (cv) {
#pragma unroll
        for (int i = 0; i < 4; ++i) gH[(32 + n0 + i) * nv + v] = L.sc * b[i] * hg[HG_RJ2 * 16 + n0 + i];
      }
    }
    // (ᶜf³ + ᶜω³) × CT12(ᶜu), Rayleigh sponge (advection.jl:228,275-277; remaining_tendency.jl:166)
    deta4(u1, mw, vl, a);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      FT wz = L.sc * (dxi4<FT, 16>(u2, i) - a[i]) * hg[HG_RJ2 * 16 + n0 + i];
      FT tot = hg[HG_COR3 * 16 + n0 + i] + wz;
      t1[i] += tot * U2[i]; t2[i] -= tot * U1[i];
      if (P.rayleigh) { t1[i] -= L.bruh * u1[i]; t2[i] -= L.bruh * u2[i]; }
      if (P.hs) { t1[i] -= hs_d[i] * u1[i]; t2[i] -= hs_d[i] * u2[i]; }
    }
  }
  __syncthreads();  // s_U1, s_U2, s_K complete
  // ---- face level v: ᶠω¹², mass flux, u₃ tendency (advection.jl:233-237,273-278)
  FT X1[4], X2[4];
  {
    FT d3[4], rl[4], a1[4], a2[4], b1[4], b2[4], kl[4], lap[4];
    deta4(u3, mw, vl, d3);
    const int vm = v > 0 ? v - 1 : 0;
    sget(s_r, rl, j, vm); sget(s_u1, a1, j, vm); sget(s_u2, a2, j, vm); sget(s_U1, b1, j, vm); sget(s_U2, b2, j, vm); sget(s_K, kl, j, vm);
    const bool any_v3 = P.viscous && __any_sync(FULLM, L.bvf != FT(0));
    if (any_v3) {  // β wdivₕ(gradₕ u₃) on faces (viscous_sponge.jl:64)
      FT R1[4], R2[4], g2[4];
      deta4(u3, md, vl, g2);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        FT g1 = dxi4<FT, 0>(u3, i), J2 = hg[HG_J2 * 16 + n0 + i];
        R1[i] = J2 * (hg[HG_GI11 * 16 + n0 + i] * g1 + hg[HG_GI12 * 16 + n0 + i] * g2[i]);
        R2[i] = J2 * (hg[HG_GI12 * 16 + n0 + i] * g1 + hg[HG_GI22 * 16 + n0 + i] * g2[i]);
      }
      div4<FT, 16>(R1, R2, mw, vl, lap);
    }
    FT* gF = Ytf + (size_t)e * 16 * nf;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const FT J2 = hg[HG_J2 * 16 + n0 + i], rJ2 = hg[HG_RJ2 * 16 + n0 + i];
      FT jt1 = J2 * L.sf * L.dzf * hg[HG_COR1 * 16 + n0 + i] + d3[i];
      FT jt2 = J2 * L.sf * L.dzf * hg[HG_COR2 * 16 + n0 + i] - dxi4<FT, 16>(u3, i);
      FT Vn, ub1, ub2, dk = FT(0);
      if (interior) {
        jt1 -= (u2[i] - a2[i]); jt2 += (u1[i] - a1[i]);
        Vn = FT(0.5) * (rl[i] * L.mclo + rho[i] * L.mc);
        ub1 = FT(0.5) * (b1[i] * L.sclo + U1[i] * L.sc) * rJ2;
        ub2 = FT(0.5) * (b2[i] * L.sclo + U2[i] * L.sc) * rJ2;
        dk = K[i] - kl[i];
      } else if (v == 0) {
        Vn = rho[i] * L.mc; ub1 = U1[i] * L.sc * rJ2; ub2 = U2[i] * L.sc * rJ2;
      } else {
        Vn = rl[i] * L.mclo; ub1 = b1[i] * L.sclo * rJ2; ub2 = b2[i] * L.sclo * rJ2;
      }
      Vn *= L.g33lo * u3[i];
      X1[i] = jt2 * Vn; X2[i] = -jt1 * Vn;
      FT t3 = -(jt1 * ub2 - jt2 * ub1) - dk;
      if (any_v3) t3 += L.bvf * (L.sf2i * lap[i] * rJ2);
      if (fv) gF[(n0 + i) * nf + v] = t3;
    }
  }
  sput(s_X1, X1, j, v); sput(s_X2, X2, j, v);
  __syncthreads();
  if (cv) {
    FT h1[4], h2[4];
    sget(s_X1, h1, j, v + 1); sget(s_X2, h2, j, v + 1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      FT irm = FT(0.5) * L.rmc * rcp_(rho[i]);
      gT[(16 + n0 + i) * nv + v] = t1[i] - (X1[i] + h1[i]) * irm;
      gT[(32 + n0 + i) * nv + v] = t2[i] - (X2[i] + h2[i]) * irm;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// The three outputs (uₕ, ρe_tot, u₃) are independent given the DSSed ∇² fields: blockIdx.y selects one, which
// triples the number of CTAs and cuts registers per thread (the single-kernel version was memory-latency
// bound: long-scoreboard 5.0 stall cycles per issue at 24 warps/SM, profiles/r1_ncu_summary.md).
template <class FT>
__global__ void __launch_bounds__(CT, (sizeof(FT) == 4 ? 4 : 2))
k2_exp_c(Par<FT> P, const FT* __restrict__ hgeo, const VLev<FT>* __restrict__ vlev, const FT* __restrict__ Yc,
         const FT* __restrict__ H, FT* __restrict__ Ytc, FT* __restrict__ Ytf) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FT* hg = reinterpret_cast<FT*>(smem_raw);
  FT* s_w = hg + HG_ELEM * 16;
  FT* s_a = s_w + SLAB;
  B200_ROW_PROLOGUE
  const int part = blockIdx.y;
  const FT* gH = H + (size_t)e * P.ncf * 16 * nv;
  FT* gT = Ytc + (size_t)e * P.ncf * 16 * nv;
  FT* gF = Ytf + (size_t)e * 16 * nf;
  FT a[4], b[4];
  if (part == 0) {  // ∇⁴uₕ = δ_div·wgradₕ(divₕ(∇²u)) − wcurlₕ(curlₕ(∇²u))  (hyperdiffusion.jl:273-276)
    FT L1[4], L2[4], old1[4], old2[4];
    ld4(L1, gH, nv, j, v, cv, FT(0)); ld4(L2, gH + 16 * nv, nv, j, v, cv, FT(0));
    ld4(old1, gT + 16 * nv, nv, j, v, cv, FT(0)); ld4(old2, gT + 32 * nv, nv, j, v, cv, FT(0));
    __syncthreads();
    FT U1[4], U2[4], D2[4], ze[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      FT J2 = hg[HG_J2 * 16 + n0 + i];
      U1[i] = J2 * (hg[HG_GI11 * 16 + n0 + i] * L1[i] + hg[HG_GI12 * 16 + n0 + i] * L2[i]);
      U2[i] = J2 * (hg[HG_GI12 * 16 + n0 + i] * L1[i] + hg[HG_GI22 * 16 + n0 + i] * L2[i]);
    }
    div4<FT, 0>(U1, U2, md, vl, D2);
    deta4(L1, md, vl, a);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      D2[i] *= hg[HG_RJ2 * 16 + n0 + i];
      ze[i] = (dxi4<FT, 0>(L2, i) - a[i]) * hg[HG_RJ2 * 16 + n0 + i];
    }
    deta4(D2, mw, vl, a); deta4(ze, mw, vl, b);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      FT rJ2 = hg[HG_RJ2 * 16 + n0 + i];
      FT dz1 = dxi4<FT, 16>(ze, i), dz2 = b[i];
      FT Qa = L.sc * (P.ddf * dxi4<FT, 16>(D2, i) - (hg[HG_GC11 * 16 + n0 + i] * dz2 - hg[HG_GC12 * 16 + n0 + i] * dz1) * rJ2);
      FT Qb = L.sc * (P.ddf * a[i] - (hg[HG_GC12 * 16 + n0 + i] * dz2 - hg[HG_GC22 * 16 + n0 + i] * dz1) * rJ2);
      if (cv) { gT[(16 + n0 + i) * nv + v] = old1[i] - P.nu4v * Qa; gT[(32 + n0 + i) * nv + v] = old2[i] - P.nu4v * Qb; }
    }
  } else if (part == 1) {  // Yₜ.ρe_tot −= ν₄ₛ wdivₕ(ρ gradₕ(∇²s_d))  (hyperdiffusion.jl:291,307)
    FT rho[4], Ls[4], old3[4], Q1[4], Q2[4];
    ld4(rho, Yc + (size_t)e * P.ncf * 16 * nv, nv, j, v, cv, FT(1));
    ld4(Ls, gH + 48 * nv, nv, j, v, cv, FT(0));
    ld4(old3, gT + 48 * nv, nv, j, v, cv, FT(0));
    __syncthreads();
    deta4(Ls, md, vl, a);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      FT g1 = dxi4<FT, 0>(Ls, i), rj = rho[i] * hg[HG_J2 * 16 + n0 + i];
      Q1[i] = rj * (hg[HG_GI11 * 16 + n0 + i] * g1 + hg[HG_GI12 * 16 + n0 + i] * a[i]);
      Q2[i] = rj * (hg[HG_GI12 * 16 + n0 + i] * g1 + hg[HG_GI22 * 16 + n0 + i] * a[i]);
    }
    div4<FT, 16>(Q1, Q2, mw, vl, b);
    if (cv) {
#pragma unroll
      for (int i = 0; i < 4; ++i) gT[(48 + n0 + i) * nv + v] = old3[i] - P.nu4s * (L.sc * b[i] * hg[HG_RJ2 * 16 + n0 + i]);
    }
  } else {  // Yₜ.f.u₃ −= ν₄ᵥ ᶠwinterp(ᶜJ ρ, C3(∇⁴u))  (hyperdiffusion.jl:277)
    FT rho[4], L3[4], oldf[4], P1[4], P2[4], q[4], w[4];
    ld4(rho, Yc + (size_t)e * P.ncf * 16 * nv, nv, j, v, cv, FT(1));
    ld4(L3, gH + 32 * nv, nv, j, v, cv, FT(0));
    ld4(oldf, gF, nf, j, v, fv, FT(0));
    __syncthreads();
    deta4(L3, md, vl, a);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      FT g1 = dxi4<FT, 0>(L3, i), J2 = hg[HG_J2 * 16 + n0 + i];
      P1[i] = J2 * (hg[HG_GI11 * 16 + n0 + i] * g1 + hg[HG_GI12 * 16 + n0 + i] * a[i]);
      P2[i] = J2 * (hg[HG_GI12 * 16 + n0 + i] * g1 + hg[HG_GI22 * 16 + n0 + i] * a[i]);
    }
    div4<FT, 16>(P1, P2, mw, vl, b);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      q[i] = L.sc * b[i] * hg[HG_RJ2 * 16 + n0 + i];
      w[i] = L.mc * rho[i];
    }
    sput(s_w, w, j, v); sput(s_a, q, j, v);
    __syncthreads();
    if (fv) {
      FT wl[4], ql[4];
      const int vm = v > 0 ? v - 1 : 0;
      sget(s_w, wl, j, vm); sget(s_a, ql, j, vm);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        FT val = (v == 0) ? q[i] : (v == nv ? ql[i] : (wl[i] * ql[i] + w[i] * q[i]) / (wl[i] + w[i]));
        gF[(n0 + i) * nf + v] = oldf[i] - P.nu4v * val;
      }
    }
  }
}

}  // namespace b200

namespace b200 {

// ---------------------------------------------------------------------------------------------
// k2_imp_stage — fused implicit stage, second generation (out of place: U → N).
//
// Same arithmetic as k_imp_stage (cache_imp! → Wfact → T_imp! residual → ldiv! → U −= ΔU → cache_imp!
// → T_post_imp!), restructured for the B200 after the first ncu profile (24 % warp occupancy, 213 M
// instructions, 16 of 256 threads active in the Thomas sweep):
//   * 11 shared slabs instead of 18 (coefficients live in registers of the (node, level) owner and the
//     solver slabs alias the dead thermodynamic slabs) ⇒ 4 CTAs/SM so other CTAs cover the Thomas sweep;
//   * thermodynamics evaluated once with transcendental functions (Π, Φ_r) and once without (only
//     h_tot is needed after the Newton update);
//   * the Schur tridiagonal is assembled from one per-face coefficient A = dtγ·ᶠinterp(ρJ)g³³/J2
//     instead of calling the centre-row routine twice per face;
//   * reciprocal-based Thomas sweep (one division per row).
// The u₃ boundary filter of cache_imp! is applied on load, so the input may carry unfiltered boundary
// values; uₕ is copied through (its Newton increment is identically zero because R_uₕ = dtγ·0).
template <class FT>
__global__ void __launch_bounds__(NT, (sizeof(FT) == 4 ? 4 : 2))
k2_imp_stage(Par<FT> P, const FT* __restrict__ hgeo, const VLev<FT>* __restrict__ vlev, const FT* __restrict__ Yc,
             const FT* __restrict__ Yf, FT* __restrict__ Nc, FT* __restrict__ Nf, FT dtg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<FT> sm(smem_raw);
  VLev<FT>& V = *reinterpret_cast<VLev<FT>*>(sm.take(sizeof(VLev<FT>) / sizeof(FT)));
  FT* hg = sm.take(HG_ELEM * 16);
  FT *s_rho = sm.take(SLAB), *s_u3 = sm.take(SLAB), *s_h = sm.take(SLAB), *s_Kh = sm.take(SLAB), *s_M = sm.take(SLAB),
     *s_A = sm.take(SLAB);
  FT *s_Pi = sm.take(SLAB), *s_thv = sm.take(SLAB), *s_thp = sm.take(SLAB), *s_phr = sm.take(SLAB), *s_dp = sm.take(SLAB);
  FT *s_l = s_Pi, *s_d = s_thv, *s_u = s_thp, *s_r = s_phr;  // solver slabs alias dead thermodynamic slabs
  const int e = blockIdx.x, nv = P.nv, nf = nv + 1;
  const FT kap = P.R_d / P.cv_d;
  load_vlev(&V, vlev);
  load_hgeo(hg, hgeo, e);
  const FT* gY = Yc + (size_t)e * P.ncf * 16 * nv;
  const FT* gYf = Yf + (size_t)e * 16 * nf;
  FT* gN = Nc + (size_t)e * P.ncf * 16 * nv;
  FT* gNf = Nf + (size_t)e * 16 * nf;
  FT r_re[NIT], r_u1[NIT], r_u2[NIT];
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int idx = threadIdx.x + it * NT, n = idx >> 6, v = idx & 63, o = n * LVP + v;
    r_re[it] = r_u1[it] = r_u2[it] = FT(0);
    if (v < nv) {
      s_rho[o] = gY[n * nv + v];
      r_u1[it] = gY[(16 + n) * nv + v]; r_u2[it] = gY[(32 + n) * nv + v]; r_re[it] = gY[(48 + n) * nv + v];
      gN[(16 + n) * nv + v] = r_u1[it]; gN[(32 + n) * nv + v] = r_u2[it];
      for (int q = 4; q < P.ncf; ++q) gN[(q * 16 + n) * nv + v] = gY[(q * 16 + n) * nv + v];  // passive tracers: ΔU = 0
    }
    if (v < nf) s_u3[o] = (v == 0 || v == nv) ? FT(0) : gYf[n * nf + v];
  }
  __syncthreads();
  // ---- phase 1: centre thermodynamics
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int idx = threadIdx.x + it * NT, n = idx >> 6, v = idx & 63, o = n * LVP + v;
    if (v < nv) {
      FT a1 = r_u1[it], a2 = r_u2[it];
      FT c1 = hg[HG_GI11 * 16 + n] * a1 + hg[HG_GI12 * 16 + n] * a2;
      FT c2 = hg[HG_GI12 * 16 + n] * a1 + hg[HG_GI22 * 16 + n] * a2;
      FT Kh = FT(0.5) * ((a1 * c1 + a2 * c2) * V.sc2i[v]);
      FT lo = s_u3[o], hi = s_u3[o + 1];
      FT K = Kh + FT(0.25) * (lo * (V.g33f[v] * lo) + hi * (V.g33f[v + 1] * hi));
      Pt<FT> t = thermo(P, s_rho[o], r_re[it], K, V.phic[v]);
      s_Kh[o] = Kh; s_h[o] = t.h; s_Pi[o] = t.Pi; s_thv[o] = t.thv; s_thp[o] = t.thp; s_phr[o] = t.phir;
      s_dp[o] = kap * (P.T_0 * P.cp_d - K - V.phic[v]) + (P.R_d - kap * P.cv_d) * t.T;
    }
  }
  __syncthreads();
  // ---- phase 2: face mass-flux pieces  M = ᶠinterp(ρJ)u³/J2,  A = dtγ ᶠinterp(ρJ) g³³/J2 (zero on boundaries)
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int idx = threadIdx.x + it * NT, n = idx >> 6, f = idx & 63, o = n * LVP + f;
    if (f < nf) {
      FT M = FT(0), A = FT(0);
      if (f > 0 && f < nv) {
        FT mr = rho_mface(V, s_rho, o, f);
        A = dtg * mr * V.g33f[f];
        M = mr * (V.g33f[f] * s_u3[o]);
      }
      s_M[o] = M; s_A[o] = A;
    }
  }
  __syncthreads();
  // ---- phase 3: Schur tridiagonal and right-hand side of face row f (manual_sparse_jacobian.jl:746-868)
  FT cl[NIT], cd[NIT], cu[NIT], cr[NIT];
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int idx = threadIdx.x + it * NT, n = idx >> 6, f = idx & 63, o = n * LVP + f;
    cl[it] = cu[it] = cr[it] = FT(0); cd[it] = FT(-1);
    if (f < nf) {
      FT beta = P.rayleigh ? V.brw[f] : FT(0);
      cd[it] = dtg * (-beta) - FT(1);
      if (f > 0 && f < nv) {
        FT rlo = s_rho[o - 1], rhi = s_rho[o];
        FT irf = rcp_(FT(0.5) * (rlo + rhi));
        FT dPi = s_Pi[o] - s_Pi[o - 1];
        FT buoy = P.cp_d * (FT(0.5) * (s_thv[o - 1] + s_thv[o])) * dPi * irf;
        FT ur_lo = dtg * (irf * s_dp[o - 1] + buoy * FT(0.5)), ur_hi = dtg * (-irf * s_dp[o] + buoy * FT(0.5));
        FT ue_lo = dtg * irf * kap, ue_hi = -ue_lo;
        FT x_lo = irf * (-kap * rlo), x_hi = -irf * (-kap * rhi);
        FT k0 = FT(0.5) * V.g33f[f] * s_u3[o];
        FT l = dtg * (x_lo * (FT(0.5) * V.g33f[f - 1] * s_u3[o - 1]));
        FT d = dtg * ((x_lo * k0 + x_hi * k0) - beta) - FT(1);
        FT u = dtg * (x_hi * (FT(0.5) * V.g33f[f + 1] * s_u3[o + 1]));
        // centre rows f-1 ("a") and f ("b"): ru_lo = A[k]/m_c[k], ru_hi = −A[k+1]/m_c[k], eu = ru·ᶠinterp(h)
        FT ima = V.rmc[f - 1], imb = V.rmc[f];
        FT Am = s_A[o - 1], A0 = s_A[o], Ap = s_A[o + 1];
        FT hm = (f > 1) ? FT(0.5) * (s_h[o - 2] + s_h[o - 1]) : FT(0);
        FT h0 = FT(0.5) * (s_h[o - 1] + s_h[o]);
        FT hp = (f < nv - 1) ? FT(0.5) * (s_h[o] + s_h[o + 1]) : FT(0);
        FT ru_lo_a = Am * ima, ru_hi_a = -A0 * ima, ru_lo_b = A0 * imb, ru_hi_b = -Ap * imb;
        l += ur_lo * ru_lo_a + ue_lo * (ru_lo_a * hm);
        d += ur_lo * ru_hi_a + ur_hi * ru_lo_b + ue_lo * (ru_hi_a * h0) + ue_hi * (ru_lo_b * h0);
        u += ur_hi * ru_hi_b + ue_hi * (ru_hi_b * hp);
        // R = dtγ·T_imp(U): face part + couplings to the centre residuals of rows f-1 and f
        FT Mm = s_M[o - 1], M0 = s_M[o], Mp = s_M[o + 1];
        FT rr_a = -dtg * (M0 - Mm) * ima, rr_b = -dtg * (Mp - M0) * imb;
        FT re_a = -dtg * (M0 * h0 - Mm * hm) * ima, re_b = -dtg * (Mp * hp - M0 * h0) * imb;
        FT tf = -(V.dphif[f] - (s_phr[o] - s_phr[o - 1]) + P.cp_d * (FT(0.5) * (s_thp[o - 1] + s_thp[o])) * dPi) - beta * s_u3[o];
        cl[it] = l; cd[it] = d; cu[it] = u;
        cr[it] = dtg * tf + ur_lo * rr_a + ur_hi * rr_b + ue_lo * re_a + ue_hi * re_b;
      }
    }
  }
  __syncthreads();  // all reads of the thermodynamic slabs are done: reuse them for the solver
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int idx = threadIdx.x + it * NT, n = idx >> 6, f = idx & 63, o = n * LVP + f;
    if (f < nf) { s_l[o] = cl[it]; s_d[o] = cd[it]; s_u[o] = cu[it]; s_r[o] = cr[it]; }
  }
  __syncthreads();
  if (threadIdx.x < 16) {  // Thomas sweep, one column per thread (BlockArrowheadSolve → Thomas)
    const FT *l = s_l + threadIdx.x * LVP, *d = s_d + threadIdx.x * LVP;
    FT *u = s_u + threadIdx.x * LVP, *r = s_r + threadIdx.x * LVP;
    FT rd = rcp_(d[0]);
    FT cp = u[0] * rd, dp = r[0] * rd;
    u[0] = cp; r[0] = dp;
    for (int i = 1; i < nf; ++i) {
      FT li = l[i];
      rd = rcp_(d[i] - li * cp);
      cp = u[i] * rd;
      dp = (r[i] - li * dp) * rd;
      u[i] = cp; r[i] = dp;
    }
    FT x = dp;
    for (int i = nf - 2; i >= 0; --i) { x = r[i] - u[i] * x; r[i] = x; }
  }
  __syncthreads();
  // ---- phase 5: U ← U − ΔU (back-substitution of the scalar rows)
  FT n_re[NIT];
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int idx = threadIdx.x + it * NT, n = idx >> 6, v = idx & 63, o = n * LVP + v;
    n_re[it] = FT(0);
    FT nr = FT(0), nu = FT(0);
    if (v < nv) {
      FT im = V.rmc[v];
      FT A0 = s_A[o], Ap = s_A[o + 1], M0 = s_M[o], Mp = s_M[o + 1];
      FT h0 = (v > 0) ? FT(0.5) * (s_h[o - 1] + s_h[o]) : FT(0);
      FT hp = (v < nv - 1) ? FT(0.5) * (s_h[o] + s_h[o + 1]) : FT(0);
      FT x0 = s_r[o], x1 = s_r[o + 1];
      FT rr = -dtg * (Mp - M0) * im, rre = -dtg * (Mp * hp - M0 * h0) * im;
      nr = s_rho[o] - ((A0 * im) * x0 + (-Ap * im) * x1 - rr);
      n_re[it] = r_re[it] - ((A0 * im * h0) * x0 + (-Ap * im * hp) * x1 - rre);
    }
    if (v < nf) nu = (v == 0 || v == nv) ? FT(0) : s_u3[o] - s_r[o];
    // (only own entries of s_rho/s_u3 are read in this phase, so they can be updated in place)
    if (v < nv) { s_rho[o] = nr; gN[n * nv + v] = nr; }
    if (v < nf) { s_u3[o] = nu; gNf[n * nf + v] = nu; }
  }
  __syncthreads();
  if (P.upwinding != 0) {
    // ---- phase 6: h_tot of the updated state (cache_imp! after the Newton update; no transcendentals needed)
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int idx = threadIdx.x + it * NT, n = idx >> 6, v = idx & 63, o = n * LVP + v;
      if (v < nv) {
        FT lo = s_u3[o], hi = s_u3[o + 1];
        FT K = s_Kh[o] + FT(0.25) * (lo * (V.g33f[v] * lo) + hi * (V.g33f[v + 1] * hi));
        FT etot = n_re[it] * rcp_(s_rho[o]);
        FT T = fmax_(P.T_min_sgs, P.T_0 + ((etot - K - V.phic[v]) + P.RT0) * P.icv);
        s_h[o] = etot + P.R_d * T;
      }
    }
    __syncthreads();
    // ---- phase 7: (upwinded − centred) enthalpy flux (implicit_tendency.jl:322-339)
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int idx = threadIdx.x + it * NT, n = idx >> 6, f = idx & 63, o = n * LVP + f;
      if (f < nf) {
        FT r = FT(0);
        if (f > 0 && f < nv) {
          FT w = V.g33f[f] * s_u3[o];
          r = rho_mface(V, s_rho, o, f) * w * upwind_minus_central(P, s_h, o, f, nv, w);
        }
        s_M[o] = r;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int idx = threadIdx.x + it * NT, n = idx >> 6, v = idx & 63, o = n * LVP + v;
    if (v < nv) {
      FT e2 = n_re[it];
      if (P.upwinding != 0) e2 += dtg * (-(s_M[o + 1] - s_M[o]) * V.rmc[v]);
      gN[(48 + n) * nv + v] = e2;
    }
  }
}

// k4_imp_stage — the same kernel on a QUARTER element: CTA = 64 threads = 4 columns (one GLL row), so every
// barrier only joins two warps and 16 CTAs are resident per SM; level constants and metric terms are read
// straight from global/L1 instead of being staged per CTA.
constexpr int QT = 64;
constexpr int QSLAB = 4 * LVP;
template <class FT>
__global__ void __launch_bounds__(QT, (sizeof(FT) == 4 ? 16 : 8))
k4_imp_stage(Par<FT> P, const FT* __restrict__ hgeo, const VLev<FT>* __restrict__ vlev, const FT* __restrict__ Yc,
             const FT* __restrict__ Yf, FT* __restrict__ Nc, FT* __restrict__ Nf, FT dtg) {
  __shared__ FT slab[11][QSLAB];
  const VLev<FT>& V = *vlev;  // level constants straight from global/L1 (each thread owns one level)
  FT *s_rho = slab[0], *s_u3 = slab[1], *s_h = slab[2], *s_Kh = slab[3], *s_M = slab[4], *s_A = slab[5];
  FT *s_Pi = slab[6], *s_thv = slab[7], *s_thp = slab[8], *s_phr = slab[9], *s_dp = slab[10];
  FT *s_l = s_Pi, *s_d = s_thv, *s_u = s_thp, *s_r = s_phr;  // solver slabs alias dead thermodynamic slabs
  const int e = blockIdx.x >> 2, nq0 = (blockIdx.x & 3) * 4, nv = P.nv, nf = nv + 1;
  const FT* hg = hgeo + (size_t)e * HG_N * 16;
  const FT kap = P.R_d / P.cv_d;
  const FT* gY = Yc + (size_t)e * P.ncf * 16 * nv;
  const FT* gYf = Yf + (size_t)e * 16 * nf;
  FT* gN = Nc + (size_t)e * P.ncf * 16 * nv;
  FT* gNf = Nf + (size_t)e * 16 * nf;
  FT r_re[NIT], r_u1[NIT], r_u2[NIT];
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int n = nq0 + it, v = threadIdx.x, o = it * LVP + v;
    r_re[it] = r_u1[it] = r_u2[it] = FT(0);
    if (v < nv) {
      s_rho[o] = gY[n * nv + v];
      r_u1[it] = gY[(16 + n) * nv + v]; r_u2[it] = gY[(32 + n) * nv + v]; r_re[it] = gY[(48 + n) * nv + v];
      gN[(16 + n) * nv + v] = r_u1[it]; gN[(32 + n) * nv + v] = r_u2[it];
    }
    if (v < nf) s_u3[o] = (v == 0 || v == nv) ? FT(0) : gYf[n * nf + v];
  }
  __syncthreads();
  // ---- phase 1: centre thermodynamics
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int n = nq0 + it, v = threadIdx.x, o = it * LVP + v;
    if (v < nv) {
      FT a1 = r_u1[it], a2 = r_u2[it];
      FT c1 = hg[HG_GI11 * 16 + n] * a1 + hg[HG_GI12 * 16 + n] * a2;
      FT c2 = hg[HG_GI12 * 16 + n] * a1 + hg[HG_GI22 * 16 + n] * a2;
      FT Kh = FT(0.5) * ((a1 * c1 + a2 * c2) * V.sc2i[v]);
      FT lo = s_u3[o], hi = s_u3[o + 1];
      FT K = Kh + FT(0.25) * (lo * (V.g33f[v] * lo) + hi * (V.g33f[v + 1] * hi));
      Pt<FT> t = thermo(P, s_rho[o], r_re[it], K, V.phic[v]);
      s_Kh[o] = Kh; s_h[o] = t.h; s_Pi[o] = t.Pi; s_thv[o] = t.thv; s_thp[o] = t.thp; s_phr[o] = t.phir;
      s_dp[o] = kap * (P.T_0 * P.cp_d - K - V.phic[v]) + (P.R_d - kap * P.cv_d) * t.T;
    }
  }
  __syncthreads();
  // ---- phase 2: face mass-flux pieces  M = ᶠinterp(ρJ)u³/J2,  A = dtγ ᶠinterp(ρJ) g³³/J2 (zero on boundaries)
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int n = nq0 + it, f = threadIdx.x, o = it * LVP + f; (void)n;
    if (f < nf) {
      FT M = FT(0), A = FT(0);
      if (f > 0 && f < nv) {
        FT mr = rho_mface(V, s_rho, o, f);
        A = dtg * mr * V.g33f[f];
        M = mr * (V.g33f[f] * s_u3[o]);
      }
      s_M[o] = M; s_A[o] = A;
    }
  }
  __syncthreads();
  // ---- phase 3: Schur tridiagonal and right-hand side of face row f (manual_sparse_jacobian.jl:746-868)
  FT cl[NIT], cd[NIT], cu[NIT], cr[NIT];
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int n = nq0 + it, f = threadIdx.x, o = it * LVP + f; (void)n;
    cl[it] = cu[it] = cr[it] = FT(0); cd[it] = FT(-1);
    if (f < nf) {
      FT beta = P.rayleigh ? V.brw[f] : FT(0);
      cd[it] = dtg * (-beta) - FT(1);
      if (f > 0 && f < nv) {
        FT rlo = s_rho[o - 1], rhi = s_rho[o];
        FT irf = rcp_(FT(0.5) * (rlo + rhi));
        FT dPi = s_Pi[o] - s_Pi[o - 1];
        FT buoy = P.cp_d * (FT(0.5) * (s_thv[o - 1] + s_thv[o])) * dPi * irf;
        FT ur_lo = dtg * (irf * s_dp[o - 1] + buoy * FT(0.5)), ur_hi = dtg * (-irf * s_dp[o] + buoy * FT(0.5));
        FT ue_lo = dtg * irf * kap, ue_hi = -ue_lo;
        FT x_lo = irf * (-kap * rlo), x_hi = -irf * (-kap * rhi);
        FT k0 = FT(0.5) * V.g33f[f] * s_u3[o];
        FT l = dtg * (x_lo * (FT(0.5) * V.g33f[f - 1] * s_u3[o - 1]));
        FT d = dtg * ((x_lo * k0 + x_hi * k0) - beta) - FT(1);
        FT u = dtg * (x_hi * (FT(0.5) * V.g33f[f + 1] * s_u3[o + 1]));
        // centre rows f-1 ("a") and f ("b"): ru_lo = A[k]/m_c[k], ru_hi = −A[k+1]/m_c[k], eu = ru·ᶠinterp(h)
        FT ima = V.rmc[f - 1], imb = V.rmc[f];
        FT Am = s_A[o - 1], A0 = s_A[o], Ap = s_A[o + 1];
        FT hm = (f > 1) ? FT(0.5) * (s_h[o - 2] + s_h[o - 1]) : FT(0);
        FT h0 = FT(0.5) * (s_h[o - 1] + s_h[o]);
        FT hp = (f < nv - 1) ? FT(0.5) * (s_h[o] + s_h[o + 1]) : FT(0);
        FT ru_lo_a = Am * ima, ru_hi_a = -A0 * ima, ru_lo_b = A0 * imb, ru_hi_b = -Ap * imb;
        l += ur_lo * ru_lo_a + ue_lo * (ru_lo_a * hm);
        d += ur_lo * ru_hi_a + ur_hi * ru_lo_b + ue_lo * (ru_hi_a * h0) + ue_hi * (ru_lo_b * h0);
        u += ur_hi * ru_hi_b + ue_hi * (ru_hi_b * hp);
        // R = dtγ·T_imp(U): face part + couplings to the centre residuals of rows f-1 and f
        FT Mm = s_M[o - 1], M0 = s_M[o], Mp = s_M[o + 1];
        FT rr_a = -dtg * (M0 - Mm) * ima, rr_b = -dtg * (Mp - M0) * imb;
        FT re_a = -dtg * (M0 * h0 - Mm * hm) * ima, re_b = -dtg * (Mp * hp - M0 * h0) * imb;
        FT tf = -(V.dphif[f] - (s_phr[o] - s_phr[o - 1]) + P.cp_d * (FT(0.5) * (s_thp[o - 1] + s_thp[o])) * dPi) - beta * s_u3[o];
        cl[it] = l; cd[it] = d; cu[it] = u;
        cr[it] = dtg * tf + ur_lo * rr_a + ur_hi * rr_b + ue_lo * re_a + ue_hi * re_b;
      }
    }
  }
  __syncthreads();  // all reads of the thermodynamic slabs are done: reuse them for the solver
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int n = nq0 + it, f = threadIdx.x, o = it * LVP + f; (void)n;
    if (f < nf) { s_l[o] = cl[it]; s_d[o] = cd[it]; s_u[o] = cu[it]; s_r[o] = cr[it]; }
  }
  __syncthreads();
  if (threadIdx.x < 4) {  // Thomas sweep, one column per thread (BlockArrowheadSolve → Thomas)
    const FT *l = s_l + threadIdx.x * LVP, *d = s_d + threadIdx.x * LVP;
    FT *u = s_u + threadIdx.x * LVP, *r = s_r + threadIdx.x * LVP;
    FT rd = rcp_(d[0]);
    FT cp = u[0] * rd, dp = r[0] * rd;
    u[0] = cp; r[0] = dp;
    for (int i = 1; i < nf; ++i) {
      FT li = l[i];
      rd = rcp_(d[i] - li * cp);
      cp = u[i] * rd;
      dp = (r[i] - li * dp) * rd;
      u[i] = cp; r[i] = dp;
    }
    FT x = dp;
    for (int i = nf - 2; i >= 0; --i) { x = r[i] - u[i] * x; r[i] = x; }
  }
  __syncthreads();
  // ---- phase 5: U ← U − ΔU (back-substitution of the scalar rows)
  FT n_re[NIT];
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int n = nq0 + it, v = threadIdx.x, o = it * LVP + v;
    n_re[it] = FT(0);
    FT nr = FT(0), nu = FT(0);
    if (v < nv) {
      FT im = V.rmc[v];
      FT A0 = s_A[o], Ap = s_A[o + 1], M0 = s_M[o], Mp = s_M[o + 1];
      FT h0 = (v > 0) ? FT(0.5) * (s_h[o - 1] + s_h[o]) : FT(0);
      FT hp = (v < nv - 1) ? FT(0.5) * (s_h[o] + s_h[o + 1]) : FT(0);
      FT x0 = s_r[o], x1 = s_r[o + 1];
      FT rr = -dtg * (Mp - M0) * im, rre = -dtg * (Mp * hp - M0 * h0) * im;
      nr = s_rho[o] - ((A0 * im) * x0 + (-Ap * im) * x1 - rr);
      n_re[it] = r_re[it] - ((A0 * im * h0) * x0 + (-Ap * im * hp) * x1 - rre);
    }
    if (v < nf) nu = (v == 0 || v == nv) ? FT(0) : s_u3[o] - s_r[o];
    // (only own entries of s_rho/s_u3 are read in this phase, so they can be updated in place)
    if (v < nv) { s_rho[o] = nr; gN[n * nv + v] = nr; }
    if (v < nf) { s_u3[o] = nu; gNf[n * nf + v] = nu; }
  }
  __syncthreads();
  if (P.upwinding != 0) {
    // ---- phase 6: h_tot of the updated state (cache_imp! after the Newton update; no transcendentals needed)
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int n = nq0 + it, v = threadIdx.x, o = it * LVP + v;
      if (v < nv) {
        FT lo = s_u3[o], hi = s_u3[o + 1];
        FT K = s_Kh[o] + FT(0.25) * (lo * (V.g33f[v] * lo) + hi * (V.g33f[v + 1] * hi));
        FT etot = n_re[it] * rcp_(s_rho[o]);
        FT T = fmax_(P.T_min_sgs, P.T_0 + ((etot - K - V.phic[v]) + P.RT0) * P.icv);
        s_h[o] = etot + P.R_d * T;
      }
    }
    __syncthreads();
    // ---- phase 7: (upwinded − centred) enthalpy flux (implicit_tendency.jl:322-339)
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int n = nq0 + it, f = threadIdx.x, o = it * LVP + f; (void)n;
      if (f < nf) {
        FT r = FT(0);
        if (f > 0 && f < nv) {
          FT w = V.g33f[f] * s_u3[o];
          r = rho_mface(V, s_rho, o, f) * w * upwind_minus_central(P, s_h, o, f, nv, w);
        }
        s_M[o] = r;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int n = nq0 + it, v = threadIdx.x, o = it * LVP + v;
    if (v < nv) {
      FT e2 = n_re[it];
      if (P.upwinding != 0) e2 += dtg * (-(s_M[o + 1] - s_M[o]) * V.rmc[v]);
      gN[(48 + n) * nv + v] = e2;
    }
  }
}


}  // namespace b200

namespace b200 {

// ---------------------------------------------------------------------------------------------
// k3_imp_stage — fused implicit stage, third generation: ONE WARP PER COLUMN.
//
// Lane l owns levels l and l+32 of its column in registers; every vertical neighbour is a warp shuffle, the
// Thomas sweep runs on lane 0 over a 1 KB warp-private shared buffer between two __syncwarp()s, and there is
// NO block-level barrier at all (the second generation spent 3.8 stall cycles per issue at its 8
// __syncthreads, profiles/r1_ncu_summary.md).  CTA = 8 warps = 8 columns; grid = columns/8.  Same arithmetic
// as k2_imp_stage.  Column c of element e, node n: centre data at ((e·4+f)·16+n)·Nv, faces at (e·16+n)·(Nv+1).
template <class FT>
struct Col2 { FT a[2]; };

template <class FT>
__device__ __forceinline__ void up2(const FT (&x)[2], FT (&hi)[2], int lane) {  // value at level v+1
  hi[0] = __shfl_down_sync(FULLM, x[0], 1);
  FT t = __shfl_sync(FULLM, x[1], 0);
  if (lane == 31) hi[0] = t;
  hi[1] = __shfl_down_sync(FULLM, x[1], 1);
}
template <class FT>
__device__ __forceinline__ void dn2(const FT (&x)[2], FT (&lo)[2], int lane) {  // value at level v-1
  lo[0] = __shfl_up_sync(FULLM, x[0], 1);
  lo[1] = __shfl_up_sync(FULLM, x[1], 1);
  FT t = __shfl_sync(FULLM, x[0], 31);
  if (lane == 0) lo[1] = t;
}

template <class FT>
__global__ void __launch_bounds__(256)
k3_imp_stage(Par<FT> P, const FT* __restrict__ hgeo, const VLev<FT>* __restrict__ vlev, const FT* __restrict__ Yc,
             const FT* __restrict__ Yf, FT* __restrict__ Nc, FT* __restrict__ Nf, FT dtg, int ncols) {
  __shared__ FT sws[8][4][LV];  // per-warp Thomas workspace: l, d, u, r
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 8 + w;
  if (col >= ncols) return;  // whole warp exits together
  const int e = col >> 4, n = col & 15, nv = P.nv, nf = nv + 1;
  const FT kap = P.R_d / P.cv_d;
  const FT* gY = Yc + ((size_t)e * P.ncf * 16 + n) * nv;
  const FT* gYf = Yf + ((size_t)e * 16 + n) * nf;
  FT* gN = Nc + ((size_t)e * P.ncf * 16 + n) * nv;
  FT* gNf = Nf + ((size_t)e * 16 + n) * nf;
  const size_t cs = (size_t)16 * nv;  // component stride
  const FT g11 = hgeo[((size_t)e * HG_N + HG_GI11) * 16 + n], g12 = hgeo[((size_t)e * HG_N + HG_GI12) * 16 + n],
           g22 = hgeo[((size_t)e * HG_N + HG_GI22) * 16 + n];
  FT rho[2], re[2], u3[2], Kh[2];
  FT mc[2], rmc[2], g33[2], phic[2], sc[2];
  bool cv[2], fv[2], fin[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int v = lane + 32 * q;
    cv[q] = v < nv; fv[q] = v < nf; fin[q] = v > 0 && v < nv;
    const int vc = cv[q] ? v : nv - 1, vf = fv[q] ? v : nv;
    mc[q] = vlev->mc[vc]; rmc[q] = vlev->rmc[vc]; phic[q] = vlev->phic[vc]; sc[q] = vlev->sc2i[vc]; g33[q] = vlev->g33f[vf];
    rho[q] = cv[q] ? gY[v] : FT(1);
    FT a1 = cv[q] ? gY[cs + v] : FT(0), a2 = cv[q] ? gY[2 * cs + v] : FT(0);
    re[q] = cv[q] ? gY[3 * cs + v] : FT(0);
    if (cv[q]) { gN[cs + v] = a1; gN[2 * cs + v] = a2; }
    u3[q] = fin[q] ? gYf[v] : FT(0);  // cache_imp! boundary filter applied on load
    FT c1 = g11 * a1 + g12 * a2, c2 = g12 * a1 + g22 * a2;
    Kh[q] = FT(0.5) * ((a1 * c1 + a2 * c2) * sc[q]);
  }
  // ---- centre thermodynamics
  FT w3[2], w3h[2], g33h[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) w3[q] = g33[q] * u3[q] * u3[q];
  up2(w3, w3h, lane);
  FT hh[2], Pi[2], thv[2], thp[2], phr[2], dp[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    FT K = Kh[q] + FT(0.25) * (w3[q] + w3h[q]);
    Pt<FT> t = thermo(P, rho[q], re[q], K, phic[q]);
    hh[q] = t.h; Pi[q] = t.Pi; thv[q] = t.thv; thp[q] = t.thp; phr[q] = t.phir;
    dp[q] = kap * (P.T_0 * P.cp_d - K - phic[q]) + (P.R_d - kap * P.cv_d) * t.T;
  }
  // ---- face quantities: M = ᶠinterp(ρJ)u³/J2,  A = dtγ ᶠinterp(ρJ) g³³/J2  (zero on the boundary faces)
  FT rm[2], rml[2], hl[2], hl2[2], hu[2], M[2], A[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) rm[q] = rho[q] * mc[q];
  dn2(rm, rml, lane); dn2(hh, hl, lane); dn2(hl, hl2, lane); up2(hh, hu, lane);
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    FT mr = FT(0.5) * (rml[q] + rm[q]);
    A[q] = fin[q] ? dtg * mr * g33[q] : FT(0);
    M[q] = fin[q] ? mr * (g33[q] * u3[q]) : FT(0);
  }
  FT Am[2], Ap[2], Mm[2], Mp[2], rl[2], Pil[2], thvl[2], thpl[2], phrl[2], dpl[2], u3m[2], u3p[2], g33m[2], g33p[2], rmcl[2];
  dn2(A, Am, lane); up2(A, Ap, lane); dn2(M, Mm, lane); up2(M, Mp, lane);
  dn2(rho, rl, lane); dn2(Pi, Pil, lane); dn2(thv, thvl, lane); dn2(thp, thpl, lane); dn2(phr, phrl, lane); dn2(dp, dpl, lane);
  {
    FT gu[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) gu[q] = g33[q] * u3[q];
    dn2(gu, u3m, lane); up2(gu, u3p, lane);  // g³³u₃ at faces f-1 and f+1
  }
  dn2(rmc, rmcl, lane);
  FT* sl = sws[w][0]; FT* sd = sws[w][1]; FT* su = sws[w][2]; FT* sr = sws[w][3];
  FT ru_lo[2], ru_hi[2], eu_lo[2], eu_hi[2], rr[2], rre[2];  // centre-row pieces reused in the back-substitution
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int f = lane + 32 * q;
    FT beta = P.rayleigh ? vlev->brw[fv[q] ? f : nv] : FT(0);
    FT l = FT(0), d = dtg * (-beta) - FT(1), u = FT(0), r = FT(0);
    FT hm = FT(0.5) * (hl2[q] + hl[q]), h0 = FT(0.5) * (hl[q] + hh[q]), hp = FT(0.5) * (hh[q] + hu[q]);
    if (f <= 1) hm = FT(0);
    if (f >= nv - 1) hp = FT(0);
    // centre row f ("b") pieces (also used after the solve): ru_lo = A[f]/m_c[f], ru_hi = −A[f+1]/m_c[f]
    ru_lo[q] = A[q] * rmc[q]; ru_hi[q] = -Ap[q] * rmc[q];
    FT h0c = (f > 0) ? h0 : FT(0);
    eu_lo[q] = ru_lo[q] * h0c; eu_hi[q] = ru_hi[q] * hp;
    rr[q] = -dtg * (Mp[q] - M[q]) * rmc[q];
    rre[q] = -dtg * (Mp[q] * hp - M[q] * h0c) * rmc[q];
    if (fin[q]) {
      FT irf = rcp_(FT(0.5) * (rl[q] + rho[q]));
      FT dPi = Pi[q] - Pil[q];
      FT buoy = P.cp_d * (FT(0.5) * (thvl[q] + thv[q])) * dPi * irf;
      FT ur_lo = dtg * (irf * dpl[q] + buoy * FT(0.5)), ur_hi = dtg * (-irf * dp[q] + buoy * FT(0.5));
      FT ue_lo = dtg * irf * kap, ue_hi = -ue_lo;
      FT x_lo = irf * (-kap * rl[q]), x_hi = -irf * (-kap * rho[q]);
      FT k0 = FT(0.5) * (g33[q] * u3[q]);
      l = dtg * (x_lo * (FT(0.5) * u3m[q]));
      d = dtg * ((x_lo * k0 + x_hi * k0) - beta) - FT(1);
      u = dtg * (x_hi * (FT(0.5) * u3p[q]));
      FT ima = rmcl[q];
      FT ru_lo_a = Am[q] * ima, ru_hi_a = -A[q] * ima;
      l += ur_lo * ru_lo_a + ue_lo * (ru_lo_a * hm);
      d += ur_lo * ru_hi_a + ur_hi * ru_lo[q] + ue_lo * (ru_hi_a * h0) + ue_hi * (ru_lo[q] * h0);
      u += ur_hi * ru_hi[q] + ue_hi * (ru_hi[q] * hp);
      FT rr_a = -dtg * (M[q] - Mm[q]) * ima, re_a = -dtg * (M[q] * h0 - Mm[q] * hm) * ima;
      FT tf = -(vlev->dphif[f] - (phr[q] - phrl[q]) + P.cp_d * (FT(0.5) * (thpl[q] + thp[q])) * dPi) - beta * u3[q];
      r = dtg * tf + ur_lo * rr_a + ur_hi * rr[q] + ue_lo * re_a + ue_hi * rre[q];
    }
    if (fv[q]) { sl[f] = l; sd[f] = d; su[f] = u; sr[f] = r; }
  }
  __syncwarp();
  if (lane == 0) {  // Thomas sweep (BlockArrowheadSolve → tridiagonal solve of the Schur complement)
    FT rd = rcp_(sd[0]);
    FT cp = su[0] * rd, dq = sr[0] * rd;
    su[0] = cp; sr[0] = dq;
    for (int i = 1; i < nf; ++i) {
      FT li = sl[i];
      rd = rcp_(sd[i] - li * cp);
      cp = su[i] * rd;
      dq = (sr[i] - li * dq) * rd;
      su[i] = cp; sr[i] = dq;
    }
    FT x = dq;
    for (int i = nf - 2; i >= 0; --i) { x = sr[i] - su[i] * x; sr[i] = x; }
  }
  __syncwarp();
  FT x[2], xp[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) x[q] = fv[q] ? sr[lane + 32 * q] : FT(0);
  up2(x, xp, lane);
  // ---- U ← U − ΔU
  FT nre[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int v = lane + 32 * q;
    FT nr = rho[q] - (ru_lo[q] * x[q] + ru_hi[q] * xp[q] - rr[q]);
    nre[q] = re[q] - (eu_lo[q] * x[q] + eu_hi[q] * xp[q] - rre[q]);
    rho[q] = cv[q] ? nr : FT(1);
    u3[q] = fin[q] ? u3[q] - x[q] : FT(0);
    if (cv[q]) gN[v] = rho[q];
    if (fv[q]) gNf[v] = u3[q];
  }
  if (P.upwinding != 0) {
    // ---- cache_imp! after the Newton update (only h_tot is needed) and T_post_imp!
#pragma unroll
    for (int q = 0; q < 2; ++q) w3[q] = g33[q] * u3[q] * u3[q];
    up2(w3, w3h, lane);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      FT K = Kh[q] + FT(0.25) * (w3[q] + w3h[q]);
      FT etot = nre[q] * rcp_(rho[q]);
      FT T = fmax_(P.T_min_sgs, P.T_0 + ((etot - K - phic[q]) + P.RT0) * P.icv);
      hh[q] = etot + P.R_d * T;
      rm[q] = rho[q] * mc[q];
    }
    dn2(rm, rml, lane); dn2(hh, hl, lane); dn2(hl, hl2, lane); up2(hh, hu, lane);
    FT F[2], Fp[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int f = lane + 32 * q;
      F[q] = FT(0);
      if (fin[q]) {
        FT wv = g33[q] * u3[q];
        FT am = hl[q], ap = hh[q];
        FT cen = FT(0.5) * (am + ap), upw;
        if (P.upwinding == 3 && f >= 2 && f <= nv - 2) {
          if (wv >= FT(0)) upw = am + vl_slope(hl2[q], am, ap) / FT(2) * (FT(1) - wv * P.dt);
          else upw = ap - vl_slope(am, ap, hu[q]) / FT(2) * (FT(1) + wv * P.dt);
        } else {
          upw = wv >= FT(0) ? am : ap;
        }
        F[q] = FT(0.5) * (rml[q] + rm[q]) * wv * (upw - cen);
      }
    }
    up2(F, Fp, lane);
#pragma unroll
    for (int q = 0; q < 2; ++q) nre[q] += dtg * (-(Fp[q] - F[q]) * rmc[q]);
  }
#pragma unroll
  for (int q = 0; q < 2; ++q)
    if (cv[q]) gN[3 * cs + lane + 32 * q] = nre[q];
}

}  // namespace b200
