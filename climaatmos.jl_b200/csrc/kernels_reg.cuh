// kernels_reg.cuh — register-resident element kernels (round-1 optimisation of the explicit tendency).
//
// Mapping: one element per CTA, ONE THREAD PER LEVEL (lane = level, 2 warps for Nv+1 ≤ 64), and each
// thread holds all 16 GLL nodes of its level in registers.  Consequences on B200:
//   * global loads/stores are naturally coalesced (level is the fastest index of VIJFH) and go
//     straight to registers — no shared-memory staging of field slabs at all;
//   * every horizontal operator (4×4 contraction with D or Dw along ξ¹/ξ²) is thread-local FMA work
//     on register arrays with the matrix entries coming from the constant bank;
//   * vertical neighbours (k±1) are warp shuffles, with a 16-value shared-memory hand-off at the
//     boundary between the two warps;
//   * shared memory only holds the 11×16 per-node metric terms of the element (broadcast reads).
// The ncu profile of the first (shared-memory-staged) kernels showed them LSU/latency-bound at
// 24 % warp occupancy with ≈500 LDS per point (profiles/r1_ncu_summary.md); this layout needs
// ≈10 broadcast LDS per point.
//
//   k_exp_s   scalar part of remaining_tendency! before the DSS: ρₜ, ρe_totₜ (split-form flux
//             divergences advection.jl:48,59; viscous sponge viscous_sponge.jl:79) and ∇²(s_d − s_d,r)
//             (hyperdiffusion.jl:142-147)
//   k_exp_m   momentum part: uₕₜ, u₃ₜ (split-form PGF advection.jl:82-88, vector-invariant vertical
//             advection advection.jl:228-278, Rayleigh/viscous sponges) and ∇²u (hyperdiffusion.jl:141)
//   k_exp_c   apply_hyperdiffusion_tendency! (hyperdiffusion.jl:273-307) after the DSS
#pragma once
#include "common.cuh"

namespace b200 {

__constant__ float c_Df[32];   // D[16] then Dw[16]
__constant__ double c_Dd[32];
template <class FT> __device__ __forceinline__ FT cM(int k);
template <> __device__ __forceinline__ float cM<float>(int k) { return c_Df[k]; }
template <> __device__ __forceinline__ double cM<double>(int k) { return c_Dd[k]; }

constexpr unsigned FULLM = 0xffffffffu;
constexpr int RT = 64;  // threads per element CTA (levels)

// W = 0: strong matrix D, W = 16: weak matrix Dw.  n = j*4 + i.
template <class FT, int W>
__device__ __forceinline__ FT dxi(const FT (&a)[16], int i, int j) {
  return cM<FT>(W + i * 4 + 0) * a[j * 4 + 0] + cM<FT>(W + i * 4 + 1) * a[j * 4 + 1] +
         cM<FT>(W + i * 4 + 2) * a[j * 4 + 2] + cM<FT>(W + i * 4 + 3) * a[j * 4 + 3];
}
template <class FT, int W>
__device__ __forceinline__ FT deta(const FT (&a)[16], int i, int j) {
  return cM<FT>(W + j * 4 + 0) * a[0 + i] + cM<FT>(W + j * 4 + 1) * a[4 + i] + cM<FT>(W + j * 4 + 2) * a[8 + i] +
         cM<FT>(W + j * 4 + 3) * a[12 + i];
}

template <class FT>
__device__ __forceinline__ void ld16(FT (&a)[16], const FT* __restrict__ g, int nlev, int v, bool ok, FT dflt) {
#pragma unroll
  for (int n = 0; n < 16; ++n) a[n] = ok ? g[n * nlev + v] : dflt;
}
template <class FT>
__device__ __forceinline__ void st16(const FT (&a)[16], FT* __restrict__ g, int nlev, int v, bool ok) {
  if (ok) {
#pragma unroll
    for (int n = 0; n < 16; ++n) g[n * nlev + v] = a[n];
  }
}

// value of the level below (v-1) / above (v+1); lanes at the ends of the column keep their own value
template <class FT>
__device__ __forceinline__ void shfl_up16(const FT (&a)[16], FT (&lo)[16]) {
#pragma unroll
  for (int n = 0; n < 16; ++n) lo[n] = __shfl_up_sync(FULLM, a[n], 1);
}
template <class FT>
__device__ __forceinline__ void shfl_dn16(const FT (&a)[16], FT (&hi)[16]) {
#pragma unroll
  for (int n = 0; n < 16; ++n) hi[n] = __shfl_down_sync(FULLM, a[n], 1);
}
// hand-off between the two warps of a column: publish, (sync), fix
template <class FT>
__device__ __forceinline__ void pub_up(const FT (&a)[16], FT* x) {  // warp 0 lane 31 → level 32's "below"
  if (threadIdx.x == 31) {
#pragma unroll
    for (int n = 0; n < 16; ++n) x[n] = a[n];
  }
}
template <class FT>
__device__ __forceinline__ void fix_up(FT (&lo)[16], const FT* x) {
  if (threadIdx.x == 32) {
#pragma unroll
    for (int n = 0; n < 16; ++n) lo[n] = x[n];
  }
}
template <class FT>
__device__ __forceinline__ void pub_dn(const FT (&a)[16], FT* x) {  // warp 1 lane 0 → level 31's "above"
  if (threadIdx.x == 32) {
#pragma unroll
    for (int n = 0; n < 16; ++n) x[n] = a[n];
  }
}
template <class FT>
__device__ __forceinline__ void fix_dn(FT (&hi)[16], const FT* x) {
  if (threadIdx.x == 31) {
#pragma unroll
    for (int n = 0; n < 16; ++n) hi[n] = x[n];
  }
}

template <class FT>
struct Lev {  // per-thread level constants
  FT sc, mc, rmc, phi, g33lo, g33hi, bruh, bvc;   // centre v
  FT sf, sf2i, dzf, mclo, sclo, bvf;              // face v (mclo/sclo: centre v-1)
};
template <class FT>
__device__ __forceinline__ Lev<FT> load_lev(const VLev<FT>* __restrict__ V, int v, int nv) {
  Lev<FT> L;
  const int vc = v < nv ? v : nv - 1, vm = v > 0 ? v - 1 : 0, vf = v <= nv ? v : nv, vf1 = v + 1 <= nv ? v + 1 : nv;
  L.sc = V->sc2i[vc]; L.mc = V->mc[vc]; L.rmc = V->rmc[vc]; L.phi = V->phic[vc]; L.g33lo = V->g33f[vf]; L.g33hi = V->g33f[vf1];
  L.bruh = V->bruh[vc]; L.bvc = V->bvc[vc];
  L.sf = V->sf[vf]; L.sf2i = V->sf2i[vf]; L.dzf = V->dzf[vf]; L.mclo = V->mc[vm < nv ? vm : nv - 1];
  L.sclo = V->sc2i[vm < nv ? vm : nv - 1]; L.bvf = V->bvf[vf];
  return L;
}

// ---------------------------------------------------------------------------------------------
template <class FT>
__global__ void __launch_bounds__(RT) k_exp_s(Par<FT> P, const FT* __restrict__ hgeo, const VLev<FT>* __restrict__ vlev,
                                              const FT* __restrict__ Yc, const FT* __restrict__ Yf, FT* __restrict__ Ytc,
                                              FT* __restrict__ H) {
  __shared__ FT hg[HG_ELEM * 16];
  __shared__ FT xch[16];
  const int e = blockIdx.x, v = threadIdx.x, nv = P.nv, nf = nv + 1;
  const bool cv = v < nv, fv = v < nf;
  for (int k = threadIdx.x; k < HG_ELEM * 16; k += RT) hg[k] = hgeo[(size_t)e * HG_N * 16 + k];
  const Lev<FT> L = load_lev(vlev, v, nv);
  const FT* gY = Yc + (size_t)e * 64 * nv;
  FT rho[16], U1[16], U2[16], hh[16], ss[16];
  {
    FT u1[16], u2[16], re[16], u3[16], u3h[16];
    ld16(rho, gY, nv, v, cv, FT(1)); ld16(u1, gY + 16 * nv, nv, v, cv, FT(0)); ld16(u2, gY + 32 * nv, nv, v, cv, FT(0));
    ld16(re, gY + 48 * nv, nv, v, cv, FT(0)); ld16(u3, Yf + (size_t)e * 16 * nf, nf, v, fv, FT(0));
    shfl_dn16(u3, u3h);
    pub_dn(u3, xch);
    __syncthreads();  // also covers hg
    fix_dn(u3h, xch);
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      FT c1 = hg[HG_GI11 * 16 + n] * u1[n] + hg[HG_GI12 * 16 + n] * u2[n];
      FT c2 = hg[HG_GI12 * 16 + n] * u1[n] + hg[HG_GI22 * 16 + n] * u2[n];
      FT K = FT(0.5) * ((u1[n] * c1 + u2[n] * c2) * L.sc + FT(0.5) * (u3[n] * (L.g33lo * u3[n]) + u3h[n] * (L.g33hi * u3h[n])));
      Pt<FT> t = thermo(P, rho[n], re[n], K, L.phi);
      hh[n] = t.h;
      FT sdv = P.cp_d * (t.T - P.T_0) + L.phi;
      ss[n] = sdv - t.sdr;
      U1[n] = hg[HG_J2 * 16 + n] * c1; U2[n] = hg[HG_J2 * 16 + n] * c2;
      re[n] = sdv;  // keep s_d for the viscous sponge in the dead ρe_tot registers
    }
    // viscous sponge: β wdivₕ(ρ gradₕ(s_d))  (viscous_sponge.jl:79), only above zd
    FT vis[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) vis[n] = FT(0);
    if (P.viscous && L.bvc != FT(0)) {
      FT S1[16], S2[16];
#pragma unroll
      for (int n = 0; n < 16; ++n) {
        const int i = n & 3, j = n >> 2;
        FT g1 = dxi<FT, 0>(re, i, j), g2 = deta<FT, 0>(re, i, j);
        FT rj = rho[n] * hg[HG_J2 * 16 + n];
        S1[n] = rj * (hg[HG_GI11 * 16 + n] * g1 + hg[HG_GI12 * 16 + n] * g2);
        S2[n] = rj * (hg[HG_GI12 * 16 + n] * g1 + hg[HG_GI22 * 16 + n] * g2);
      }
#pragma unroll
      for (int n = 0; n < 16; ++n) {
        const int i = n & 3, j = n >> 2;
        vis[n] = L.bvc * (L.sc * (dxi<FT, 16>(S1, i, j) + deta<FT, 16>(S2, i, j)) * hg[HG_RJ2 * 16 + n]);
      }
    }
#pragma unroll
    for (int n = 0; n < 16; ++n) u3[n] = vis[n];  // park in dead registers
    // ---- split-form flux divergences (advection.jl:48,59)
    FT F1[16], F2[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) { F1[n] = rho[n] * U1[n]; F2[n] = rho[n] * U2[n]; }
    FT wd[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const int i = n & 3, j = n >> 2;
      wd[n] = (dxi<FT, 16>(F1, i, j) + deta<FT, 16>(F2, i, j)) * hg[HG_RJ2 * 16 + n] * L.sc;
    }
    FT* gT = Ytc + (size_t)e * 64 * nv;
    if (cv) {
#pragma unroll
      for (int n = 0; n < 16; ++n) gT[n * nv + v] = -wd[n];
    }
    FT G1[16], G2[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) { G1[n] = F1[n] * hh[n]; G2[n] = F2[n] * hh[n]; }
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const int i = n & 3, j = n >> 2;
      FT wdh = (dxi<FT, 16>(G1, i, j) + deta<FT, 16>(G2, i, j)) * hg[HG_RJ2 * 16 + n] * L.sc;
      FT gh1 = dxi<FT, 0>(hh, i, j), gh2 = deta<FT, 0>(hh, i, j);
      FT et = -(FT(0.5) * wdh + FT(0.5) * (hh[n] * wd[n] + (F1[n] * gh1 + F2[n] * gh2) * hg[HG_RJ2 * 16 + n] * L.sc));
      if (cv) gT[(48 + n) * nv + v] = et + u3[n];
    }
  }
  // ---- ∇²(s_d − s_d,r) = wdivₕ(gradₕ(·))  (hyperdiffusion.jl:142-147)
  if (H) {
    FT Q1[16], Q2[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const int i = n & 3, j = n >> 2;
      FT g1 = dxi<FT, 0>(ss, i, j), g2 = deta<FT, 0>(ss, i, j);
      Q1[n] = hg[HG_J2 * 16 + n] * (hg[HG_GI11 * 16 + n] * g1 + hg[HG_GI12 * 16 + n] * g2);
      Q2[n] = hg[HG_J2 * 16 + n] * (hg[HG_GI12 * 16 + n] * g1 + hg[HG_GI22 * 16 + n] * g2);
    }
    FT* gH = H + (size_t)e * 64 * nv + 48 * nv;
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const int i = n & 3, j = n >> 2;
      FT Ls = L.sc * (dxi<FT, 16>(Q1, i, j) + deta<FT, 16>(Q2, i, j)) * hg[HG_RJ2 * 16 + n];
      if (cv) gH[n * nv + v] = Ls;
    }
  }
}

// ---------------------------------------------------------------------------------------------
template <class FT>
__global__ void __launch_bounds__(RT) k_exp_m(Par<FT> P, const FT* __restrict__ hgeo, const VLev<FT>* __restrict__ vlev,
                                              const FT* __restrict__ Yc, const FT* __restrict__ Yf, FT* __restrict__ Ytc,
                                              FT* __restrict__ Ytf, FT* __restrict__ H) {
  __shared__ FT hg[HG_ELEM * 16];
  __shared__ FT xch[6][16];
  const int e = blockIdx.x, v = threadIdx.x, nv = P.nv, nf = nv + 1;
  const bool cv = v < nv, fv = v < nf, interior = (v > 0 && v < nv);
  for (int k = threadIdx.x; k < HG_ELEM * 16; k += RT) hg[k] = hgeo[(size_t)e * HG_N * 16 + k];
  const Lev<FT> L = load_lev(vlev, v, nv);
  const FT* gY = Yc + (size_t)e * 64 * nv;
  FT rho[16], u1[16], u2[16], u3[16], U1[16], U2[16], t1[16], t2[16], dK[16];
  ld16(rho, gY, nv, v, cv, FT(1)); ld16(u1, gY + 16 * nv, nv, v, cv, FT(0)); ld16(u2, gY + 32 * nv, nv, v, cv, FT(0));
  ld16(u3, Yf + (size_t)e * 16 * nf, nf, v, fv, FT(0));
  FT u3c[16];  // ᶜinterp(u₃) (covariant, component-wise)
  {
    FT re[16], u3h[16];
    ld16(re, gY + 48 * nv, nv, v, cv, FT(0));
    shfl_dn16(u3, u3h);
    pub_dn(u3, xch[0]);
    __syncthreads();
    fix_dn(u3h, xch[0]);
    FT sE[16], Pi[16], th[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      FT c1 = hg[HG_GI11 * 16 + n] * u1[n] + hg[HG_GI12 * 16 + n] * u2[n];
      FT c2 = hg[HG_GI12 * 16 + n] * u1[n] + hg[HG_GI22 * 16 + n] * u2[n];
      FT K = FT(0.5) * ((u1[n] * c1 + u2[n] * c2) * L.sc + FT(0.5) * (u3[n] * (L.g33lo * u3[n]) + u3h[n] * (L.g33hi * u3h[n])));
      Pt<FT> t = thermo(P, rho[n], re[n], K, L.phi);
      sE[n] = (K + L.phi) - t.phir; Pi[n] = t.Pi; th[n] = t.thp;
      U1[n] = hg[HG_J2 * 16 + n] * c1; U2[n] = hg[HG_J2 * 16 + n] * c2;
      dK[n] = K;
      u3c[n] = FT(0.5) * (u3[n] + u3h[n]);
      re[n] = t.thp * t.Pi;
    }
    // split-form PGF (advection.jl:82-88)
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const int i = n & 3, j = n >> 2;
      t1[n] = -(dxi<FT, 0>(sE, i, j) + P.cp_d * (th[n] * dxi<FT, 0>(Pi, i, j) + dxi<FT, 0>(re, i, j) - Pi[n] * dxi<FT, 0>(th, i, j)) / FT(2));
      t2[n] = -(deta<FT, 0>(sE, i, j) + P.cp_d * (th[n] * deta<FT, 0>(Pi, i, j) + deta<FT, 0>(re, i, j) - Pi[n] * deta<FT, 0>(th, i, j)) / FT(2));
    }
  }
  // ---- ∇²u: horizontal components (hyperdiffusion.jl:141) — also the viscous-sponge Laplacian
  {
    FT D2[16], ze[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const int i = n & 3, j = n >> 2;
      D2[n] = (dxi<FT, 0>(U1, i, j) + deta<FT, 0>(U2, i, j)) * hg[HG_RJ2 * 16 + n];
      ze[n] = (dxi<FT, 0>(u2, i, j) - deta<FT, 0>(u1, i, j)) * hg[HG_RJ2 * 16 + n];
    }
    FT* gH = H ? H + (size_t)e * 64 * nv : nullptr;
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const int i = n & 3, j = n >> 2;
      FT dz1 = dxi<FT, 16>(ze, i, j), dz2 = deta<FT, 16>(ze, i, j);
      FT L1 = L.sc * (dxi<FT, 16>(D2, i, j) - (hg[HG_GC11 * 16 + n] * dz2 - hg[HG_GC12 * 16 + n] * dz1) * hg[HG_RJ2 * 16 + n]);
      FT L2 = L.sc * (deta<FT, 16>(D2, i, j) - (hg[HG_GC12 * 16 + n] * dz2 - hg[HG_GC22 * 16 + n] * dz1) * hg[HG_RJ2 * 16 + n]);
      if (gH && cv) { gH[n * nv + v] = L1; gH[(16 + n) * nv + v] = L2; }
      if (P.viscous) { t1[n] += L.bvc * L1; t2[n] += L.bvc * L2; }
    }
    if (gH) {  // ∇²u₃ = wdivₕ(gradₕ(ᶜinterp(u₃))) on the flat shell
      FT P1[16], P2[16];
#pragma unroll
      for (int n = 0; n < 16; ++n) {
        const int i = n & 3, j = n >> 2;
        FT g1 = dxi<FT, 0>(u3c, i, j), g2 = deta<FT, 0>(u3c, i, j);
        P1[n] = hg[HG_J2 * 16 + n] * (hg[HG_GI11 * 16 + n] * g1 + hg[HG_GI12 * 16 + n] * g2);
        P2[n] = hg[HG_J2 * 16 + n] * (hg[HG_GI12 * 16 + n] * g1 + hg[HG_GI22 * 16 + n] * g2);
      }
#pragma unroll
      for (int n = 0; n < 16; ++n) {
        const int i = n & 3, j = n >> 2;
        FT L3 = L.sc * (dxi<FT, 16>(P1, i, j) + deta<FT, 16>(P2, i, j)) * hg[HG_RJ2 * 16 + n];
        if (cv) gH[(32 + n) * nv + v] = L3;
      }
    }
  }
  // ---- (ᶜf³ + ᶜω³) × CT12(ᶜu), Rayleigh sponge (advection.jl:228,275-277; remaining_tendency.jl:166)
#pragma unroll
  for (int n = 0; n < 16; ++n) {
    const int i = n & 3, j = n >> 2;
    FT wz = L.sc * (dxi<FT, 16>(u2, i, j) - deta<FT, 16>(u1, i, j)) * hg[HG_RJ2 * 16 + n];
    FT tot = hg[HG_COR3 * 16 + n] + wz;
    t1[n] += tot * U2[n]; t2[n] -= tot * U1[n];
    if (P.rayleigh) { t1[n] -= L.bruh * u1[n]; t2[n] -= L.bruh * u2[n]; }
  }
  // ---- face level v: ᶠω¹², mass flux, u₃ tendency (advection.jl:233-237,273-278)
  FT X1[16], X2[16];
  {
    FT rl[16], a1[16], a2[16], b1[16], b2[16], kl[16];
    shfl_up16(rho, rl); shfl_up16(u1, a1); shfl_up16(u2, a2); shfl_up16(U1, b1); shfl_up16(U2, b2); shfl_up16(dK, kl);
    pub_up(rho, xch[0]); pub_up(u1, xch[1]); pub_up(u2, xch[2]); pub_up(U1, xch[3]); pub_up(U2, xch[4]); pub_up(dK, xch[5]);
    __syncthreads();
    fix_up(rl, xch[0]); fix_up(a1, xch[1]); fix_up(a2, xch[2]); fix_up(b1, xch[3]); fix_up(b2, xch[4]); fix_up(kl, xch[5]);
    __syncthreads();
    FT* gF = Ytf + (size_t)e * 16 * nf;
    FT R1[16], R2[16];
    const bool vis3 = P.viscous && L.bvf != FT(0);
    if (vis3) {
#pragma unroll
      for (int n = 0; n < 16; ++n) {
        const int i = n & 3, j = n >> 2;
        FT g1 = dxi<FT, 0>(u3, i, j), g2 = deta<FT, 0>(u3, i, j);
        R1[n] = hg[HG_J2 * 16 + n] * (hg[HG_GI11 * 16 + n] * g1 + hg[HG_GI12 * 16 + n] * g2);
        R2[n] = hg[HG_J2 * 16 + n] * (hg[HG_GI12 * 16 + n] * g1 + hg[HG_GI22 * 16 + n] * g2);
      }
    }
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const int i = n & 3, j = n >> 2;
      const FT J2 = hg[HG_J2 * 16 + n], rJ2 = hg[HG_RJ2 * 16 + n];
      FT jt1 = J2 * L.sf * L.dzf * hg[HG_COR1 * 16 + n] + deta<FT, 16>(u3, i, j);
      FT jt2 = J2 * L.sf * L.dzf * hg[HG_COR2 * 16 + n] - dxi<FT, 16>(u3, i, j);
      FT Vn, ub1, ub2, dk = FT(0);
      if (interior) {
        jt1 -= (u2[n] - a2[n]); jt2 += (u1[n] - a1[n]);
        Vn = FT(0.5) * (rl[n] * L.mclo + rho[n] * L.mc);
        ub1 = FT(0.5) * (b1[n] * L.sclo + U1[n] * L.sc) * rJ2;
        ub2 = FT(0.5) * (b2[n] * L.sclo + U2[n] * L.sc) * rJ2;
        dk = dK[n] - kl[n];
      } else if (v == 0) {
        Vn = rho[n] * L.mc; ub1 = U1[n] * L.sc * rJ2; ub2 = U2[n] * L.sc * rJ2;
      } else {  // top face: extrapolate from the last centre (= level below)
        Vn = rl[n] * L.mclo; ub1 = b1[n] * L.sclo * rJ2; ub2 = b2[n] * L.sclo * rJ2;
      }
      Vn *= L.g33lo * u3[n];
      X1[n] = jt2 * Vn; X2[n] = -jt1 * Vn;
      FT t3 = -(jt1 * ub2 - jt2 * ub1) - dk;
      if (vis3) t3 += L.bvf * (L.sf2i * (dxi<FT, 16>(R1, i, j) + deta<FT, 16>(R2, i, j)) * rJ2);
      if (fv) gF[n * nf + v] = t3;
    }
  }
  {
    FT h1[16], h2[16];
    shfl_dn16(X1, h1); shfl_dn16(X2, h2);
    pub_dn(X1, xch[0]); pub_dn(X2, xch[1]);
    __syncthreads();
    fix_dn(h1, xch[0]); fix_dn(h2, xch[1]);
    FT* gT = Ytc + (size_t)e * 64 * nv;
    if (cv) {
#pragma unroll
      for (int n = 0; n < 16; ++n) {
        FT rm = rho[n] * L.mc;
        gT[(16 + n) * nv + v] = t1[n] - FT(0.5) * (X1[n] + h1[n]) / rm;
        gT[(32 + n) * nv + v] = t2[n] - FT(0.5) * (X2[n] + h2[n]) / rm;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
template <class FT>
__global__ void __launch_bounds__(RT) k_exp_c(Par<FT> P, const FT* __restrict__ hgeo, const VLev<FT>* __restrict__ vlev,
                                              const FT* __restrict__ Yc, const FT* __restrict__ H, FT* __restrict__ Ytc,
                                              FT* __restrict__ Ytf) {
  __shared__ FT hg[HG_ELEM * 16];
  __shared__ FT xch[3][16];
  const int e = blockIdx.x, v = threadIdx.x, nv = P.nv, nf = nv + 1;
  const bool cv = v < nv, fv = v < nf;
  for (int k = threadIdx.x; k < HG_ELEM * 16; k += RT) hg[k] = hgeo[(size_t)e * HG_N * 16 + k];
  const Lev<FT> L = load_lev(vlev, v, nv);
  const FT* gH = H + (size_t)e * 64 * nv;
  FT* gT = Ytc + (size_t)e * 64 * nv;
  FT rho[16];
  ld16(rho, Yc + (size_t)e * 64 * nv, nv, v, cv, FT(1));
  __syncthreads();
  {  // ∇⁴uₕ = δ_div·wgradₕ(divₕ(∇²u)) − wcurlₕ(curlₕ(∇²u))  (hyperdiffusion.jl:273-276)
    FT L1[16], L2[16], D2[16], ze[16];
    ld16(L1, gH, nv, v, cv, FT(0)); ld16(L2, gH + 16 * nv, nv, v, cv, FT(0));
    {
      FT U1[16], U2[16];
#pragma unroll
      for (int n = 0; n < 16; ++n) {
        U1[n] = hg[HG_J2 * 16 + n] * (hg[HG_GI11 * 16 + n] * L1[n] + hg[HG_GI12 * 16 + n] * L2[n]);
        U2[n] = hg[HG_J2 * 16 + n] * (hg[HG_GI12 * 16 + n] * L1[n] + hg[HG_GI22 * 16 + n] * L2[n]);
      }
#pragma unroll
      for (int n = 0; n < 16; ++n) {
        const int i = n & 3, j = n >> 2;
        D2[n] = (dxi<FT, 0>(U1, i, j) + deta<FT, 0>(U2, i, j)) * hg[HG_RJ2 * 16 + n];
        ze[n] = (dxi<FT, 0>(L2, i, j) - deta<FT, 0>(L1, i, j)) * hg[HG_RJ2 * 16 + n];
      }
    }
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const int i = n & 3, j = n >> 2;
      FT dz1 = dxi<FT, 16>(ze, i, j), dz2 = deta<FT, 16>(ze, i, j);
      FT Qa = L.sc * (P.ddf * dxi<FT, 16>(D2, i, j) - (hg[HG_GC11 * 16 + n] * dz2 - hg[HG_GC12 * 16 + n] * dz1) * hg[HG_RJ2 * 16 + n]);
      FT Qb = L.sc * (P.ddf * deta<FT, 16>(D2, i, j) - (hg[HG_GC12 * 16 + n] * dz2 - hg[HG_GC22 * 16 + n] * dz1) * hg[HG_RJ2 * 16 + n]);
      if (cv) { gT[(16 + n) * nv + v] -= P.nu4v * Qa; gT[(32 + n) * nv + v] -= P.nu4v * Qb; }
    }
  }
  {  // Yₜ.ρe_tot −= ν₄ₛ wdivₕ(ρ gradₕ(∇²s_d))  (hyperdiffusion.jl:291,307)
    FT Ls[16], Q1[16], Q2[16];
    ld16(Ls, gH + 48 * nv, nv, v, cv, FT(0));
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const int i = n & 3, j = n >> 2;
      FT g1 = dxi<FT, 0>(Ls, i, j), g2 = deta<FT, 0>(Ls, i, j);
      FT rj = rho[n] * hg[HG_J2 * 16 + n];
      Q1[n] = rj * (hg[HG_GI11 * 16 + n] * g1 + hg[HG_GI12 * 16 + n] * g2);
      Q2[n] = rj * (hg[HG_GI12 * 16 + n] * g1 + hg[HG_GI22 * 16 + n] * g2);
    }
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const int i = n & 3, j = n >> 2;
      FT Le = L.sc * (dxi<FT, 16>(Q1, i, j) + deta<FT, 16>(Q2, i, j)) * hg[HG_RJ2 * 16 + n];
      if (cv) gT[(48 + n) * nv + v] -= P.nu4s * Le;
    }
  }
  {  // Yₜ.f.u₃ −= ν₄ᵥ ᶠwinterp(ᶜJ ρ, C3(∇⁴u))  (hyperdiffusion.jl:277)
    FT L3[16], P1[16], P2[16], w[16], a[16], q[16];
    ld16(L3, gH + 32 * nv, nv, v, cv, FT(0));
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const int i = n & 3, j = n >> 2;
      FT g1 = dxi<FT, 0>(L3, i, j), g2 = deta<FT, 0>(L3, i, j);
      P1[n] = hg[HG_J2 * 16 + n] * (hg[HG_GI11 * 16 + n] * g1 + hg[HG_GI12 * 16 + n] * g2);
      P2[n] = hg[HG_J2 * 16 + n] * (hg[HG_GI12 * 16 + n] * g1 + hg[HG_GI22 * 16 + n] * g2);
    }
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const int i = n & 3, j = n >> 2;
      q[n] = L.sc * (dxi<FT, 16>(P1, i, j) + deta<FT, 16>(P2, i, j)) * hg[HG_RJ2 * 16 + n];
      w[n] = L.mc * rho[n];
      a[n] = w[n] * q[n];
    }
    FT wl[16], al[16], ql[16];
    shfl_up16(w, wl); shfl_up16(a, al); shfl_up16(q, ql);
    pub_up(w, xch[0]); pub_up(a, xch[1]); pub_up(q, xch[2]);
    __syncthreads();
    fix_up(wl, xch[0]); fix_up(al, xch[1]); fix_up(ql, xch[2]);
    FT* gF = Ytf + (size_t)e * 16 * nf;
    if (fv) {
#pragma unroll
      for (int n = 0; n < 16; ++n) {
        FT val = (v == 0) ? q[n] : (v == nv ? ql[n] : (al[n] + a[n]) / (wl[n] + w[n]));
        gF[n * nf + v] -= P.nu4v * val;
      }
    }
  }
}

}  // namespace b200
