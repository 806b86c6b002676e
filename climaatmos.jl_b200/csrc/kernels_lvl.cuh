// kernels_lvl.cuh — hyperdiffusion apply with ONE THREAD PER (element, level): all 16 GLL nodes of the level in registers.
//
// Why (round 2, profiles/r2_tma_experiment.md): ncu shows k5_exp_c (thread = level × GLL row, kernels_pair.cuh) waiting on the MIO —
// short_scoreboard + mio_throttle are ≈ 8 of its 17 stall cycles per issue: every ξ²-contraction is a shuffle reduce-scatter (3 SHFL
// per value) and the metric pairs come from shared memory.  With the whole 4 × 4 level in one thread BOTH contractions are thread-local
// FFMA2 chains with the matrix entries from the constant bank: no shuffles, no shared memory, no block barrier (part 2 excepted: one
// value per node from the level below), and the global accesses are 128-byte lines (a warp = 32 consecutive levels of one node).
// The same idea lost for the pre-DSS kernel in round 1 (255 registers, instruction-fetch-bound); the three parts of the hyperdiffusion
// apply are small enough (≈ 100 registers each).
//
// Same arithmetic and operation ORDER as the kernel it replaced (k5_exp_c, thread = level × GLL row; removed in round 2) — the butterfly
// summation order of its shuffle reduce-scatter is reproduced term by term — and it was bitwise identical to it on the B200 (Float32 and
// Float64, nv = 2, 10, 63; profiles/r2_k7_exp_c.md) before that kernel was deleted: 70.6 → 65.1 µs at he30/ze63.
#pragma once
#include "kernels_pair.cuh"

namespace b200 {

constexpr int LVL_EPB = 4;  // elements per CTA (64 threads = levels each)

// o[j] = Σ_k M_W[j][k]·a[k] over the rows of the level, in the summation order of deta4p (kernels_pair.cuh):
//   o_j = fma(a_j, M[j][j], a_{j^2}·M[j][j^2]) + fma(a_{j^1}, M[j][j^1], a_{j^3}·M[j][j^3])
template <class FT, int W>
__device__ __forceinline__ void deta16(const P2<FT> (&a)[4][2], P2<FT> (&o)[4][2]) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const P2<FT> qa = fma2(a[j][p], cM<FT>(W * 16 + j * 4 + j), a[j ^ 2][p] * cM<FT>(W * 16 + j * 4 + (j ^ 2)));
      const P2<FT> qb = fma2(a[j ^ 1][p], cM<FT>(W * 16 + j * 4 + (j ^ 1)), a[j ^ 3][p] * cM<FT>(W * 16 + j * 4 + (j ^ 3)));
      o[j][p] = qa + qb;
    }
}
template <class FT, int W>
__device__ __forceinline__ void dxi16(const P2<FT> (&a)[4][2], P2<FT> (&o)[4][2]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) dxi4p<FT, W>(a[j], o[j]);
}
// (weak or strong) divergence of the contravariant pair (a1, a2): ∂₂a2 + ∂₁a1, in the order of div4p
template <class FT, int W>
__device__ __forceinline__ void div16(const P2<FT> (&a1)[4][2], const P2<FT> (&a2)[4][2], P2<FT> (&o)[4][2]) {
  deta16<FT, W>(a2, o);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    P2<FT> t[2];
    dxi4p<FT, W>(a1[j], t);
    o[j][0] = o[j][0] + t[0]; o[j][1] = o[j][1] + t[1];
  }
}
// g points at (node 0, level v) of a slab; node n is n·nlev further on
template <class FT>
__device__ __forceinline__ void ld16(P2<FT> (&a)[4][2], const FT* __restrict__ g, int nlev, bool ok, FT dflt) {
#pragma unroll
  for (int j = 0; j < 4; ++j) ld4q(a[j], g + 4 * j * nlev, nlev, ok, dflt);
}
template <class FT>
__device__ __forceinline__ void st16(const P2<FT> (&a)[4][2], FT* __restrict__ g, int nlev) {
#pragma unroll
  for (int j = 0; j < 4; ++j) st4q(a[j], g + 4 * j * nlev, nlev);
}
// metric pair of component c for nodes (4j + 2p, 4j + 2p + 1), straight from global memory (warp-uniform address: one L1 transaction)
#define HG16(c, j, p) ldpair(&hg[(c) * 16 + 4 * (j) + 2 * (p)])

// blockIdx.y = part (0: ∇⁴uₕ → Yₜ.uₕ, 1: ρe_tot (+ the water terms of a moist context), 2: u₃, 3 + k: passive tracer k — apply_tracer_hyperdiffusion_tendency!,
// ρχₜ_lim −= ν₄ₛ wdivₕ(ρ gradₕ ∇²χ), hyperdiffusion.jl:524-532, into Tlim = Yₜ_lim.c or Yₜ.c; the same arithmetic as part 1 on H[4 + k],
// which replaced the row-layout kernel k5_tracer_c, 21.8 µs per tracer).  One launch for all parts: per-part launches with their own
// register budgets (part 1 fits 64 registers, part 0 80) measured slower (74 vs 65 µs, profiles/r2_k7_exp_c.md).
template <class FT, int NVC>
__global__ void __launch_bounds__(LVL_EPB * 64, (sizeof(FT) == 4 ? 2 : 1))
k7_exp_c(Par<FT> P, const FT* __restrict__ hgeo, const VLev<FT>* __restrict__ vlev, const FT* __restrict__ Yc,
         const FT* __restrict__ H, FT* __restrict__ Ytc, FT* __restrict__ Ytf, FT* __restrict__ Tlim = nullptr,
         const FT* __restrict__ Hw = nullptr) {
  using V = P2<FT>;
  extern __shared__ __align__(16) unsigned char smem_raw[];  // LVL_EPB·16 values: part 2, node values of level 31 for the thread of level 32
  FT (*s_q)[16] = reinterpret_cast<FT (*)[16]>(smem_raw);     // (the two warps of an element)
  pdl_launch();
  const int v = threadIdx.x & 63, le = threadIdx.x >> 6, e = blockIdx.x * LVL_EPB + le;
  const int nv = NVC ? NVC : P.nv, nf = nv + 1;
  const bool live = e < P.nh, cv = live && v < nv, fv = live && v < nf;
  const int part = blockIdx.y;
  const int vc = v < nv ? v : nv - 1, vm = v > 0 ? v - 1 : 0;
  const FT sc = vlev->sc2i[vc], mc = vlev->mc[vc], mclo = vlev->mc[vm < nv ? vm : nv - 1];
  const FT* hg = hgeo + (size_t)(live ? e : 0) * HG_N * 16;
  pdl_wait(Yc, H, Ytc, Ytf, Tlim, Hw);
  const size_t offc = (size_t)(live ? e : 0) * P.ncf * 16 * nv + v;  // (node 0, level v) of component 0
  const int cs = 16 * nv;
  if (part == 0) {  // ∇⁴uₕ = δ_div·wgradₕ(divₕ(∇²u)) − wcurlₕ(curlₕ(∇²u))  (hyperdiffusion.jl:273-276)
    V L1[4][2], L2[4][2], U1[4][2], U2[4][2], D2[4][2], ze[4][2], a[4][2], g1[4][2];
    ld16(L1, H + offc, nv, cv, FT(0)); ld16(L2, H + offc + cs, nv, cv, FT(0));
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const V w_ = HG16(HG_J2, j, p);
        U1[j][p] = w_ * fma2(HG16(HG_GI12, j, p), L2[j][p], HG16(HG_GI11, j, p) * L1[j][p]);
        U2[j][p] = w_ * fma2(HG16(HG_GI22, j, p), L2[j][p], HG16(HG_GI12, j, p) * L1[j][p]);
      }
    div16<FT, 0>(U1, U2, D2);
    deta16<FT, 0>(L1, a);
    dxi16<FT, 0>(L2, g1);
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        D2[j][p] = D2[j][p] * HG16(HG_RJ2, j, p);
        ze[j][p] = (g1[j][p] - a[j][p]) * HG16(HG_RJ2, j, p);
      }
    V b[4][2], dD1[4][2], dz1[4][2], old1[4][2], old2[4][2];
    deta16<FT, 1>(D2, a); deta16<FT, 1>(ze, b);
    dxi16<FT, 1>(D2, dD1); dxi16<FT, 1>(ze, dz1);
    ld16(old1, Ytc + offc + cs, nv, cv, FT(0)); ld16(old2, Ytc + offc + 2 * cs, nv, cv, FT(0));
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const V rJ2 = HG16(HG_RJ2, j, p);
        const V Qa = (dD1[j][p] * P.ddf - (HG16(HG_GC11, j, p) * b[j][p] - HG16(HG_GC12, j, p) * dz1[j][p]) * rJ2) * sc;
        const V Qb = (a[j][p] * P.ddf - (HG16(HG_GC12, j, p) * b[j][p] - HG16(HG_GC22, j, p) * dz1[j][p]) * rJ2) * sc;
        old1[j][p] = old1[j][p] - Qa * P.nu4v; old2[j][p] = old2[j][p] - Qb * P.nu4v;
      }
    if (cv) { st16(old1, Ytc + offc + cs, nv); st16(old2, Ytc + offc + 2 * cs, nv); }
  } else {
    // parts 1 and 2 share the scalar Laplacian wdivₕ(w·gradₕ(L)):  part 1: L = ∇²s_d, w = ρ·J2  (hyperdiffusion.jl:291,307);
    //                                                              part 2: L = ∇²u₃,  w = J2    (hyperdiffusion.jl:277)
    V rho[4][2], Ls[4][2], g1[4][2], g2[4][2], Q1[4][2], Q2[4][2], b[4][2];
    ld16(rho, Yc + offc, nv, cv, FT(1));
    // component of the scalar: 3 = ρe_tot (part 1), 2 = ∇²u₃ (part 2), first passive tracer + k (part 3 + k; a moist context keeps
    // component 4 for the active ρq_tot, whose water terms are in part 1)
    const int comp = part == 1 ? 3 : part == 2 ? 2 : 4 + (P.moist ? 1 : 0) + (part - 3);
    ld16(Ls, H + offc + comp * cs, nv, cv, FT(0));
    deta16<FT, 0>(Ls, g2);
    dxi16<FT, 0>(Ls, g1);
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const V w_ = part != 2 ? rho[j][p] * HG16(HG_J2, j, p) : HG16(HG_J2, j, p);
        Q1[j][p] = w_ * fma2(HG16(HG_GI12, j, p), g2[j][p], HG16(HG_GI11, j, p) * g1[j][p]);
        Q2[j][p] = w_ * fma2(HG16(HG_GI22, j, p), g2[j][p], HG16(HG_GI12, j, p) * g1[j][p]);
      }
    div16<FT, 1>(Q1, Q2, b);
    if (part != 2) {
      V old3[4][2];
      FT* tgt = (part == 1 ? Ytc : Tlim) + offc + comp * cs;
      ld16(old3, tgt, nv, cv, FT(0));
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int p = 0; p < 2; ++p) old3[j][p] = old3[j][p] - ((b[j][p] * sc) * HG16(HG_RJ2, j, p)) * P.nu4s;
      if (part == 1 && P.moist) {
        // water part of the apply for a moist (0M) context, on the DSSed ∇²q_tot_eff = H[4] (this replaced the row-layout kernel k_moist_c):
        //   d = ν₄ₛ wdivₕ(ρ gradₕ(∇²q_tot_eff)):  ρq_totₜ −= d and ρₜ −= d, both in Tlim = Yₜ_lim or Yₜ   (hyperdiffusion.jl:475-484)
        //   ρe_totₜ −= ν₄ₛ wdivₕ(ρ (h_eff + Φ) gradₕ(∇²q_tot_eff))                                       (:293-307); Hw = ρ(h_eff + Φ) from k5_exp_a
        V Lq[4][2], rh[4][2], d[4][2];
        ld16(Lq, H + offc + 4 * cs, nv, cv, FT(0));
        ld16(rh, Hw + ((size_t)(live ? e : 0) * 16 * nv + v), nv, cv, FT(0));
        deta16<FT, 0>(Lq, g2);
        dxi16<FT, 0>(Lq, g1);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            const V w_ = rho[j][p] * HG16(HG_J2, j, p);
            Q1[j][p] = w_ * fma2(HG16(HG_GI12, j, p), g2[j][p], HG16(HG_GI11, j, p) * g1[j][p]);
            Q2[j][p] = w_ * fma2(HG16(HG_GI22, j, p), g2[j][p], HG16(HG_GI12, j, p) * g1[j][p]);
          }
        div16<FT, 1>(Q1, Q2, b);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int p = 0; p < 2; ++p) d[j][p] = ((b[j][p] * sc) * HG16(HG_RJ2, j, p)) * P.nu4s;
#pragma unroll
        for (int k = 0; k < 2; ++k) {  // ρ (component 0) and ρq_tot (component 4)
          V o_[4][2];
          FT* tl = Tlim + offc + (k ? 4 : 0) * cs;
          ld16(o_, tl, nv, cv, FT(0));
#pragma unroll
          for (int j = 0; j < 4; ++j) { o_[j][0] = o_[j][0] - d[j][0]; o_[j][1] = o_[j][1] - d[j][1]; }
          if (cv) st16(o_, tl, nv);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            const V w_ = rh[j][p] * HG16(HG_J2, j, p);
            Q1[j][p] = w_ * fma2(HG16(HG_GI12, j, p), g2[j][p], HG16(HG_GI11, j, p) * g1[j][p]);
            Q2[j][p] = w_ * fma2(HG16(HG_GI22, j, p), g2[j][p], HG16(HG_GI12, j, p) * g1[j][p]);
          }
        div16<FT, 1>(Q1, Q2, b);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int p = 0; p < 2; ++p) old3[j][p] = old3[j][p] - ((b[j][p] * sc) * HG16(HG_RJ2, j, p)) * P.nu4s;
      }
      if (cv) st16(old3, tgt, nv);
    } else {  // Yₜ.f.u₃ −= ν₄ᵥ ᶠwinterp(ᶜJ ρ, C3(∇⁴u)): face v from the centres v − 1 and v
      V q[4][2], ql[4][2], rlo[4][2], oldf[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int p = 0; p < 2; ++p) q[j][p] = (b[j][p] * sc) * HG16(HG_RJ2, j, p);
      // q of the level below: lane − 1 of the warp, or (first lane of the upper warp) the last lane of the lower warp through shared memory
      const int lane = threadIdx.x & 31;
      if (v == 31) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int p = 0; p < 2; ++p) { s_q[le][4 * j + 2 * p] = q[j][p].lo(); s_q[le][4 * j + 2 * p + 1] = q[j][p].hi(); }
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const FT lo = __shfl_up_sync(FULLM, q[j][p].lo(), 1), hi = __shfl_up_sync(FULLM, q[j][p].hi(), 1);
          ql[j][p] = (lane == 0 && v == 32) ? V(s_q[le][4 * j + 2 * p], s_q[le][4 * j + 2 * p + 1]) : V(lo, hi);
        }
      const size_t offf = (size_t)(live ? e : 0) * 16 * nf + v;
      ld16(rlo, Yc + offc - (v > 0 ? 1 : 0), nv, live && v > 0 && v <= nv, FT(1));  // ρ of the level below
      ld16(oldf, Ytf + offf, nf, fv, FT(0));
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const V w = rho[j][p] * mc, wl = rlo[j][p] * mclo;
          V val;
          if (v == 0) val = q[j][p];
          else if (v == nv) val = ql[j][p];
          else {
            const V num = fma2(w, q[j][p], wl * ql[j][p]), den = wl + w;
            val = V(num.lo() / den.lo(), num.hi() / den.hi());
          }
          oldf[j][p] = oldf[j][p] - val * P.nu4v;
        }
      if (fv) st16(oldf, Ytf + offf, nf);
    }
  }
}

}  // namespace b200
