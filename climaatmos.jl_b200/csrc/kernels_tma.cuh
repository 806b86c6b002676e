// kernels_tma.cuh — persistent, bulk-copy-fed versions of the element kernels of the benchmarked step.
//
// Same arithmetic, operation order and thread ↔ (level, GLL row) mapping as the kernels of kernels_pair.cuh — the results are
// BITWISE identical (tests/test_gpu_parity.py::test_persistent_bulk_copy_kernels_are_bitwise_identical) — but the CTAs are
// persistent (grid = resident CTAs, each walks elements e, e + gridDim.x, …) and the inputs of an element arrive in a
// double-buffered shared-memory stage through 1-D bulk asynchronous copies (bulk.cuh: cp.async.bulk + mbarrier, SASS UBLKCP):
//   iteration i:  wait(full[i & 1]) → registers ← stage, contractions, exchange slabs, stores of element i → __syncthreads →
//                 one thread re-arms stage i & 1 with element i + 2      (the copy of element i + 1 has been in flight since the end of
//                 iteration i − 1 and lands while element i is processed)
// What this removes from the one-CTA-per-element kernels (ncu, profiles/r1_ncu_full_session2_kernels.txt: k5_exp_c
// long_scoreboard 5.0 stalls per issue): every global load with a thread waiting on it, the per-CTA prologue (metric terms,
// level constants, derivative-matrix columns are loaded once per CTA instead of once per element) and the CTA launch/retire gaps.
#pragma once
#include "bulk.cuh"
#include "kernels_pair.cuh"

namespace b200 {

// everything of B200_ROW_PROLOGUE_NV that does not depend on the element
#define B200_ROW_PROLOGUE_PERSISTENT(NVC_)                                                                                  \
  const int lane = threadIdx.x & 31, vl = lane & 7, j = lane >> 3;                                                          \
  const int v = (threadIdx.x >> 5) * 8 + vl, nv = (NVC_) ? (NVC_) : P.nv, nf = nv + 1;                                      \
  const bool cv = v < nv, fv = v < nf;                                                                                      \
  FT md[4], mw[4];                                                                                                          \
  _Pragma("unroll") for (int k = 0; k < 4; ++k) { md[k] = cM<FT>((j ^ k) * 4 + j); mw[k] = cM<FT>(16 + (j ^ k) * 4 + j); }  \
  const Lev<FT> L = load_lev(vlev, v, nv);                                                                                  \
  const int n0 = j * 4;

// ---------------------------------------------------------------------------------------------
// k6_exp_c — apply_hyperdiffusion_tendency! after the DSS of the ∇² fields (hyperdiffusion.jl:247-307), persistent.
// blockIdx.y = part as in k5_exp_c (0: ∇⁴uₕ → Yₜ.uₕ, 1: ρe_tot, 2: u₃).  Stage = [metric terms 13·16 | 4 slabs of 16·nlev],
// slab contents per part:
//   part 0: ∇²u₁, ∇²u₂ (one copy: adjacent components of H), Yₜ.uₕ₁, Yₜ.uₕ₂ (one copy)
//   part 1: ρ, ∇²s_d, Yₜ.ρe_tot                    part 2: ρ, ∇²u₃, Yₜ.u₃ (16·(nv+1) values)
constexpr int K6_HG = HG_ELEM * 16;                  // metric terms staged per element (FT words)
constexpr int K6C_STAGE = K6_HG + 4 * 16 * LV;       // FT words per stage
template <class FT> constexpr size_t smem_k6c() { return (2 * (size_t)K6C_STAGE + 2 * XSLAB) * sizeof(FT) + 2 * sizeof(mbar_t); }

template <class FT, int NVC>
__global__ void __launch_bounds__(CT, (sizeof(FT) == 4 ? EXPC_MINB : 2))
k6_exp_c(Par<FT> P, const FT* __restrict__ hgeo, const VLev<FT>* __restrict__ vlev, const FT* __restrict__ Yc,
         const FT* __restrict__ H, FT* __restrict__ Ytc, FT* __restrict__ Ytf) {
  using V = P2<FT>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  FT* stage0 = reinterpret_cast<FT*>(smem_raw);
  constexpr int SW = K6C_STAGE;
  FT* s_w = stage0 + 2 * SW;
  FT* s_a = s_w + XSLAB;
  mbar_t* full = reinterpret_cast<mbar_t*>(s_a + XSLAB);
  pdl_launch();
  B200_ROW_PROLOGUE_PERSISTENT(NVC)
  const int part = blockIdx.y;
  const int cs = 16 * nv;                      // words per centre slab
  const size_t ec = (size_t)P.ncf * cs;        // element stride of Y.c / H / Yₜ.c
  const unsigned slab_b = (unsigned)(cs * sizeof(FT)), fslab_b = (unsigned)(16 * nf * sizeof(FT)), hg_b = (unsigned)(K6_HG * sizeof(FT));
  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  pdl_wait(Yc, H, Ytc, Ytf);
  // one thread arms stage s with element e
  auto issue = [&](int e, int s) {
    FT* st = stage0 + s * SW;
    FT* sl = st + K6_HG;
    mbar_t* bar = &full[s];
    if (part == 0) {
      mbar_expect_tx(bar, hg_b + 4 * slab_b);
      bulk_g2s(st, hgeo + (size_t)e * HG_N * 16, hg_b, bar);
      bulk_g2s(sl, H + e * ec, 2 * slab_b, bar);
      bulk_g2s(sl + 2 * cs, Ytc + e * ec + cs, 2 * slab_b, bar);
    } else if (part == 1) {
      mbar_expect_tx(bar, hg_b + 3 * slab_b);
      bulk_g2s(st, hgeo + (size_t)e * HG_N * 16, hg_b, bar);
      bulk_g2s(sl, Yc + e * ec, slab_b, bar);
      bulk_g2s(sl + cs, H + e * ec + 3 * cs, slab_b, bar);
      bulk_g2s(sl + 2 * cs, Ytc + e * ec + 3 * cs, slab_b, bar);
    } else {
      mbar_expect_tx(bar, hg_b + 2 * slab_b + fslab_b);
      bulk_g2s(st, hgeo + (size_t)e * HG_N * 16, hg_b, bar);
      bulk_g2s(sl, Yc + e * ec, slab_b, bar);
      bulk_g2s(sl + cs, H + e * ec + 2 * cs, slab_b, bar);
      bulk_g2s(sl + 2 * cs, Ytf + (size_t)e * 16 * nf, fslab_b, bar);
    }
  };
  const int e0 = blockIdx.x, de = gridDim.x;
  if (threadIdx.x == 0) {
    if (e0 < P.nh) issue(e0, 0);
    if (e0 + de < P.nh) issue(e0 + de, 1);
  }
  int it = 0;
  for (int e = e0; e < P.nh; e += de, ++it) {
    const int s = it & 1;
    const FT* hg = stage0 + s * SW;              // metric terms of element e (HGP / METRIC_FLUX read them from here)
    const FT* sl = hg + K6_HG;
    mbar_wait(&full[s], (it >> 1) & 1);
    const size_t offc = e * ec + (n0 * nv + v);
    FT* gT = Ytc + offc;
    FT* gF = Ytf + ((size_t)e * 16 * nf + (n0 * nf + v));
    const FT* q0 = sl + (n0 * nv + v);           // (row j, level v) of slab 0; the four nodes of the row are nv apart
    V a[2], b[2], g1[2];
    if (part == 0) {  // ∇⁴uₕ = δ_div·wgradₕ(divₕ(∇²u)) − wcurlₕ(curlₕ(∇²u))  (hyperdiffusion.jl:273-276)
      V L1[2], L2[2], old1[2], old2[2];
      ld4q(L1, q0, nv, cv, FT(0)); ld4q(L2, q0 + cs, nv, cv, FT(0));
      ld4q(old1, q0 + 2 * cs, nv, cv, FT(0)); ld4q(old2, q0 + 3 * cs, nv, cv, FT(0));
      V U1[2], U2[2], D2[2], ze[2], dD1[2], dz1[2];
      METRIC_FLUX(U1, U2, L1, L2, HGP(HG_J2, p))
      div4p<FT, 0>(U1, U2, md, vl, D2);
      deta4p(L1, md, vl, a);
      dxi4p<FT, 0>(L2, g1);
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        D2[p] = D2[p] * HGP(HG_RJ2, p);
        ze[p] = (g1[p] - a[p]) * HGP(HG_RJ2, p);
      }
      deta4p(D2, mw, vl, a); deta4p(ze, mw, vl, b);
      dxi4p<FT, 1>(D2, dD1); dxi4p<FT, 1>(ze, dz1);
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        V rJ2 = HGP(HG_RJ2, p);
        V Qa = (dD1[p] * P.ddf - (HGP(HG_GC11, p) * b[p] - HGP(HG_GC12, p) * dz1[p]) * rJ2) * L.sc;
        V Qb = (a[p] * P.ddf - (HGP(HG_GC12, p) * b[p] - HGP(HG_GC22, p) * dz1[p]) * rJ2) * L.sc;
        old1[p] = old1[p] - Qa * P.nu4v; old2[p] = old2[p] - Qb * P.nu4v;
      }
      if (cv) { st4q(old1, gT + 16 * nv, nv); st4q(old2, gT + 32 * nv, nv); }
    } else if (part == 1) {  // Yₜ.ρe_tot −= ν₄ₛ wdivₕ(ρ gradₕ(∇²s_d))  (hyperdiffusion.jl:291,307)
      V rho[2], Ls[2], old3[2], Q1[2], Q2[2];
      ld4q(rho, q0, nv, cv, FT(1));
      ld4q(Ls, q0 + cs, nv, cv, FT(0));
      ld4q(old3, q0 + 2 * cs, nv, cv, FT(0));
      deta4p(Ls, md, vl, a);
      dxi4p<FT, 0>(Ls, g1);
      METRIC_FLUX(Q1, Q2, g1, a, rho[p] * HGP(HG_J2, p))
      div4p<FT, 1>(Q1, Q2, mw, vl, b);
#pragma unroll
      for (int p = 0; p < 2; ++p) old3[p] = old3[p] - ((b[p] * L.sc) * HGP(HG_RJ2, p)) * P.nu4s;
      if (cv) st4q(old3, gT + 48 * nv, nv);
    } else {  // Yₜ.f.u₃ −= ν₄ᵥ ᶠwinterp(ᶜJ ρ, C3(∇⁴u))  (hyperdiffusion.jl:277)
      V rho[2], L3[2], oldf[2], P1[2], P2_[2], q[2], w[2];
      ld4q(rho, q0, nv, cv, FT(1));
      ld4q(L3, q0 + cs, nv, cv, FT(0));
      ld4q(oldf, sl + 2 * cs + (n0 * nf + v), nf, fv, FT(0));
      deta4p(L3, md, vl, a);
      dxi4p<FT, 0>(L3, g1);
      METRIC_FLUX(P1, P2_, g1, a, HGP(HG_J2, p))
      div4p<FT, 1>(P1, P2_, mw, vl, b);
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        q[p] = (b[p] * L.sc) * HGP(HG_RJ2, p);
        w[p] = rho[p] * L.mc;
      }
      sputq(s_w, w, j, v); sputq(s_a, q, j, v);
      __syncthreads();
      if (fv) {
        V wl[2], ql[2];
        const int vm = v > 0 ? v - 1 : 0;
        sgetq(s_w, wl, j, vm); sgetq(s_a, ql, j, vm);
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          V val;
          if (v == 0) val = q[p];
          else if (v == nv) val = ql[p];
          else {
            V num = fma2(w[p], q[p], wl[p] * ql[p]), den = wl[p] + w[p];
            val = V(num.lo() / den.lo(), num.hi() / den.hi());
          }
          oldf[p] = oldf[p] - val * P.nu4v;
        }
        st4q(oldf, gF, nf);
      }
    }
    __syncthreads();  // every thread is done with stage s (and with the exchange slabs): re-arm it with the element after next
    if (threadIdx.x == 0 && e + 2 * de < P.nh) {
      fence_proxy_async();
      issue(e + 2 * de, s);
    }
  }
}

}  // namespace b200
