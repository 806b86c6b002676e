// kernels_imp8.cuh — fused implicit stage, fourth generation: a WARP PER COLUMN PAIR, no shared memory, no block barriers.
//
// Same arithmetic as k5_imp_stage (kernels_imp5.cuh; cache_imp! → Wfact → T_imp! residual → ldiv! → U −= ΔU → cache_imp! → T_post_imp!;
// implicit_tendency.jl:36-98,185-339, manual_sparse_jacobian.jl:713-870,504-585).  Layout: one element per CTA of 8 warps; warp w owns
// the columns (nodes) 2w and 2w+1 as ONE f32x2 pair, lane k owns the two consecutive levels 2k and 2k+1 of that pair — so a column of
// 64 faces is exactly one warp.  Consequences on B200:
//   * vertical neighbours are registers of the same thread (level 2k ↔ 2k+1) or one shuffle from the adjacent lane: the 13 pair slabs
//     (54 KB) and the 12 block barriers of k5_imp_stage are gone;
//   * the Schur tridiagonal systems are solved inside the warp: one cyclic-reduction step eliminates the odd rows thread-locally (one
//     shuffle set), the 32 even rows are reduced by parallel cyclic reduction with shuffles (5 steps), the odd rows follow by
//     back-substitution — 68 shuffles per solve instead of 108 64-bit shared-memory accesses and 7 barriers;
//   * pointwise algebra stays packed (FFMA2/FMUL2/FADD2 over the two columns); global accesses are 256-byte contiguous per warp and
//     field (two 32-bit accesses per thread: node stride 63 words is not 8-byte aligned for odd nodes).
#pragma once
#include "kernels_imp5.cuh"

namespace b200 {

template <class FT> __device__ __forceinline__ P2<FT> shup(const P2<FT>& a, int d = 1) {
  return P2<FT>(__shfl_up_sync(FULLM, a.lo(), d), __shfl_up_sync(FULLM, a.hi(), d));
}
template <class FT> __device__ __forceinline__ P2<FT> shdn(const P2<FT>& a, int d = 1) {
  return P2<FT>(__shfl_down_sync(FULLM, a.lo(), d), __shfl_down_sync(FULLM, a.hi(), d));
}
// the thread's two rows of a column-pair field: g points at (node 2w, level 2k); the second column is `nlev` further
template <class FT>
__device__ __forceinline__ void ld8(P2<FT> (&a)[2], const FT* __restrict__ g, int nlev, bool ok0, bool ok1, FT dflt) {
  const FT a00 = ok0 ? g[0] : dflt, a01 = ok0 ? g[nlev] : dflt, a10 = ok1 ? g[1] : dflt, a11 = ok1 ? g[nlev + 1] : dflt;
  a[0] = P2<FT>(a00, a01); a[1] = P2<FT>(a10, a11);
}
template <class FT>
__device__ __forceinline__ void st8(const P2<FT> (&a)[2], FT* __restrict__ g, int nlev, bool ok0, bool ok1) {
  if (ok0) { g[0] = a[0].lo(); g[nlev] = a[0].hi(); }
  if (ok1) { g[1] = a[1].lo(); g[nlev + 1] = a[1].hi(); }
}

// Tridiagonal solves inside the warp: rows 2k (p = 0) and 2k+1 (p = 1) of lane k, coefficients (l, d, u) and NR right-hand sides of
// the thread's column pair.  Rows outside the system must be identity rows (l = u = 0, y = 0); row 0 has l = 0 and the last row u = 0.
// The solutions replace the right-hand sides.  (cyclic reduction of the odd rows + PCR of the 32 even rows + back-substitution)
template <class FT, int NR>
__device__ __forceinline__ void warp_tridiag_n(int lane, const P2<FT> (&l)[2], const P2<FT> (&d)[2], const P2<FT> (&u)[2], P2<FT> (&y)[NR][2]) {
  using V2 = P2<FT>;
  V2 a[2], c[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const V2 rd = rcpn2(d[p]);
    a[p] = l[p] * rd; c[p] = u[p] * rd;
#pragma unroll
    for (int r = 0; r < NR; ++r) y[r][p] = y[r][p] * rd;
  }
  // eliminate x[2k−1] (odd row of lane k−1) and x[2k+1] (own odd row) from the even row 2k
  const V2 a1u = shup(a[1]), c1u = shup(c[1]);  // lane 0: a[0] = 0, so its own values are harmless
  V2 A, C, Y[NR];
  {
    const V2 rd = rcpn2(V2(FT(1)) - fma2(a[0], c1u, c[0] * a[1]));
    A = -((a[0] * a1u) * rd);
    C = -((c[0] * c[1]) * rd);
#pragma unroll
    for (int r = 0; r < NR; ++r) Y[r] = (y[r][0] - fma2(a[0], shup(y[r][1]), c[0] * y[r][1])) * rd;
  }
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) {
    const bool hm = lane >= s, hp = lane + s < 32;
    V2 Am = shup(A, s), Cm = shup(C, s), Ap = shdn(A, s), Cp = shdn(C, s);
    if (!hm) { Am = V2(FT(0)); Cm = V2(FT(0)); }
    if (!hp) { Ap = V2(FT(0)); Cp = V2(FT(0)); }
    const V2 rd = rcpn2(V2(FT(1)) - fma2(C, Ap, A * Cm));
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      V2 Ym = shup(Y[r], s), Yp = shdn(Y[r], s);
      if (!hm) Ym = V2(FT(0));
      if (!hp) Yp = V2(FT(0));
      Y[r] = (Y[r] - fma2(C, Yp, A * Ym)) * rd;
    }
    A = -((A * Am) * rd);
    C = -((C * Cp) * rd);
  }
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const V2 x0d = shdn(Y[r]);  // lane 31: c[1] = 0 (last row)
    y[r][1] = y[r][1] - fma2(a[1], Y[r], c[1] * x0d);
    y[r][0] = Y[r];
  }
}
template <class FT>
__device__ __forceinline__ void warp_tridiag(int lane, const P2<FT> (&l)[2], const P2<FT> (&d)[2], const P2<FT> (&u)[2], const P2<FT> (&r)[2],
                                             P2<FT> (&x)[2]) {
  P2<FT> y[1][2] = {{r[0], r[1]}};
  warp_tridiag_n<FT, 1>(lane, l, d, u, y);
  x[0] = y[0][0]; x[1] = y[0][1];
}

// LDIV / MOIST: as for k5_imp_stage — LDIV = true is ldiv!(ΔY, J, R) of the hook path on the state snapshot Wfact kept ((Rc, Rf) the
// right-hand side, (Nc, Nf) receive ΔY); MOIST = true the 0M-moist context (active ρq_tot = component 4: moist thermodynamic state,
// κ_m per point, the (ρq_tot, u₃) / (u₃, ρq_tot) blocks, q_tot transport and its post-Newton correction).
template <class FT, int NVC, bool LDIV = false, bool MOIST = false>
__global__ void __launch_bounds__(256, sizeof(FT) == 4 ? 2 : 1)
k8_imp_stage(Par<FT> P, const FT* __restrict__ hgeo, const VLev<FT>* __restrict__ vlev, const FT* __restrict__ Yc,
             const FT* __restrict__ Yf, FT* __restrict__ Nc, FT* __restrict__ Nf, FT dtg, const FT* __restrict__ Rc = nullptr,
             const FT* __restrict__ Rf = nullptr) {
  using V2 = P2<FT>;
  constexpr int Q0 = MOIST ? 5 : 4;  // first passive tracer
  pdl_launch();
  const int e = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5, n0 = 2 * w, nv = NVC ? NVC : P.nv, nf = nv + 1;
  const FT kap = P.R_d / P.cv_d;
  // per-level constants of the thread's two levels v = 2·lane + p
  bool cv[2], fv[2], interior[2];
  FT sc2i[2], phi[2], mc[2], mclo[2], rmc[2], rmclo[2], g33lo[2], g33hi[2], g33m[2], dphif[2], beta[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int v = 2 * lane + p;
    cv[p] = v < nv; fv[p] = v < nf; interior[p] = v > 0 && v < nv;
    const int vm = v > 0 ? v - 1 : 0;
    const int vc = cv[p] ? v : nv - 1, vmc = vm < nv ? vm : nv - 1, vf = fv[p] ? v : nv, vf1 = v + 1 <= nv ? v + 1 : nv;
    sc2i[p] = vlev->sc2i[vc]; phi[p] = vlev->phic[vc]; mc[p] = vlev->mc[vc]; mclo[p] = vlev->mc[vmc]; rmc[p] = vlev->rmc[vc];
    rmclo[p] = vlev->rmc[vmc]; g33lo[p] = vlev->g33f[vf]; g33hi[p] = vlev->g33f[vf1]; g33m[p] = vlev->g33f[vm < nf ? vm : nv];
    dphif[p] = vlev->dphif[vf]; beta[p] = P.rayleigh ? vlev->brw[vf] : FT(0);
  }
  const FT* hgp = hgeo + (size_t)e * HG_N * 16 + n0;
  const V2 g11 = ldpair(hgp + HG_GI11 * 16), g12 = ldpair(hgp + HG_GI12 * 16), g22 = ldpair(hgp + HG_GI22 * 16);
  pdl_wait(Yc, Yf, Nc, Nf, Rc, Rf);
  const int cs = 16 * nv;
  const FT* gY = Yc + ((size_t)e * P.ncf * 16 * nv + n0 * nv + 2 * lane);
  const FT* gYf = Yf + ((size_t)e * 16 * nf + n0 * nf + 2 * lane);
  FT* gN = Nc + ((size_t)e * P.ncf * 16 * nv + n0 * nv + 2 * lane);
  FT* gNf = Nf + ((size_t)e * 16 * nf + n0 * nf + 2 * lane);
  V2 rho[2], u1[2], u2[2], re[2], u3[2];
  ld8(rho, gY, nv, cv[0], cv[1], FT(1)); ld8(u1, gY + cs, nv, cv[0], cv[1], FT(0)); ld8(u2, gY + 2 * cs, nv, cv[0], cv[1], FT(0));
  ld8(re, gY + 3 * cs, nv, cv[0], cv[1], FT(0)); ld8(u3, gYf, nf, interior[0], interior[1], FT(0));  // u₃ boundary filter on load
  V2 rq[2], Rr[2], R1[2], R2[2], Re[2], R3[2], Rq[2];
  if (MOIST) ld8(rq, gY + 4 * cs, nv, cv[0], cv[1], FT(0));
  if (LDIV) {  // the right-hand side; Δuₕ = −R_uₕ ((uₕ,uₕ) = −I), Δ(ρχ) = −R_ρχ (passive tracers: the fallback −I block)
    const FT* gR = Rc + ((size_t)e * P.ncf * 16 * nv + n0 * nv + 2 * lane);
    ld8(Rr, gR, nv, cv[0], cv[1], FT(0)); ld8(R1, gR + cs, nv, cv[0], cv[1], FT(0)); ld8(R2, gR + 2 * cs, nv, cv[0], cv[1], FT(0));
    ld8(Re, gR + 3 * cs, nv, cv[0], cv[1], FT(0)); ld8(R3, Rf + ((size_t)e * 16 * nf + n0 * nf + 2 * lane), nf, fv[0], fv[1], FT(0));
    if (MOIST) ld8(Rq, gR + 4 * cs, nv, cv[0], cv[1], FT(0));
    const V2 m1[2] = {-R1[0], -R1[1]}, m2[2] = {-R2[0], -R2[1]};
    st8(m1, gN + cs, nv, cv[0], cv[1]); st8(m2, gN + 2 * cs, nv, cv[0], cv[1]);
    for (int q = Q0; q < P.ncf; ++q) {
      V2 t[2];
      ld8(t, gR + q * cs, nv, cv[0], cv[1], FT(0));
      const V2 m[2] = {-t[0], -t[1]};
      st8(m, gN + q * cs, nv, cv[0], cv[1]);
    }
  } else {  // uₕ is copied through (R_uₕ = 0); passive tracers: ΔU = 0
    st8(u1, gN + cs, nv, cv[0], cv[1]); st8(u2, gN + 2 * cs, nv, cv[0], cv[1]);
    for (int q = Q0; q < P.ncf; ++q) {
      V2 t[2];
      ld8(t, gY + q * cs, nv, cv[0], cv[1], FT(0));
      st8(t, gN + q * cs, nv, cv[0], cv[1]);
    }
  }
  // ---- centre thermodynamics and face mass-flux pieces
  V2 Kh[2], h[2], Pi[2], thv[2], thp[2], phr[2], dp[2], A[2], M[2], kapv[2], dpq[2], qv[2], ck1[2], ck2[2];
  const V2 u3d0 = shdn(u3[0]), rhou1 = shup(rho[1]);
  V2 u3h[2] = {u3[1], u3d0};     // u₃ at face v + 1
  V2 rlo[2] = {rhou1, rho[0]};   // ρ at centre v − 1
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const V2 c1 = fma2(g12, u2[p], g11 * u1[p]), c2 = fma2(g22, u2[p], g12 * u1[p]);
    Kh[p] = (fma2(u2[p], c2, u1[p] * c1) * sc2i[p]) * FT(0.5);
    ck1[p] = c1 * sc2i[p]; ck2[p] = c2 * sc2i[p];  // ∂K/∂uₕ = CT12(uₕ) (ldiv! only)
    h[p] = V2(FT(0)); Pi[p] = V2(FT(1)); thv[p] = V2(FT(0)); thp[p] = V2(FT(0)); phr[p] = V2(FT(1)); dp[p] = V2(FT(0));
    kapv[p] = V2(kap); dpq[p] = V2(FT(0)); qv[p] = V2(FT(0));
    if (cv[p]) {
      const V2 K = Kh[p] + (u3[p] * (u3[p] * g33lo[p]) + u3h[p] * (u3h[p] * g33hi[p])) * FT(0.25);
      Pt2<FT> t;
      if constexpr (MOIST) {
        Mst2<FT> m;
        t = thermo2m(P, rho[p], re[p], rq[p], K, phi[p], m);
        dpq[p] = dp_drhoq2(P, m, kapv[p]);
        qv[p] = div2(rq[p], rho[p]);
        dp[p] = t.T * (V2(P.R_d) - kapv[p] * P.cv_d) + ((V2(P.T_0 * P.cp_d) - K) - phi[p]) * kapv[p];
      } else {
        t = thermo2(P, rho[p], re[p], K, phi[p]);
        dp[p] = fma2(t.T, V2(P.R_d - kap * P.cv_d), ((V2(P.T_0 * P.cp_d) - K) - phi[p]) * kap);  // ∂p/∂ρ (manual_sparse_jacobian.jl:816-818)
      }
      h[p] = t.h; Pi[p] = t.Pi; thv[p] = t.thv; thp[p] = t.thp; phr[p] = pgf_aux2(t);
    }
    A[p] = M[p] = V2(FT(0));
    if (interior[p]) {
      const V2 mr = fma2(rho[p], V2(mc[p]), rlo[p] * mclo[p]) * FT(0.5);
      A[p] = (mr * dtg) * g33lo[p];
      M[p] = mr * (u3[p] * g33lo[p]);
    }
  }
  // ---- vertical neighbours: level v − 1 (m1), v − 2 (m2), v + 1 (p1) of the thread's two levels
  const V2 hu0 = shup(h[0]), hu1 = shup(h[1]), hd0 = shdn(h[0]);
  const V2 Au1 = shup(A[1]), Ad0 = shdn(A[0]), Mu1 = shup(M[1]), Md0 = shdn(M[0]), u3u1 = shup(u3[1]);
  const V2 Piu1 = shup(Pi[1]), thvu1 = shup(thv[1]), thpu1 = shup(thp[1]), phru1 = shup(phr[1]), dpu1 = shup(dp[1]);
  const V2 h_m1[2] = {hu1, h[0]}, h_m2[2] = {hu0, hu1}, h_p1[2] = {h[1], hd0};
  const V2 A_m1[2] = {Au1, A[0]}, A_p1[2] = {A[1], Ad0}, M_m1[2] = {Mu1, M[0]}, M_p1[2] = {M[1], Md0}, u3_m1[2] = {u3u1, u3[0]};
  const V2 Pi_m1[2] = {Piu1, Pi[0]}, thv_m1[2] = {thvu1, thv[0]}, thp_m1[2] = {thpu1, thp[0]}, phr_m1[2] = {phru1, phr[0]}, dp_m1[2] = {dpu1, dp[0]};
  // MOIST: κ_m, ∂p/∂ρq_tot at v − 1, q_tot at v − 1, v − 2, v + 1;  LDIV: the right-hand side and ∂K/∂uₕ at centre v − 1
  V2 kap_m1[2], dpq_m1[2], q_m1[2], q_m2[2], q_p1[2], Rr_m1[2], R1_m1[2], R2_m1[2], Re_m1[2], Rq_m1[2], ck1_m1[2], ck2_m1[2];
  if (MOIST) {
    const V2 ku = shup(kapv[1]), du = shup(dpq[1]), qu0 = shup(qv[0]), qu1 = shup(qv[1]), qd0 = shdn(qv[0]);
    kap_m1[0] = ku; kap_m1[1] = kapv[0]; dpq_m1[0] = du; dpq_m1[1] = dpq[0];
    q_m1[0] = qu1; q_m1[1] = qv[0]; q_m2[0] = qu0; q_m2[1] = qu1; q_p1[0] = qv[1]; q_p1[1] = qd0;
  }
  if (LDIV) {
    Rr_m1[0] = shup(Rr[1]); Rr_m1[1] = Rr[0]; R1_m1[0] = shup(R1[1]); R1_m1[1] = R1[0]; R2_m1[0] = shup(R2[1]); R2_m1[1] = R2[0];
    Re_m1[0] = shup(Re[1]); Re_m1[1] = Re[0]; ck1_m1[0] = shup(ck1[1]); ck1_m1[1] = ck1[0]; ck2_m1[0] = shup(ck2[1]); ck2_m1[1] = ck2[0];
    if (MOIST) { Rq_m1[0] = shup(Rq[1]); Rq_m1[1] = Rq[0]; }
  }
  // ---- Schur tridiagonal and right-hand side of face row v (manual_sparse_jacobian.jl:746-868)
  V2 R0[2], E0[2], a0[2], a1[2], b0[2], b1[2], cl[2], cd[2], cu[2], cr[2], Q0s[2], c0[2], c1q[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int v = 2 * lane + p;
    const V2 hl = h_m1[p], hm2 = h_m2[p], hp1 = h_p1[p];
    const V2 hf0 = v > 0 ? (hl + h[p]) * FT(0.5) : V2(FT(0));
    const V2 hfp = v < nv - 1 ? (h[p] + hp1) * FT(0.5) : V2(FT(0));
    const V2 Ap = A_p1[p], Mp = M_p1[p];
    V2 qf0(FT(0)), qfp(FT(0)), qfm(FT(0));  // MOIST: ᶠinterp(q_tot) at faces v, v + 1, v − 1
    {
      const V2 rr = ((Mp - M[p]) * (-dtg)) * rmc[p], rre = ((Mp * hfp - M[p] * hf0) * (-dtg)) * rmc[p];
      a0[p] = A[p] * rmc[p]; a1[p] = -(Ap * rmc[p]);
      b0[p] = a0[p] * hf0; b1[p] = a1[p] * hfp;
      R0[p] = rho[p] + rr; E0[p] = re[p] + rre;
      if constexpr (MOIST) {
        qf0 = v > 0 ? (q_m1[p] + qv[p]) * FT(0.5) : V2(FT(0));
        qfp = v < nv - 1 ? (qv[p] + q_p1[p]) * FT(0.5) : V2(FT(0));
        qfm = v > 1 ? (q_m2[p] + q_m1[p]) * FT(0.5) : V2(FT(0));
        c0[p] = a0[p] * qf0; c1q[p] = a1[p] * qfp;
        Q0s[p] = rq[p] + ((Mp * qfp - M[p] * qf0) * (-dtg)) * rmc[p];
      }
    }
    cl[p] = cu[p] = V2(FT(0)); cd[p] = V2(dtg * (-beta[p]) - FT(1));
    cr[p] = LDIV ? R3[p] : V2(FT(0));  // boundary rows of ldiv!: x = R₃/(−dtγβ − 1)
    if (interior[p]) {
      const V2 hfm = v > 1 ? (hm2 + hl) * FT(0.5) : V2(FT(0));
      const V2 Am = A_m1[p], Mm = M_m1[p], u3m = u3_m1[p];
      const V2 irf = rcpn2((rlo[p] + rho[p]) * FT(0.5));
      V2 dPi, dphr;
      pgf_diff2(P, Pi_m1[p], Pi[p], phr_m1[p], phr[p], dPi, dphr);
      const V2 buoy = ((((thv_m1[p] + thv[p]) * FT(0.5)) * P.cp_d) * dPi) * irf;
      const V2 hb = buoy * FT(0.5);
      const V2 ur_lo = fma2(irf, dp_m1[p], hb) * dtg, ur_hi = (hb - irf * dp[p]) * dtg;
      V2 ue_lo, ue_hi, x_lo, x_hi, uq_lo(FT(0)), uq_hi(FT(0));
      if constexpr (MOIST) {
        ue_lo = (irf * dtg) * kap_m1[p]; ue_hi = -((irf * dtg) * kapv[p]);
        x_lo = irf * (rlo[p] * (-kap_m1[p])); x_hi = -(irf * (rho[p] * (-kapv[p])));
        uq_lo = (irf * dtg) * dpq_m1[p]; uq_hi = -((irf * dtg) * dpq[p]);  // (u₃, ρq_tot): dtγ ᶠp_grad_matrix ⋅ Diag(∂p/∂ρq_tot)
      } else {
        ue_lo = (irf * dtg) * kap; ue_hi = -ue_lo;
        x_lo = irf * (rlo[p] * (-kap)); x_hi = -(irf * (rho[p] * (-kap)));
      }
      const V2 k0 = u3[p] * (FT(0.5) * g33lo[p]);
      V2 l = (x_lo * (u3m * (FT(0.5) * g33m[p]))) * dtg;
      V2 d = (fma2(x_hi, k0, x_lo * k0) - beta[p]) * dtg - FT(1);
      V2 u = (x_hi * (u3h[p] * (FT(0.5) * g33hi[p]))) * dtg;
      const V2 ru_lo_a = Am * rmclo[p], ru_hi_a = -(A[p] * rmclo[p]), ru_lo_b = a0[p], ru_hi_b = a1[p];
      l = l + fma2(ue_lo, ru_lo_a * hfm, ur_lo * ru_lo_a);
      d = d + (fma2(ur_hi, ru_lo_b, ur_lo * ru_hi_a) + fma2(ue_hi, ru_lo_b * hf0, ue_lo * (ru_hi_a * hf0)));
      u = u + fma2(ue_hi, ru_hi_b * hfp, ur_hi * ru_hi_b);
      const V2 rr_a = ((M[p] - Mm) * (-dtg)) * rmclo[p], rr_b = ((Mp - M[p]) * (-dtg)) * rmc[p];
      const V2 Mh0 = M[p] * hf0;
      const V2 re_a = ((Mh0 - Mm * hfm) * (-dtg)) * rmclo[p], re_b = ((Mp * hfp - Mh0) * (-dtg)) * rmc[p];
      const V2 tf = -((V2(dphif[p]) - dphr) + (((thp_m1[p] + thp[p]) * FT(0.5)) * P.cp_d) * dPi) - u3[p] * beta[p];
      V2 moist_rhs(FT(0));
      if constexpr (MOIST) {  // Schur terms of the ρq_tot column/row
        l = l + uq_lo * (ru_lo_a * qfm);
        d = d + fma2(uq_hi, ru_lo_b * qf0, uq_lo * (ru_hi_a * qf0));
        u = u + uq_hi * (ru_hi_b * qfp);
        const V2 Mq0 = M[p] * qf0;
        const V2 rq_a = ((Mq0 - Mm * qfm) * (-dtg)) * rmclo[p], rq_b = ((Mp * qfp - Mq0) * (-dtg)) * rmc[p];
        moist_rhs = LDIV ? fma2(uq_lo, Rq_m1[p], uq_hi * Rq[p]) : fma2(uq_lo, rq_a, uq_hi * rq_b);
      }
      cl[p] = l; cd[p] = d; cu[p] = u;
      if (LDIV) {  // Schur right-hand side R₃ + A₃ρ R_ρ + A₃e R_ρe + A₃uₕ R_uₕ
        const V2 xl = x_lo * dtg, xh = x_hi * dtg;
        cr[p] = R3[p] + (fma2(ur_lo, Rr_m1[p], ur_hi * Rr[p]) + fma2(ue_lo, Re_m1[p], ue_hi * Re[p])) +
                (fma2(xl * ck1_m1[p], R1_m1[p], (xh * ck1[p]) * R1[p]) + fma2(xl * ck2_m1[p], R2_m1[p], (xh * ck2[p]) * R2[p]));
      } else
        cr[p] = fma2(tf, V2(dtg), fma2(ur_lo, rr_a, ur_hi * rr_b) + fma2(ue_lo, re_a, ue_hi * re_b));
      if (MOIST) cr[p] = cr[p] + moist_rhs;
    }
  }
  V2 x0[2];
  warp_tridiag(lane, cl, cd, cu, cr, x0);
  const V2 x1[2] = {x0[1], shdn(x0[0])};  // ΔU.f.u₃ at face v + 1
  if (LDIV) {  // ΔY: Δu₃ = x, Δρ = A_ρ3 x − R_ρ, Δρe_tot = A_e3 x − R_ρe, Δρq_tot = A_q3 x − R_ρq (A₁₁ = −I)
    V2 dr[2], de[2], dq[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      dr[p] = fma2(a1[p], x1[p], a0[p] * x0[p]) - Rr[p];
      de[p] = fma2(b1[p], x1[p], b0[p] * x0[p]) - Re[p];
      if (MOIST) dq[p] = fma2(c1q[p], x1[p], c0[p] * x0[p]) - Rq[p];
    }
    st8(dr, gN, nv, cv[0], cv[1]); st8(de, gN + 3 * cs, nv, cv[0], cv[1]);
    if (MOIST) st8(dq, gN + 4 * cs, nv, cv[0], cv[1]);
    st8(x0, gNf, nf, fv[0], fv[1]);
    return;
  }
  // ---- U ← U − ΔU (back-substitution of the scalar rows)
  V2 nr[2], nre[2], nu[2], nu1[2], nq[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int v = 2 * lane + p;
    nr[p] = R0[p] - fma2(a1[p], x1[p], a0[p] * x0[p]);
    nre[p] = E0[p] - fma2(b1[p], x1[p], b0[p] * x0[p]);
    if (MOIST) nq[p] = Q0s[p] - fma2(c1q[p], x1[p], c0[p] * x0[p]);
    nu[p] = interior[p] ? u3[p] - x0[p] : V2(FT(0));
    nu1[p] = (v + 1 < nv) ? u3h[p] - x1[p] : V2(FT(0));
  }
  st8(nr, gN, nv, cv[0], cv[1]);
  st8(nu, gNf, nf, fv[0], fv[1]);
  if (P.upwinding != 0) {
    // ---- h_tot of the updated state (cache_imp! after the Newton update), then the (upwinded − centred) enthalpy flux
    V2 hn[2], rn[2], qn[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      hn[p] = V2(FT(0)); rn[p] = cv[p] ? nr[p] : V2(FT(1)); qn[p] = V2(FT(0));
      if (cv[p]) {
        const V2 K = Kh[p] + (nu[p] * (nu[p] * g33lo[p]) + nu1[p] * (nu1[p] * g33hi[p])) * FT(0.25);
        if constexpr (MOIST) {
          Mst2<FT> m;
          hn[p] = thermo2m(P, nr[p], nre[p], nq[p], K, phi[p], m).h;
          qn[p] = div2(nq[p], nr[p]);
        } else {
          const V2 etot = nre[p] * rcpn2(nr[p]);
          const V2 T = max2(P.T_min_sgs, fma2(((etot - K) - phi[p]) + P.RT0, V2(P.icv), V2(P.T_0)));
          hn[p] = fma2(T, V2(P.R_d), etot);
        }
      }
    }
    const V2 hnu0 = shup(hn[0]), hnu1 = shup(hn[1]), hnd0 = shdn(hn[0]), rnu1 = shup(rn[1]);
    const V2 hn_m1[2] = {hnu1, hn[0]}, hn_m2[2] = {hnu0, hnu1}, hn_p1[2] = {hn[1], hnd0}, rn_m1[2] = {rnu1, rn[0]};
    V2 qn_m1[2], qn_m2[2], qn_p1[2];
    if (MOIST) {
      const V2 qu0 = shup(qn[0]), qu1 = shup(qn[1]), qd0 = shdn(qn[0]);
      qn_m1[0] = qu1; qn_m1[1] = qn[0]; qn_m2[0] = qu0; qn_m2[1] = qu1; qn_p1[0] = qn[1]; qn_p1[1] = qd0;
    }
    V2 flx[2], flq[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int v = 2 * lane + p;
      flx[p] = V2(FT(0)); flq[p] = V2(FT(0));
      if (interior[p]) {
        const V2 wv = nu[p] * g33lo[p];
        const V2 mr = fma2(nr[p], V2(mc[p]), rn_m1[p] * mclo[p]) * FT(0.5);
        flx[p] = (mr * wv) * upw_minus_central2(P, wv, hn_m2[p], hn_m1[p], hn[p], hn_p1[p], v, nv);
        if (MOIST) flq[p] = (mr * wv) * upw_minus_central2(P, wv, qn_m2[p], qn_m1[p], qn[p], qn_p1[p], v, nv);
      }
    }
    const V2 fp[2] = {flx[1], shdn(flx[0])};
#pragma unroll
    for (int p = 0; p < 2; ++p) nre[p] = nre[p] + ((-(fp[p] - flx[p])) * rmc[p]) * dtg;
    if (MOIST) {
      const V2 fq[2] = {flq[1], shdn(flq[0])};
#pragma unroll
      for (int p = 0; p < 2; ++p) nq[p] = nq[p] + ((-(fq[p] - flq[p])) * rmc[p]) * dtg;
    }
  }
  st8(nre, gN + 3 * cs, nv, cv[0], cv[1]);
  if (MOIST) st8(nq, gN + 4 * cs, nv, cv[0], cv[1]);
}

}  // namespace b200

namespace b200 {

// ---------------------------------------------------------------------------------------------
// Hook kernels in the warp-per-column-pair layout (the hook-by-hook path a ClimaTimeSteppers integration drives):
//   k8_t_imp       implicit_tendency! (implicit_tendency.jl:36-98,185-298; the diffusion part is added by k_vdiff_tend2)
//   k8_t_post_imp  correct_implicit_advection_tendency! (:322-339)
// Same arithmetic as k_t_imp2 / k_t_post_imp2 (kernels_vdiff.cuh: one point per thread, quarter element per CTA, 12 shared profiles),
// without shared memory and barriers; the state is used as it comes (cache_imp! has filtered u₃ on the boundary faces before).
template <class FT, int NVC, bool MOIST = false>
__global__ void __launch_bounds__(256, sizeof(FT) == 4 ? 2 : 1)
k8_t_imp(Par<FT> P, const FT* __restrict__ hgeo, const VLev<FT>* __restrict__ vlev, const FT* __restrict__ Yc, const FT* __restrict__ Yf,
         FT* __restrict__ Ytc, FT* __restrict__ Ytf) {
  using V2 = P2<FT>;
  const int e = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5, n0 = 2 * w, nv = NVC ? NVC : P.nv, nf = nv + 1;
  bool cv[2], fv[2], interior[2];
  FT sc2i[2], phi[2], mc[2], mclo[2], rmc[2], g33lo[2], g33hi[2], dphif[2], beta[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int v = 2 * lane + p;
    cv[p] = v < nv; fv[p] = v < nf; interior[p] = v > 0 && v < nv;
    const int vm = v > 0 ? v - 1 : 0;
    const int vc = cv[p] ? v : nv - 1, vmc = vm < nv ? vm : nv - 1, vf = fv[p] ? v : nv, vf1 = v + 1 <= nv ? v + 1 : nv;
    sc2i[p] = vlev->sc2i[vc]; phi[p] = vlev->phic[vc]; mc[p] = vlev->mc[vc]; mclo[p] = vlev->mc[vmc]; rmc[p] = vlev->rmc[vc];
    g33lo[p] = vlev->g33f[vf]; g33hi[p] = vlev->g33f[vf1]; dphif[p] = vlev->dphif[vf]; beta[p] = P.rayleigh ? vlev->brw[vf] : FT(0);
  }
  const FT* hgp = hgeo + (size_t)e * HG_N * 16 + n0;
  const V2 g11 = ldpair(hgp + HG_GI11 * 16), g12 = ldpair(hgp + HG_GI12 * 16), g22 = ldpair(hgp + HG_GI22 * 16);
  const int cs = 16 * nv;
  const FT* gY = Yc + ((size_t)e * P.ncf * 16 * nv + n0 * nv + 2 * lane);
  FT* gT = Ytc + ((size_t)e * P.ncf * 16 * nv + n0 * nv + 2 * lane);
  V2 rho[2], u1[2], u2[2], re[2], u3[2], rq[2];
  ld8(rho, gY, nv, cv[0], cv[1], FT(1)); ld8(u1, gY + cs, nv, cv[0], cv[1], FT(0)); ld8(u2, gY + 2 * cs, nv, cv[0], cv[1], FT(0));
  ld8(re, gY + 3 * cs, nv, cv[0], cv[1], FT(0)); ld8(u3, Yf + ((size_t)e * 16 * nf + n0 * nf + 2 * lane), nf, fv[0], fv[1], FT(0));
  if (MOIST) ld8(rq, gY + 4 * cs, nv, cv[0], cv[1], FT(0));
  const V2 zero[2] = {V2(FT(0)), V2(FT(0))};
  st8(zero, gT + cs, nv, cv[0], cv[1]); st8(zero, gT + 2 * cs, nv, cv[0], cv[1]);
  for (int q = MOIST ? 5 : 4; q < P.ncf; ++q) st8(zero, gT + q * cs, nv, cv[0], cv[1]);
  const V2 u3h[2] = {u3[1], shdn(u3[0])}, rlo[2] = {shup(rho[1]), rho[0]};
  V2 h[2], Pi[2], thp[2], phr[2], M[2], qv[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    h[p] = V2(FT(0)); Pi[p] = V2(FT(1)); thp[p] = V2(FT(0)); phr[p] = V2(FT(1)); qv[p] = V2(FT(0));
    if (cv[p]) {
      const V2 c1 = fma2(g12, u2[p], g11 * u1[p]), c2 = fma2(g22, u2[p], g12 * u1[p]);
      const V2 K = (fma2(u2[p], c2, u1[p] * c1) * sc2i[p]) * FT(0.5) + (u3[p] * (u3[p] * g33lo[p]) + u3h[p] * (u3h[p] * g33hi[p])) * FT(0.25);
      Pt2<FT> t;
      if constexpr (MOIST) { Mst2<FT> m; t = thermo2m(P, rho[p], re[p], rq[p], K, phi[p], m); qv[p] = div2(rq[p], rho[p]); }
      else t = thermo2(P, rho[p], re[p], K, phi[p]);
      h[p] = t.h; Pi[p] = t.Pi; thp[p] = t.thp; phr[p] = pgf_aux2(t);
    }
    M[p] = V2(FT(0));
    if (interior[p]) M[p] = (fma2(rho[p], V2(mc[p]), rlo[p] * mclo[p]) * FT(0.5)) * (u3[p] * g33lo[p]);  // ᶠinterp(ρJ) u³ / J2
  }
  const V2 h_m1[2] = {shup(h[1]), h[0]}, h_p1[2] = {h[1], shdn(h[0])}, M_p1[2] = {M[1], shdn(M[0])};
  const V2 Pi_m1[2] = {shup(Pi[1]), Pi[0]}, thp_m1[2] = {shup(thp[1]), thp[0]}, phr_m1[2] = {shup(phr[1]), phr[0]};
  V2 q_m1[2], q_p1[2];
  if (MOIST) { q_m1[0] = shup(qv[1]); q_m1[1] = qv[0]; q_p1[0] = qv[1]; q_p1[1] = shdn(qv[0]); }
  V2 rt[2], et[2], qt[2], tf[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int v = 2 * lane + p;
    const V2 hf0 = v > 0 ? (h_m1[p] + h[p]) * FT(0.5) : V2(FT(0)), hfp = v < nv - 1 ? (h[p] + h_p1[p]) * FT(0.5) : V2(FT(0));
    const V2 Mp = M_p1[p];
    rt[p] = -((Mp - M[p]) * rmc[p]);
    et[p] = -((Mp * hfp - M[p] * hf0) * rmc[p]);
    if constexpr (MOIST) {
      const V2 qf0 = v > 0 ? (q_m1[p] + qv[p]) * FT(0.5) : V2(FT(0)), qfp = v < nv - 1 ? (qv[p] + q_p1[p]) * FT(0.5) : V2(FT(0));
      qt[p] = -((Mp * qfp - M[p] * qf0) * rmc[p]);
    }
    tf[p] = V2(FT(0));
    if (interior[p]) {
      V2 dPi, dphr;
      pgf_diff2(P, Pi_m1[p], Pi[p], phr_m1[p], phr[p], dPi, dphr);
      tf[p] = -((V2(dphif[p]) - dphr) + (((thp_m1[p] + thp[p]) * FT(0.5)) * P.cp_d) * dPi);
    }
    if (P.rayleigh) tf[p] = tf[p] - u3[p] * beta[p];
  }
  st8(rt, gT, nv, cv[0], cv[1]); st8(et, gT + 3 * cs, nv, cv[0], cv[1]);
  if (MOIST) st8(qt, gT + 4 * cs, nv, cv[0], cv[1]);
  st8(tf, Ytf + ((size_t)e * 16 * nf + n0 * nf + 2 * lane), nf, fv[0], fv[1]);
}

template <class FT, int NVC, bool MOIST = false>
__global__ void __launch_bounds__(256, sizeof(FT) == 4 ? 2 : 1)
k8_t_post_imp(Par<FT> P, const FT* __restrict__ hgeo, const VLev<FT>* __restrict__ vlev, const FT* __restrict__ Yc, const FT* __restrict__ Yf,
              FT* __restrict__ Ytc, FT* __restrict__ Ytf) {
  using V2 = P2<FT>;
  const int e = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5, n0 = 2 * w, nv = NVC ? NVC : P.nv, nf = nv + 1;
  bool cv[2], fv[2], interior[2];
  FT sc2i[2], phi[2], mc[2], mclo[2], rmc[2], g33lo[2], g33hi[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int v = 2 * lane + p;
    cv[p] = v < nv; fv[p] = v < nf; interior[p] = v > 0 && v < nv;
    const int vm = v > 0 ? v - 1 : 0;
    const int vc = cv[p] ? v : nv - 1, vmc = vm < nv ? vm : nv - 1, vf = fv[p] ? v : nv, vf1 = v + 1 <= nv ? v + 1 : nv;
    sc2i[p] = vlev->sc2i[vc]; phi[p] = vlev->phic[vc]; mc[p] = vlev->mc[vc]; mclo[p] = vlev->mc[vmc]; rmc[p] = vlev->rmc[vc];
    g33lo[p] = vlev->g33f[vf]; g33hi[p] = vlev->g33f[vf1];
  }
  const FT* hgp = hgeo + (size_t)e * HG_N * 16 + n0;
  const V2 g11 = ldpair(hgp + HG_GI11 * 16), g12 = ldpair(hgp + HG_GI12 * 16), g22 = ldpair(hgp + HG_GI22 * 16);
  const int cs = 16 * nv;
  const FT* gY = Yc + ((size_t)e * P.ncf * 16 * nv + n0 * nv + 2 * lane);
  FT* gT = Ytc + ((size_t)e * P.ncf * 16 * nv + n0 * nv + 2 * lane);
  V2 rho[2], u1[2], u2[2], re[2], u3[2], rq[2];
  ld8(rho, gY, nv, cv[0], cv[1], FT(1)); ld8(u1, gY + cs, nv, cv[0], cv[1], FT(0)); ld8(u2, gY + 2 * cs, nv, cv[0], cv[1], FT(0));
  ld8(re, gY + 3 * cs, nv, cv[0], cv[1], FT(0)); ld8(u3, Yf + ((size_t)e * 16 * nf + n0 * nf + 2 * lane), nf, fv[0], fv[1], FT(0));
  if (MOIST) ld8(rq, gY + 4 * cs, nv, cv[0], cv[1], FT(0));
  const V2 zero[2] = {V2(FT(0)), V2(FT(0))};
  st8(zero, gT, nv, cv[0], cv[1]); st8(zero, gT + cs, nv, cv[0], cv[1]); st8(zero, gT + 2 * cs, nv, cv[0], cv[1]);
  for (int q = MOIST ? 5 : 4; q < P.ncf; ++q) st8(zero, gT + q * cs, nv, cv[0], cv[1]);
  st8(zero, Ytf + ((size_t)e * 16 * nf + n0 * nf + 2 * lane), nf, fv[0], fv[1]);
  const V2 u3h[2] = {u3[1], shdn(u3[0])};
  V2 h[2], qv[2], rn[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    h[p] = V2(FT(0)); qv[p] = V2(FT(0)); rn[p] = cv[p] ? rho[p] : V2(FT(1));
    if (cv[p]) {
      const V2 c1 = fma2(g12, u2[p], g11 * u1[p]), c2 = fma2(g22, u2[p], g12 * u1[p]);
      const V2 K = (fma2(u2[p], c2, u1[p] * c1) * sc2i[p]) * FT(0.5) + (u3[p] * (u3[p] * g33lo[p]) + u3h[p] * (u3h[p] * g33hi[p])) * FT(0.25);
      if constexpr (MOIST) { Mst2<FT> m; h[p] = thermo2m(P, rho[p], re[p], rq[p], K, phi[p], m).h; qv[p] = div2(rq[p], rho[p]); }
      else h[p] = thermo2(P, rho[p], re[p], K, phi[p]).h;
    }
  }
  const V2 hu0 = shup(h[0]), hu1 = shup(h[1]), hd0 = shdn(h[0]), rnu1 = shup(rn[1]);
  const V2 h_m1[2] = {hu1, h[0]}, h_m2[2] = {hu0, hu1}, h_p1[2] = {h[1], hd0}, rn_m1[2] = {rnu1, rn[0]};
  V2 q_m1[2], q_m2[2], q_p1[2];
  if (MOIST) {
    const V2 qu0 = shup(qv[0]), qu1 = shup(qv[1]), qd0 = shdn(qv[0]);
    q_m1[0] = qu1; q_m1[1] = qv[0]; q_m2[0] = qu0; q_m2[1] = qu1; q_p1[0] = qv[1]; q_p1[1] = qd0;
  }
  V2 flx[2], flq[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int v = 2 * lane + p;
    flx[p] = V2(FT(0)); flq[p] = V2(FT(0));
    if (interior[p]) {
      const V2 wv = u3[p] * g33lo[p];
      const V2 mr = fma2(rho[p], V2(mc[p]), rn_m1[p] * mclo[p]) * FT(0.5);
      flx[p] = (mr * wv) * upw_minus_central2(P, wv, h_m2[p], h_m1[p], h[p], h_p1[p], v, nv);
      if (MOIST) flq[p] = (mr * wv) * upw_minus_central2(P, wv, q_m2[p], q_m1[p], qv[p], q_p1[p], v, nv);
    }
  }
  const V2 fp[2] = {flx[1], shdn(flx[0])};
  V2 et[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) et[p] = (-(fp[p] - flx[p])) * rmc[p];
  st8(et, gT + 3 * cs, nv, cv[0], cv[1]);
  if (MOIST) {
    const V2 fq[2] = {flq[1], shdn(flq[0])};
    V2 qt[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) qt[p] = (-(fq[p] - flq[p])) * rmc[p];
    st8(qt, gT + 4 * cs, nv, cv[0], cv[1]);
  }
}

}  // namespace b200
