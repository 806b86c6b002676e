// kernels_imp5.cuh — fused implicit stage, third generation (row layout, packed two-lane arithmetic).
//
// Same arithmetic as k2_imp_stage (cache_imp! → Wfact → T_imp! residual → ldiv! → U −= ΔU → cache_imp! →
// T_post_imp!; implicit_tendency.jl:36-98,185-339, manual_sparse_jacobian.jl:713-870,504-585), restructured after the
// SASS/ncu reading of k2_imp_stage (4008 instructions per thread for 4 points: 1100 integer/address, 519 LDS/STS,
// 350 branch instructions; 62 % issue-active, barrier stall 3.8 per issue):
//   * thread (v, j) owns FOUR columns (nodes 4j..4j+3) at level v as two f32x2 pairs: the per-point algebra is
//     FFMA2/FMUL2/FADD2, index arithmetic and level constants are paid once per four points;
//   * the implicit stage has no horizontal coupling, so a warp is 32 consecutive levels of one row: global
//     accesses are 128-byte lines, vertical neighbours (v±1, v±2) are 64-bit shared loads of a pair slab
//     s[jp][v] (jp = 2j + p, stride 65 pairs: conflict-free for level-major and for column-sequential access);
//   * level constants and the three metric terms come straight from global memory (L1 broadcast);
//   * the 16 Schur tridiagonal systems are solved by PARALLEL CYCLIC REDUCTION in the same thread layout (six
//     steps of 2^k-strided eliminations on normalised rows, every thread busy) — a one-sided Thomas sweep by 16
//     lanes cost ≈60 µs of the 220 µs kernel and a two-sided one ≈30 µs (measured in round 1, profiles/r1_ncu_summary.md).
// 13 pair slabs (54 KB Float32) ⇒ 4 CTAs/SM.
#pragma once
#include "common.cuh"
#include "kernels_implicit.cuh"
#include "kernels_row.cuh"
#include "pair.cuh"
#include "thermo2.cuh"
#include "moist2.cuh"

namespace b200 {

constexpr int PLV = 65;          // pair-slab level stride (in pairs)
constexpr int PSLAB = 8 * PLV;   // pairs per slab (8 pair-columns × 65)
constexpr int IMP5_SLABS = 13, IMP5_SLABS_MOIST = 16;
template <class FT> constexpr size_t smem_imp5(bool moist = false) { return (size_t)(moist ? IMP5_SLABS_MOIST : IMP5_SLABS) * PSLAB * sizeof(P2<FT>); }

// van Leer limited slope (same value as vl_slope in kernels_implicit.cuh, written with min/max instructions)
template <class FT>
__device__ __forceinline__ FT vl_slope5(FT am, FT a0, FT ap) {
  const FT d = ((a0 - am) + (ap - a0)) / FT(2);
  const FT mn = mn_(mn_(am, a0), ap), mx = mx_(mx_(am, a0), ap);
  const FT lim = mn_(abs_(d), mn_(FT(2) * (a0 - mn), FT(2) * (mx - a0)));
  return d > FT(0) ? lim : (d < FT(0) ? -lim : FT(0));
}

// (upwinded − centred) face value per lane for the post-Newton correction (implicit_tendency.jl:322-339)
template <class FT>
__device__ __forceinline__ P2<FT> upw_minus_central2(const Par<FT>& P, P2<FT> w, P2<FT> am2, P2<FT> am, P2<FT> ap, P2<FT> ap2, int v, int nv) {
  FT d[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const FT wk = k ? w.hi() : w.lo(), a_m2 = k ? am2.hi() : am2.lo(), a_m = k ? am.hi() : am.lo(), a_p = k ? ap.hi() : ap.lo(),
             a_p2 = k ? ap2.hi() : ap2.lo();
    // the w < 0 branch is the mirror image of the w ≥ 0 one (slope(a,b,c) = −slope(c,b,a) exactly)
    const bool pos = wk >= FT(0);
    const FT y0 = pos ? a_m2 : a_p2, y1 = pos ? a_m : a_p, y2 = pos ? a_p : a_m;
    FT upv = y1;
    if (P.upwinding == 3 && v >= 2 && v <= nv - 2) upv = y1 + vl_slope5(y0, y1, y2) / FT(2) * (FT(1) - abs_(wk) * P.dt);
    else if (P.upwinding == 2 && nv >= 3) upv = upwind3_face(a_m2, a_m, a_p, a_p2, v, nv, wk);  // ᶠupwind3 (third_order)
    d[k] = upv - FT(0.5) * (a_m + a_p);
  }
  return P2<FT>(d[0], d[1]);
}

// g already points at (node n0, level v) of the thread; the four nodes are nlev apart (uniform offsets)
template <class FT>
__device__ __forceinline__ void ld2g(P2<FT> (&a)[2], const FT* __restrict__ g, int nlev, bool ok, FT dflt) {
  FT t[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) t[i] = ok ? g[i * nlev] : dflt;
  a[0] = P2<FT>(t[0], t[1]); a[1] = P2<FT>(t[2], t[3]);
}
template <class FT>
__device__ __forceinline__ void st2g(const P2<FT> (&a)[2], FT* __restrict__ g, int nlev) {
  g[0] = a[0].lo(); g[nlev] = a[0].hi(); g[2 * nlev] = a[1].lo(); g[3 * nlev] = a[1].hi();
}

// NVC: compile-time number of levels (63 for every production configuration: all global offsets become immediates); 0 = run-time
// LDIV = false: the fused implicit stage  N = U − J(U)⁻¹·dtγ·T_imp(U) [+ T_post_imp! correction]  (b200_implicit_stage, the fused stepper).
// LDIV = true:  ldiv!(ΔY, J, R) of the hook path (jacobian.jl:78-82) with the SAME coefficient code: (Yc, Yf) is the snapshot of the
//               state that Wfact was called with (b200_wfact keeps S bytes instead of writing 15 coefficient planes), (Rc, Rf) the
//               right-hand side, (Nc, Nf) receive ΔY.  Differences to the stage: the residuals come from R instead of T_imp, the
//               (u₃, uₕ) bidiagonal blocks enter the Schur right-hand side (R_uₕ = 0 in the stage), Δuₕ = −R_uₕ, Δ(ρχ) = −R_ρχ.
// MOIST (microphysics_model 0M): component 4 of Y.c is the active ρq_tot — moist thermodynamic state (moist.cuh), κ_m = R_m/cv_m per
//               point, the (ρq_tot, u₃) and (u₃, ρq_tot) blocks (manual_sparse_jacobian.jl:770-790, 827-831; A₁₁ stays −I, so the
//               ApproximateBlockArrowheadIterativeSolve of :538-578 is the exact arrowhead solve with one more rank-one term in the
//               Schur tridiagonal), central transport of q_tot (implicit_tendency.jl:210-214) and its post-Newton correction.
template <class FT, int NVC, bool LDIV = false, bool MOIST = false>
__global__ void __launch_bounds__(256, (MOIST && sizeof(FT) == 8) ? 1 : 2)
k5_imp_stage(Par<FT> P, const FT* __restrict__ hgeo, const VLev<FT>* __restrict__ vlev, const FT* __restrict__ Yc,
             const FT* __restrict__ Yf, FT* __restrict__ Nc, FT* __restrict__ Nf, FT dtg, const FT* __restrict__ Rc = nullptr,
             const FT* __restrict__ Rf = nullptr) {
  using V2 = P2<FT>;
  pdl_launch();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  V2* sb = reinterpret_cast<V2*>(smem_raw);
  V2 *s_rho = sb, *s_u3 = sb + PSLAB, *s_h = sb + 2 * PSLAB, *s_A = sb + 3 * PSLAB, *s_M = sb + 4 * PSLAB,
     *s_dp = sb + 5 * PSLAB, *s_Pi = sb + 6 * PSLAB, *s_thv = sb + 7 * PSLAB, *s_thp = sb + 8 * PSLAB,
     *s_phr = sb + 9 * PSLAB, *s_d = sb + 10 * PSLAB, *s_u = sb + 11 * PSLAB, *s_r = sb + 12 * PSLAB;
  V2 *s_kap = sb + 13 * PSLAB, *s_dpq = sb + 14 * PSLAB, *s_q = sb + 15 * PSLAB;  // MOIST only
  constexpr int Q0 = MOIST ? 5 : 4;  // first passive tracer
  const int e = blockIdx.x, v = threadIdx.x & 63, j = threadIdx.x >> 6, n0 = j * 4, nv = NVC ? NVC : P.nv, nf = nv + 1;
  const bool cv = v < nv, fv = v < nf, interior = v > 0 && v < nv;
  const int vm = v > 0 ? v - 1 : 0, vm2 = v > 1 ? v - 2 : 0, vp = v < LV - 1 ? v + 1 : v;
  const int o0 = (2 * j) * PLV + v;  // pair-slab offset of pair p: o0 + p·PLV
  const FT kap = P.R_d / P.cv_d;
  // level constants (centre v, face v) straight from global memory
  const int vc = cv ? v : nv - 1, vmc = vm < nv ? vm : nv - 1, vf = fv ? v : nv, vf1 = v + 1 <= nv ? v + 1 : nv;
  const FT sc2i = vlev->sc2i[vc], phi = vlev->phic[vc], mc = vlev->mc[vc], mclo = vlev->mc[vmc], rmc = vlev->rmc[vc],
           rmclo = vlev->rmc[vmc], g33lo = vlev->g33f[vf], g33hi = vlev->g33f[vf1], g33m = vlev->g33f[vm],
           dphif = vlev->dphif[vf], beta = P.rayleigh ? vlev->brw[vf] : FT(0);
  pdl_wait(Yc, Yf, Nc, Nf, Rc, Rf);
  const int cs = 16 * nv;  // component stride of Y.c
  const FT* gY = Yc + ((size_t)e * P.ncf * 16 * nv + n0 * nv + v);   // (ρ, node n0, level v) of this thread
  const FT* gYf = Yf + ((size_t)e * 16 * nf + n0 * nf + v);
  FT* gN = Nc + ((size_t)e * P.ncf * 16 * nv + n0 * nv + v);
  FT* gNf = Nf + ((size_t)e * 16 * nf + n0 * nf + v);
  V2 rho[2], u1[2], u2[2], re[2], u3[2];
  ld2g(rho, gY, nv, cv, FT(1)); ld2g(u1, gY + cs, nv, cv, FT(0)); ld2g(u2, gY + 2 * cs, nv, cv, FT(0));
  ld2g(re, gY + 3 * cs, nv, cv, FT(0)); ld2g(u3, gYf, nf, interior, FT(0));  // u₃ boundary filter on load
  V2 rq[2], Rq[2], Rq_lo[2];
  if (MOIST) ld2g(rq, gY + 4 * cs, nv, cv, FT(0));
#pragma unroll
  for (int p = 0; p < 2; ++p) { s_rho[o0 + p * PLV] = rho[p]; s_u3[o0 + p * PLV] = u3[p]; }
  // ldiv!: the right-hand side at this level and the level below (second load: an L1 hit), and the state of the level below
  V2 Rr[2], R1[2], R2[2], Re[2], R3[2], Rr_lo[2], R1_lo[2], R2_lo[2], Re_lo[2], u1_lo[2], u2_lo[2];
  if (LDIV) {
    const FT* gR = Rc + ((size_t)e * P.ncf * 16 * nv + n0 * nv + v);
    ld2g(Rr, gR, nv, cv, FT(0)); ld2g(R1, gR + cs, nv, cv, FT(0)); ld2g(R2, gR + 2 * cs, nv, cv, FT(0)); ld2g(Re, gR + 3 * cs, nv, cv, FT(0));
    ld2g(R3, Rf + ((size_t)e * 16 * nf + n0 * nf + v), nf, fv, FT(0));
    ld2g(Rr_lo, gR - 1, nv, interior, FT(0)); ld2g(R1_lo, gR + cs - 1, nv, interior, FT(0));
    ld2g(R2_lo, gR + 2 * cs - 1, nv, interior, FT(0)); ld2g(Re_lo, gR + 3 * cs - 1, nv, interior, FT(0));
    ld2g(u1_lo, gY + cs - 1, nv, interior, FT(0)); ld2g(u2_lo, gY + 2 * cs - 1, nv, interior, FT(0));
    if (MOIST) { ld2g(Rq, gR + 4 * cs, nv, cv, FT(0)); ld2g(Rq_lo, gR + 4 * cs - 1, nv, interior, FT(0)); }
    if (cv) {  // Δuₕ = −R_uₕ ((uₕ,uₕ) = −I), Δ(ρχ) = −R_ρχ (passive tracers: the fallback −I block, manual_sparse_jacobian.jl:476-481)
      V2 m1[2] = {-R1[0], -R1[1]}, m2[2] = {-R2[0], -R2[1]};
      st2g(m1, gN + cs, nv); st2g(m2, gN + 2 * cs, nv);
      for (int q = Q0; q < P.ncf; ++q) {
        V2 t[2];
        ld2g(t, gR + q * cs, nv, true, FT(0));
        V2 m[2] = {-t[0], -t[1]};
        st2g(m, gN + q * cs, nv);
      }
    }
  } else if (cv) {  // uₕ is copied through (R_uₕ = 0); passive tracers: ΔU = 0
    st2g(u1, gN + cs, nv); st2g(u2, gN + 2 * cs, nv);
    for (int q = Q0; q < P.ncf; ++q) {
      V2 t[2];
      ld2g(t, gY + q * cs, nv, true, FT(0));
      st2g(t, gN + q * cs, nv);
    }
  }
  const FT* hgp = hgeo + (size_t)e * HG_N * 16 + n0;
  V2 Kh[2], ck1[2], ck2[2], ck1_lo[2], ck2_lo[2];  // ck: ∂K/∂uₕ = CT12(uₕ) at this centre and the one below (ldiv! only)
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const V2 g11 = ldpair(hgp + HG_GI11 * 16 + 2 * p), g12 = ldpair(hgp + HG_GI12 * 16 + 2 * p), g22 = ldpair(hgp + HG_GI22 * 16 + 2 * p);
    const V2 c1 = fma2(g12, u2[p], g11 * u1[p]), c2 = fma2(g22, u2[p], g12 * u1[p]);
    Kh[p] = (fma2(u2[p], c2, u1[p] * c1) * sc2i) * FT(0.5);
    if (LDIV) {
      const FT sc2i_lo = vlev->sc2i[vmc];
      ck1[p] = c1 * sc2i; ck2[p] = c2 * sc2i;
      ck1_lo[p] = fma2(g12, u2_lo[p], g11 * u1_lo[p]) * sc2i_lo; ck2_lo[p] = fma2(g22, u2_lo[p], g12 * u1_lo[p]) * sc2i_lo;
    }
  }
  __syncthreads();  // (1) ρ, u₃ slabs
  // ---- centre thermodynamics (level v) and face mass-flux pieces (face v)
  V2 u3h[2], rlo[2], h[2], A[2], M[2], kapv[2], qv[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    u3h[p] = s_u3[o0 + p * PLV - v + vp];
    rlo[p] = s_rho[o0 + p * PLV - v + vm];
    V2 Pi(FT(1)), thv(FT(0)), thp(FT(0)), phr(FT(1)), dp(FT(0)), dpq(FT(0));
    h[p] = V2(FT(0)); kapv[p] = V2(kap); qv[p] = V2(FT(0));
    if (cv) {
      const V2 K = Kh[p] + (u3[p] * (u3[p] * g33lo) + u3h[p] * (u3h[p] * g33hi)) * FT(0.25);
      Pt2<FT> t;
      if constexpr (MOIST) {
        Mst2<FT> m;
        t = thermo2m(P, rho[p], re[p], rq[p], K, phi, m);
        dpq = dp_drhoq2(P, m, kapv[p]);  // ᶜ∂p∂ρq_tot_field! (:670-690), ᶜkappa_m_field! (:653-662)
        qv[p] = div2(rq[p], rho[p]);
        dp = t.T * (V2(P.R_d) - kapv[p] * P.cv_d) + ((V2(P.T_0 * P.cp_d) - K) - phi) * kapv[p];
      } else {
        t = thermo2(P, rho[p], re[p], K, phi);
        // ∂p/∂ρ at fixed ρe_tot (manual_sparse_jacobian.jl:816-818)
        dp = fma2(t.T, V2(P.R_d - kap * P.cv_d), ((V2(P.T_0 * P.cp_d) - K) - phi) * kap);
      }
      h[p] = t.h; Pi = t.Pi; thv = t.thv; thp = t.thp; phr = pgf_aux2(t);  // Φ_r (Float64) or p (Float32), see thermo2.cuh
    }
    A[p] = M[p] = V2(FT(0));
    if (interior) {  // M = ᶠinterp(ρJ)u³/J2,  A = dtγ ᶠinterp(ρJ) g³³/J2
      const V2 mr = fma2(rho[p], V2(mc), rlo[p] * mclo) * FT(0.5);
      A[p] = (mr * dtg) * g33lo;
      M[p] = mr * (u3[p] * g33lo);
    }
    const int o = o0 + p * PLV;
    s_h[o] = h[p]; s_Pi[o] = Pi; s_thv[o] = thv; s_thp[o] = thp; s_phr[o] = phr; s_dp[o] = dp;
    s_A[o] = A[p]; s_M[o] = M[p];
    if (MOIST) { s_kap[o] = kapv[p]; s_dpq[o] = dpq; s_q[o] = qv[p]; }
  }
  __syncthreads();  // (2) thermodynamic and flux slabs
  // ---- Schur tridiagonal and right-hand side of face row v (manual_sparse_jacobian.jl:746-868)
  // kept across the solve for the back-substitution  ρ_new = R0 − a0·x[v] − a1·x[v+1],  ρe_new = E0 − b0·x[v] − b1·x[v+1]
  V2 R0[2], E0[2], a0[2], a1[2], b0[2], b1[2], cl[2], cd[2], cu[2], cr[2], Q0s[2], c0[2], c1[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int o = o0 + p * PLV, om = o - v + vm, om2 = o - v + vm2, op = o - v + vp;
    const V2 hl = s_h[om], hm2 = s_h[om2], hp1 = s_h[op];
    const V2 hf0 = v > 0 ? (hl + h[p]) * FT(0.5) : V2(FT(0));        // ᶠinterp(h) at faces v, v+1
    const V2 hfp = v < nv - 1 ? (h[p] + hp1) * FT(0.5) : V2(FT(0));
    const V2 Ap = s_A[op], Mp = s_M[op];
    V2 qf0(FT(0)), qfp(FT(0)), qfm(FT(0));  // MOIST: ᶠinterp(q_tot) at faces v, v+1, v−1
    {
      const V2 rr = ((Mp - M[p]) * (-dtg)) * rmc, rre = ((Mp * hfp - M[p] * hf0) * (-dtg)) * rmc;
      a0[p] = A[p] * rmc; a1[p] = -(Ap * rmc);
      b0[p] = a0[p] * hf0; b1[p] = a1[p] * hfp;
      R0[p] = rho[p] + rr; E0[p] = re[p] + rre;
      if constexpr (MOIST) {
        const V2 ql = s_q[om], qm2 = s_q[om2], qp1 = s_q[op];
        qf0 = v > 0 ? (ql + qv[p]) * FT(0.5) : V2(FT(0));
        qfp = v < nv - 1 ? (qv[p] + qp1) * FT(0.5) : V2(FT(0));
        qfm = v > 1 ? (qm2 + ql) * FT(0.5) : V2(FT(0));
        c0[p] = a0[p] * qf0; c1[p] = a1[p] * qfp;
        Q0s[p] = rq[p] + ((Mp * qfp - M[p] * qf0) * (-dtg)) * rmc;
      }
    }
    cl[p] = cu[p] = V2(FT(0)); cd[p] = V2(dtg * (-beta) - FT(1));
    cr[p] = LDIV ? R3[p] : V2(FT(0));  // boundary rows of ldiv!: x = R₃/(−dtγβ − 1)
    if (interior) {
      const V2 hfm = v > 1 ? (hm2 + hl) * FT(0.5) : V2(FT(0));
      const V2 Am = s_A[om], Mm = s_M[om], u3m = s_u3[om];
      const V2 Pil = s_Pi[om], thvl = s_thv[om], thpl = s_thp[om], phrl = s_phr[om], dpl = s_dp[om];
      const V2 Pi = s_Pi[o], thv = s_thv[o], thp = s_thp[o], phr = s_phr[o], dp = s_dp[o];
      const V2 irf = rcpn2((rlo[p] + rho[p]) * FT(0.5));
      V2 dPi, dphr;
      pgf_diff2(P, Pil, Pi, phrl, phr, dPi, dphr);
      const V2 buoy = ((((thvl + thv) * FT(0.5)) * P.cp_d) * dPi) * irf;
      const V2 hb = buoy * FT(0.5);
      const V2 ur_lo = fma2(irf, dpl, hb) * dtg, ur_hi = (hb - irf * dp) * dtg;
      V2 ue_lo, ue_hi, x_lo, x_hi, uq_lo(FT(0)), uq_hi(FT(0));
      if constexpr (MOIST) {
        const V2 kl = s_kap[om];
        ue_lo = (irf * dtg) * kl; ue_hi = -((irf * dtg) * kapv[p]);
        x_lo = irf * (rlo[p] * (-kl)); x_hi = -(irf * (rho[p] * (-kapv[p])));
        uq_lo = (irf * dtg) * s_dpq[om]; uq_hi = -((irf * dtg) * s_dpq[o]);  // (u₃, ρq_tot): dtγ ᶠp_grad_matrix ⋅ Diag(∂p/∂ρq_tot)
      } else {
        ue_lo = (irf * dtg) * kap; ue_hi = -ue_lo;
        x_lo = irf * (rlo[p] * (-kap)); x_hi = -(irf * (rho[p] * (-kap)));
      }
      const V2 k0 = u3[p] * (FT(0.5) * g33lo);
      V2 l = (x_lo * (u3m * (FT(0.5) * g33m))) * dtg;
      V2 d = (fma2(x_hi, k0, x_lo * k0) - beta) * dtg - FT(1);
      V2 u = (x_hi * (u3h[p] * (FT(0.5) * g33hi))) * dtg;
      // centre rows v-1 ("a") and v ("b"): ru_lo = A[k]/m_c[k], ru_hi = −A[k+1]/m_c[k], eu = ru·ᶠinterp(h)
      const V2 ru_lo_a = Am * rmclo, ru_hi_a = -(A[p] * rmclo), ru_lo_b = a0[p], ru_hi_b = a1[p];
      l = l + fma2(ue_lo, ru_lo_a * hfm, ur_lo * ru_lo_a);
      d = d + (fma2(ur_hi, ru_lo_b, ur_lo * ru_hi_a) + fma2(ue_hi, ru_lo_b * hf0, ue_lo * (ru_hi_a * hf0)));
      u = u + fma2(ue_hi, ru_hi_b * hfp, ur_hi * ru_hi_b);
      // R = dtγ·T_imp(U): face part + couplings to the centre residuals of rows v-1 and v
      const V2 rr_a = ((M[p] - Mm) * (-dtg)) * rmclo, rr_b = ((Mp - M[p]) * (-dtg)) * rmc;
      const V2 Mh0 = M[p] * hf0;
      const V2 re_a = ((Mh0 - Mm * hfm) * (-dtg)) * rmclo, re_b = ((Mp * hfp - Mh0) * (-dtg)) * rmc;
      const V2 tf = -((V2(dphif) - dphr) + (((thpl + thp) * FT(0.5)) * P.cp_d) * dPi) - u3[p] * beta;
      V2 moist_rhs(FT(0));
      if constexpr (MOIST) {  // Schur terms of the ρq_tot column/row; qu = ru·ᶠinterp(q_tot)
        l = l + uq_lo * (ru_lo_a * qfm);
        d = d + fma2(uq_hi, ru_lo_b * qf0, uq_lo * (ru_hi_a * qf0));
        u = u + uq_hi * (ru_hi_b * qfp);
        const V2 Mq0 = M[p] * qf0;
        const V2 rq_a = ((Mq0 - Mm * qfm) * (-dtg)) * rmclo, rq_b = ((Mp * qfp - Mq0) * (-dtg)) * rmc;
        moist_rhs = LDIV ? fma2(uq_lo, Rq_lo[p], uq_hi * Rq[p]) : fma2(uq_lo, rq_a, uq_hi * rq_b);
      }
      cl[p] = l; cd[p] = d; cu[p] = u;
      if (LDIV) {  // Schur right-hand side R₃ + A₃ρ R_ρ + A₃e R_ρe + A₃uₕ R_uₕ  (A₃uₕ = dtγ·(−κρ/ᶠρ)·CT12(uₕ), manual_sparse_jacobian.jl:855-868)
        const V2 xl = x_lo * dtg, xh = x_hi * dtg;
        cr[p] = R3[p] + (fma2(ur_lo, Rr_lo[p], ur_hi * Rr[p]) + fma2(ue_lo, Re_lo[p], ue_hi * Re[p])) +
                (fma2(xl * ck1_lo[p], R1_lo[p], (xh * ck1[p]) * R1[p]) + fma2(xl * ck2_lo[p], R2_lo[p], (xh * ck2[p]) * R2[p]));
      } else
        cr[p] = fma2(tf, V2(dtg), fma2(ur_lo, rr_a, ur_hi * rr_b) + fma2(ue_lo, re_a, ue_hi * re_b));
      if (MOIST) cr[p] = cr[p] + moist_rhs;
    }
  }
  V2 x0[2], x1[2];  // ΔU.f.u₃ at faces v and v+1
  {
    // ---- parallel cyclic reduction on the normalised rows a·x[v−s] + x[v] + c·x[v+s] = y, s = 1, 2, 4, …: every
    // thread reduces its own four rows, no serial sweep.  Double-buffered slabs, one barrier per step.
    V2 a[2], c[2], y[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const V2 rd = rcpn2(cd[p]);
      a[p] = cl[p] * rd; c[p] = cu[p] * rd; y[p] = cr[p] * rd;
    }
    int buf = 0;
    for (int st = 1; st < nf; st <<= 1, buf ^= 1) {
      V2 *ba = buf ? s_Pi : s_d, *bc = buf ? s_thv : s_u, *by = buf ? s_thp : s_r;
#pragma unroll
      for (int p = 0; p < 2; ++p) { const int o = o0 + p * PLV; ba[o] = a[p]; bc[o] = c[p]; by[o] = y[p]; }
      __syncthreads();
      const bool hm = v >= st, hp = v + st < LV;
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int o = o0 + p * PLV;
        V2 am(FT(0)), cm(FT(0)), ym(FT(0)), ap(FT(0)), cp(FT(0)), yp(FT(0));
        if (hm) { am = ba[o - st]; cm = bc[o - st]; ym = by[o - st]; }
        if (hp) { ap = ba[o + st]; cp = bc[o + st]; yp = by[o + st]; }
        const V2 rd = rcpn2(V2(FT(1)) - fma2(c[p], ap, a[p] * cm));
        y[p] = (y[p] - fma2(c[p], yp, a[p] * ym)) * rd;
        a[p] = -((a[p] * am) * rd);
        c[p] = -((c[p] * cp) * rd);
      }
    }
    // publish x so that thread v can read x[v+1]; the buffer written here was last read two barriers ago
    V2* bx = buf ? s_Pi : s_d;
#pragma unroll
    for (int p = 0; p < 2; ++p) { x0[p] = y[p]; bx[o0 + p * PLV] = y[p]; }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < 2; ++p) x1[p] = bx[o0 + p * PLV - v + vp];
  }
  if (LDIV) {  // ΔY: Δu₃ = x, Δρ = A_ρ3 x − R_ρ, Δρe_tot = A_e3 x − R_ρe (back-substitution of the scalar rows; A₁₁ = −I)
    V2 dr[2], de[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      dr[p] = fma2(a1[p], x1[p], a0[p] * x0[p]) - Rr[p];
      de[p] = fma2(b1[p], x1[p], b0[p] * x0[p]) - Re[p];
    }
    if (cv) { st2g(dr, gN, nv); st2g(de, gN + 3 * cs, nv); }
    if (MOIST && cv) {
      V2 dq[2];
#pragma unroll
      for (int p = 0; p < 2; ++p) dq[p] = fma2(c1[p], x1[p], c0[p] * x0[p]) - Rq[p];
      st2g(dq, gN + 4 * cs, nv);
    }
    if (fv) st2g(x0, gNf, nf);
    return;
  }
  // ---- U ← U − ΔU (back-substitution of the scalar rows)
  V2 nr[2], nre[2], nu[2], nu1[2], nq[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    nr[p] = R0[p] - fma2(a1[p], x1[p], a0[p] * x0[p]);
    nre[p] = E0[p] - fma2(b1[p], x1[p], b0[p] * x0[p]);
    if (MOIST) nq[p] = Q0s[p] - fma2(c1[p], x1[p], c0[p] * x0[p]);
    nu[p] = interior ? u3[p] - x0[p] : V2(FT(0));
    nu1[p] = (v + 1 < nv) ? u3h[p] - x1[p] : V2(FT(0));
  }
  if (cv) st2g(nr, gN, nv);
  if (fv) st2g(nu, gNf, nf);
  if (P.upwinding != 0) {
    // ---- h_tot of the updated state (cache_imp! after the Newton update), then the (upwinded − centred) enthalpy flux
    V2 hn[2], qn[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      hn[p] = V2(FT(0)); qn[p] = V2(FT(0));
      if (cv) {
        const V2 K = Kh[p] + (nu[p] * (nu[p] * g33lo) + nu1[p] * (nu1[p] * g33hi)) * FT(0.25);
        if constexpr (MOIST) {
          Mst2<FT> m;
          hn[p] = thermo2m(P, nr[p], nre[p], nq[p], K, phi, m).h;
          qn[p] = div2(nq[p], nr[p]);
        } else {
          const V2 etot = nre[p] * rcpn2(nr[p]);
          const V2 T = max2(P.T_min_sgs, fma2(((etot - K) - phi) + P.RT0, V2(P.icv), V2(P.T_0)));
          hn[p] = fma2(T, V2(P.R_d), etot);
        }
      }
      s_h[o0 + p * PLV] = hn[p]; s_rho[o0 + p * PLV] = cv ? nr[p] : V2(FT(1));
      if (MOIST) s_q[o0 + p * PLV] = qn[p];
    }
    __syncthreads();  // (5)
    V2 flx[2], flq[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      flx[p] = V2(FT(0)); flq[p] = V2(FT(0));
      if (interior) {
        const int o = o0 + p * PLV, om = o - 1, om2 = o - v + vm2, op = o - v + vp;
        const V2 w = nu[p] * g33lo;
        const V2 mr = fma2(nr[p], V2(mc), s_rho[om] * mclo) * FT(0.5);
        flx[p] = (mr * w) * upw_minus_central2(P, w, s_h[om2], s_h[om], hn[p], s_h[op], v, nv);
        if (MOIST) flq[p] = (mr * w) * upw_minus_central2(P, w, s_q[om2], s_q[om], qn[p], s_q[op], v, nv);
      }
      s_M[o0 + p * PLV] = flx[p];
      if (MOIST) s_A[o0 + p * PLV] = flq[p];
    }
    __syncthreads();  // (6)
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const V2 fp = s_M[o0 + p * PLV - v + vp];
      nre[p] = nre[p] + ((-(fp - flx[p])) * rmc) * dtg;
      if (MOIST) { const V2 fq = s_A[o0 + p * PLV - v + vp]; nq[p] = nq[p] + ((-(fq - flq[p])) * rmc) * dtg; }
    }
  }
  if (cv) st2g(nre, gN + 3 * cs, nv);
  if (MOIST && cv) st2g(nq, gN + 4 * cs, nv);
}

}  // namespace b200
