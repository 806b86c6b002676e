// kernels_imp8d.cuh — fused implicit stage WITH implicit vertical diffusion in the warp-per-column-pair layout of kernels_imp8.cuh.
//
// Same algebra as k_imp_stage_diff (kernels_vdiff.cuh; cache_imp! → Wfact incl. update_diffusion_jacobian! → R = dtγ·T_imp(U) incl. the
// diffusion tendency → ldiv! = ApproximateBlockArrowheadIterativeSolve → N = U − ΔU → cache_imp! → T_post_imp!; reference:
// implicit_tendency.jl:36-98,185-339, vertical_diffusion_boundary_layer.jl:64-154, manual_sparse_jacobian.jl:713-870,1031-1261,538-578).
// Every tridiagonal system — (uₕ,uₕ) with two right-hand sides, A_ee = (ρe_tot,ρe_tot) (1 + n_iters + 1 solves), the preconditioner
// P of the Schur complement (1 + n_iters solves), the passive-tracer blocks — is solved inside the warp by warp_tridiag_n: no shared
// memory and no block barrier.  (An intermediate version in the packed shared-memory layout of k5_imp_stage — block-wide PCR, 68 block
// barriers, bound by the LDS/STS traffic of its PCR steps — measured ≈ 600 µs per launch at he30/ze63 and was removed; this one ≈ 310 µs.)
#pragma once
#include "kernels_imp8.cuh"
#include "kernels_vdiff.cuh"

namespace b200 {

// level v − 1, v + 1, v − 2 of the thread's two levels (lane 0 / 31 get their own values at the column ends: callers mask)
#define K8_M1(X) {shup((X)[1]), (X)[0]}
#define K8_P1(X) {(X)[1], shdn((X)[0])}
#define K8_M2(X) {shup((X)[0]), shup((X)[1])}

template <class FT, int NVC>
__global__ void __launch_bounds__(256, sizeof(FT) == 4 ? 2 : 1)
k8_imp_stage_diff(Par<FT> P, VDiff<FT> D, const FT* __restrict__ hgeo, const VLev<FT>* __restrict__ vlev, const FT* __restrict__ Yc,
                  const FT* __restrict__ Yf, FT* __restrict__ Nc, FT* __restrict__ Nf, FT dtg) {
  using V2 = P2<FT>;
  pdl_launch();
  const int e = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5, n0 = 2 * w, nv = NVC ? NVC : P.nv, nf = nv + 1;
  const FT kap = P.R_d / P.cv_d;
  bool cv[2], fv[2], interior[2], lo[2], hi[2];
  FT sc2i[2], phi[2], mc[2], mclo[2], rmc[2], rmclo[2], g33lo[2], g33hi[2], g33m[2], dphif[2], beta[2], wfac[2], is0[2], isl[2], ish[2], sc2i_lo[2], kdec[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int v = 2 * lane + p;
    cv[p] = v < nv; fv[p] = v < nf; interior[p] = v > 0 && v < nv; lo[p] = v > 0; hi[p] = v < nv - 1;
    const int vm = v > 0 ? v - 1 : 0;
    const int vc = cv[p] ? v : nv - 1, vmc = vm < nv ? vm : nv - 1, vpc = v + 1 < nv ? v + 1 : nv - 1, vf = fv[p] ? v : nv, vf1 = v + 1 <= nv ? v + 1 : nv;
    sc2i[p] = vlev->sc2i[vc]; phi[p] = vlev->phic[vc]; mc[p] = vlev->mc[vc]; mclo[p] = vlev->mc[vmc]; rmc[p] = vlev->rmc[vc];
    rmclo[p] = vlev->rmc[vmc]; g33lo[p] = vlev->g33f[vf]; g33hi[p] = vlev->g33f[vf1]; g33m[p] = vlev->g33f[vm < nf ? vm : nv];
    dphif[p] = vlev->dphif[vf]; beta[p] = P.rayleigh ? vlev->brw[vf] : FT(0);
    wfac[p] = dtg * (vlev->dzf[vf] * vlev->g33f[vf] / vlev->sf2i[vf]);
    is0[p] = sqrt(sc2i[p]); isl[p] = sqrt(vlev->sc2i[vmc]); ish[p] = sqrt(vlev->sc2i[vpc]); sc2i_lo[p] = vlev->sc2i[vmc];
    kdec[p] = D.mode == 2 ? D.kdec[vc] : FT(0);
  }
  const FT* hgp = hgeo + (size_t)e * HG_N * 16 + n0;
  const V2 g11 = ldpair(hgp + HG_GI11 * 16), g12 = ldpair(hgp + HG_GI12 * 16), g22 = ldpair(hgp + HG_GI22 * 16);
  pdl_wait(Yc, Yf, Nc, Nf);
  const int cs = 16 * nv;
  const FT* gY = Yc + ((size_t)e * P.ncf * 16 * nv + n0 * nv + 2 * lane);
  const FT* gYf = Yf + ((size_t)e * 16 * nf + n0 * nf + 2 * lane);
  FT* gN = Nc + ((size_t)e * P.ncf * 16 * nv + n0 * nv + 2 * lane);
  FT* gNf = Nf + ((size_t)e * 16 * nf + n0 * nf + 2 * lane);
  V2 rho[2], u1[2], u2[2], re[2], u3[2];
  ld8(rho, gY, nv, cv[0], cv[1], FT(1)); ld8(u1, gY + cs, nv, cv[0], cv[1], FT(0)); ld8(u2, gY + 2 * cs, nv, cv[0], cv[1], FT(0));
  ld8(re, gY + 3 * cs, nv, cv[0], cv[1], FT(0)); ld8(u3, gYf, nf, interior[0], interior[1], FT(0));  // u₃ boundary filter on load
  // ---- centre thermodynamics, face mass-flux pieces, eddy diffusivity
  V2 h[2], Pi[2], thv[2], thp[2], phr[2], dp[2], sd[2], kh[2], A[2], M[2], ir[2];
  const V2 u3h[2] = K8_P1(u3), rlo[2] = K8_M1(rho), rhi[2] = K8_P1(rho);
  V2 ul1, ul2;  // uₕ at level 1 of the column (VerticalDiffusion)
  ul1 = V2(__shfl_sync(FULLM, u1[0].lo(), 0), __shfl_sync(FULLM, u1[0].hi(), 0));
  ul2 = V2(__shfl_sync(FULLM, u2[0].lo(), 0), __shfl_sync(FULLM, u2[0].hi(), 0));
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    h[p] = V2(FT(0)); Pi[p] = V2(FT(1)); thv[p] = V2(FT(0)); thp[p] = V2(FT(0)); phr[p] = V2(FT(1)); dp[p] = V2(FT(0)); sd[p] = V2(FT(0));
    kh[p] = V2(FT(1));
    ir[p] = rcpn2(rho[p]);
    if (cv[p]) {
      const V2 c1 = fma2(g12, u2[p], g11 * u1[p]), c2 = fma2(g22, u2[p], g12 * u1[p]);
      const V2 K = (fma2(u2[p], c2, u1[p] * c1) * sc2i[p]) * FT(0.5) + (u3[p] * (u3[p] * g33lo[p]) + u3h[p] * (u3h[p] * g33hi[p])) * FT(0.25);
      const Pt2<FT> t = thermo2(P, rho[p], re[p], K, phi[p]);
      h[p] = t.h; Pi[p] = t.Pi; thv[p] = t.thv; thp[p] = t.thp; phr[p] = pgf_aux2(t);
      dp[p] = fma2(t.T, V2(P.R_d - kap * P.cv_d), ((V2(P.T_0 * P.cp_d) - K) - phi[p]) * kap);
      sd[p] = fma2(t.T - P.T_0, V2(P.cp_d), V2(phi[p]));  // dry static energy cp_d (T − T_0) + Φ
      if (D.mode == 2) {
        kh[p] = V2(kdec[p]);
      } else {  // VerticalDiffusion: C_E |uₕ(level 1)| Δz₁/2 below 850 hPa, Gaussian taper in pressure above
        const V2 n2 = fma2(ul2, fma2(g22, ul2, g12 * ul1), ul1 * fma2(g12, ul2, g11 * ul1)) * vlev->sc2i[0];
        const V2 KE = V2(sqrt(n2.lo()), sqrt(n2.hi())) * D.ce_za;
        FT kk[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const FT pk = k ? t.p.hi() : t.p.lo(), ke = k ? KE.hi() : KE.lo();
          const FT x = (FT(85000) - pk) / FT(10000);
          kk[k] = pk > FT(85000) ? ke : ke * exp_(-(x * x));
        }
        kh[p] = V2(kk[0], kk[1]);
      }
    }
    A[p] = M[p] = V2(FT(0));
    if (interior[p]) {
      const V2 mr = fma2(rho[p], V2(mc[p]), rlo[p] * mclo[p]) * FT(0.5);
      A[p] = (mr * dtg) * g33lo[p];
      M[p] = mr * (u3[p] * g33lo[p]);
    }
  }
  // ---- dtγ (J g³³ ᶠρK)/J2 at face v
  const V2 kh_m1[2] = K8_M1(kh);
  V2 wl[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    wl[p] = V2(FT(0));
    if (interior[p]) {
      const V2 rf = (rlo[p] + rho[p]) * FT(0.5);
      const V2 ik = (rcpn2(max2(D.eps, kh_m1[p])) + rcpn2(max2(D.eps, kh[p]))) * FT(0.5);
      wl[p] = (rf * rcpn2(ik)) * wfac[p];
    }
  }
  const V2 w_p1[2] = K8_P1(wl), w_m1[2] = K8_M1(wl);
  const V2 h_m1[2] = K8_M1(h), h_m2[2] = K8_M2(h), h_p1[2] = K8_P1(h);
  const V2 A_m1[2] = K8_M1(A), A_p1[2] = K8_P1(A), M_m1[2] = K8_M1(M), M_p1[2] = K8_P1(M), u3_m1[2] = K8_M1(u3);
  const V2 Pi_m1[2] = K8_M1(Pi), thv_m1[2] = K8_M1(thv), thp_m1[2] = K8_M1(thp), phr_m1[2] = K8_M1(phr), dp_m1[2] = K8_M1(dp);
  const V2 sd_m1[2] = K8_M1(sd), sd_p1[2] = K8_P1(sd);
  const V2 u1_m1[2] = K8_M1(u1), u1_p1[2] = K8_P1(u1), u2_m1[2] = K8_M1(u2), u2_p1[2] = K8_P1(u2);
  // ---- residuals, centre-row and face-row coefficients, diffusion blocks
  V2 rr[2], rre[2], a0[2], a1[2], b0[2], b1[2], el[2], ed[2], eu[2], fc[2], pl[2], pd[2], pu[2], r12[2][2];
  V2 sl[2], sd_[2], su[2], uel[2], ueh[2], b3[2], Pl[2], Pd[2], Pu[2], xl_[2], xh_[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int v = 2 * lane + p;
    const V2 hl = h_m1[p], hm2 = h_m2[p], hp1 = h_p1[p];
    const V2 hf0 = v > 0 ? (hl + h[p]) * FT(0.5) : V2(FT(0));
    const V2 hfp = v < nv - 1 ? (h[p] + hp1) * FT(0.5) : V2(FT(0));
    const V2 Ap = A_p1[p], Mp = M_p1[p];
    const V2 wh = (cv[p] && hi[p]) ? w_p1[p] : V2(FT(0)), wlo = (cv[p] && lo[p]) ? wl[p] : V2(FT(0));
    const V2 irl = rcpn2(rlo[p]), irh = rcpn2(rhi[p]);
    rr[p] = ((Mp - M[p]) * (-dtg)) * rmc[p];
    a0[p] = A[p] * rmc[p]; a1[p] = -(Ap * rmc[p]);
    b0[p] = a0[p] * hf0; b1[p] = a1[p] * hfp;
    rre[p] = ((Mp * hfp - M[p] * hf0) * (-dtg)) * rmc[p];
    el[p] = eu[p] = pl[p] = pu[p] = V2(FT(0)); ed[p] = pd[p] = V2(FT(1)); fc[p] = V2(FT(0));
    r12[0][p] = r12[1][p] = V2(FT(0));
    if (cv[p]) {
      {  // dry static energy diffusion (vertical_diffusion_boundary_layer.jl:101-102)
        const V2 fl = lo[p] ? wlo * (sd[p] - sd_m1[p]) : V2(FT(0)), fh = hi[p] ? wh * (sd_p1[p] - sd[p]) : V2(FT(0));
        rre[p] = rre[p] + (fh - fl) * rmc[p];
      }
      if (D.momentum) {  // uₕ strain-rate form on uₕ/s_c (:91-96)
        const V2 sir = ir[p] * (rmc[p] / is0[p]);
        {
          const V2 c0 = u1[p] * is0[p];
          const V2 fl = lo[p] ? wlo * (c0 - u1_m1[p] * isl[p]) : V2(FT(0)), fh = hi[p] ? wh * (u1_p1[p] * ish[p] - c0) : V2(FT(0));
          r12[0][p] = (fh - fl) * sir;
        }
        {
          const V2 c0 = u2[p] * is0[p];
          const V2 fl = lo[p] ? wlo * (c0 - u2_m1[p] * isl[p]) : V2(FT(0)), fh = hi[p] ? wh * (u2_p1[p] * ish[p] - c0) : V2(FT(0));
          r12[1][p] = (fh - fl) * sir;
        }
      }
      const V2 l_ = lo[p] ? wlo * rmc[p] : V2(FT(0)), h_ = hi[p] ? wh * rmc[p] : V2(FT(0));
      const V2 dg = -(l_ + h_);
      const V2 m = dg * (ir[p] * D.cpcv);
      el[p] = lo[p] ? l_ * (irl * D.cpcv) : V2(FT(0));
      ed[p] = m - FT(1);
      eu[p] = hi[p] ? h_ * (irh * D.cpcv) : V2(FT(0));
      fc[p] = m * rcpn2(m - FT(1));
      pl[p] = l_ * ir[p]; pd[p] = dg * ir[p] - FT(1); pu[p] = h_ * ir[p];
    }
    sl[p] = su[p] = uel[p] = ueh[p] = xl_[p] = xh_[p] = V2(FT(0)); sd_[p] = V2(dtg * (-beta[p]) - FT(1)); b3[p] = V2(FT(0));
    Pl[p] = Pu[p] = V2(FT(0)); Pd[p] = sd_[p];
    if (interior[p]) {
      const V2 hfm = v > 1 ? (hm2 + hl) * FT(0.5) : V2(FT(0));
      const V2 Am = A_m1[p], Mm = M_m1[p], u3m = u3_m1[p];
      const V2 irf = rcpn2((rlo[p] + rho[p]) * FT(0.5));
      V2 dPi, dphr;
      pgf_diff2(P, Pi_m1[p], Pi[p], phr_m1[p], phr[p], dPi, dphr);
      const V2 buoy = ((((thv_m1[p] + thv[p]) * FT(0.5)) * P.cp_d) * dPi) * irf;
      const V2 hb = buoy * FT(0.5);
      const V2 ur_lo = fma2(irf, dp_m1[p], hb) * dtg, ur_hi = (hb - irf * dp[p]) * dtg;
      const V2 ue_lo = (irf * dtg) * kap, ue_hi = -ue_lo;
      const V2 x_lo = irf * (rlo[p] * (-kap)), x_hi = -(irf * (rho[p] * (-kap)));
      const V2 k0 = u3[p] * (FT(0.5) * g33lo[p]);
      V2 l = (x_lo * (u3m * (FT(0.5) * g33m[p]))) * dtg;
      V2 d = (fma2(x_hi, k0, x_lo * k0) - beta[p]) * dtg - FT(1);
      V2 u = (x_hi * (u3h[p] * (FT(0.5) * g33hi[p]))) * dtg;
      const V2 ru_lo_a = Am * rmclo[p], ru_hi_a = -(A[p] * rmclo[p]), ru_lo_b = a0[p], ru_hi_b = a1[p];
      const V2 eu_lo_a = ru_lo_a * hfm, eu_hi_a = ru_hi_a * hf0;
      l = l + fma2(ue_lo, eu_lo_a, ur_lo * ru_lo_a);
      d = d + (fma2(ur_hi, ru_lo_b, ur_lo * ru_hi_a) + fma2(ue_hi, b0[p], ue_lo * eu_hi_a));
      u = u + fma2(ue_hi, b1[p], ur_hi * ru_hi_b);
      sl[p] = l; sd_[p] = d; su[p] = u; uel[p] = ue_lo; ueh[p] = ue_hi;
      const V2 rr_a = ((M[p] - Mm) * (-dtg)) * rmclo[p];
      const V2 tf = -((V2(dphif[p]) - dphr) + (((thp_m1[p] + thp[p]) * FT(0.5)) * P.cp_d) * dPi) - u3[p] * beta[p];
      b3[p] = fma2(tf, V2(dtg), fma2(ur_lo, rr_a, ur_hi * rr[p]));
      // preconditioner: the Schur complement with A_ee replaced by its main diagonal, T − A₃e Diag(m/(m − 1)) A_e3  (m = d_ee + 1)
      const V2 l_lo = v > 1 ? w_m1[p] * rmclo[p] : V2(FT(0)), h_lo = wl[p] * rmclo[p];
      const V2 m_lo = (-(l_lo + h_lo)) * (irl * D.cpcv);
      const V2 fc_lo = m_lo * rcpn2(m_lo - FT(1));
      const V2 ca = ue_lo * fc_lo, cb = ue_hi * fc[p];
      Pl[p] = l - ca * eu_lo_a;
      Pd[p] = d - (ca * eu_hi_a + cb * b0[p]);
      Pu[p] = u - cb * b1[p];
      xl_[p] = x_lo * dtg; xh_[p] = x_hi * dtg;
    }
  }
  // ---- Δuₕ: exact tridiagonal solves (two right-hand sides), or −R = 0 without momentum diffusion
  if (D.momentum) warp_tridiag_n<FT, 2>(lane, pl, pd, pu, r12);
  // ---- y_e = A_ee⁻¹ R_ρe
  V2 ye[1][2] = {{cv[0] ? rre[0] : V2(FT(0)), cv[1] ? rre[1] : V2(FT(0))}};
  warp_tridiag_n<FT, 1>(lane, el, ed, eu, ye);
  // ---- Schur right-hand side, x₃ = P⁻¹ b
  V2 x3[1][2];
  {
    const V2 ye_m1[2] = K8_M1(ye[0]), r1_m1[2] = K8_M1(r12[0]), r2_m1[2] = K8_M1(r12[1]);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      if (interior[p]) {
        b3[p] = b3[p] - fma2(uel[p], ye_m1[p], ueh[p] * ye[0][p]);
        if (D.momentum) {  // − A₃uₕ Δuₕ: (u₃, uₕ) = dtγ ᶠp_grad ⋅ Diag(−κρ) ⋅ CT12(uₕ) at centres v − 1 and v (manual_sparse_jacobian.jl:855-868)
          const V2 ck1 = fma2(g12, u2[p], g11 * u1[p]) * sc2i[p], ck2 = fma2(g22, u2[p], g12 * u1[p]) * sc2i[p];
          const V2 ck1l = fma2(g12, u2_m1[p], g11 * u1_m1[p]) * sc2i_lo[p], ck2l = fma2(g22, u2_m1[p], g12 * u1_m1[p]) * sc2i_lo[p];
          b3[p] = b3[p] - (fma2(xl_[p] * ck1l, r1_m1[p], (xh_[p] * ck1) * r12[0][p]) + fma2(xl_[p] * ck2l, r2_m1[p], (xh_[p] * ck2) * r12[1][p]));
        }
      }
      x3[0][p] = b3[p];
    }
  }
  warp_tridiag_n<FT, 1>(lane, Pl, Pd, Pu, x3);
  // ---- stationary iteration x ← x + P⁻¹(b − S x),  S x = T x − A₃e A_ee⁻¹ A_e3 x  (A₃e(A_ee⁻¹ + I) form as in k_ldiv_diff)
  for (int it = 0; it < D.n_iters; ++it) {
    const V2 x_p1[2] = K8_P1(x3[0]), x_m1[2] = K8_M1(x3[0]);
    V2 z[1][2], yv[2], zy[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      yv[p] = cv[p] ? fma2(b1[p], x_p1[p], b0[p] * x3[0][p]) : V2(FT(0));
      z[0][p] = yv[p];
    }
    warp_tridiag_n<FT, 1>(lane, el, ed, eu, z);
#pragma unroll
    for (int p = 0; p < 2; ++p) zy[p] = z[0][p] + yv[p];
    const V2 zy_m1[2] = K8_M1(zy);
    V2 r3[1][2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int v = 2 * lane + p;
      V2 tx = sd_[p] * x3[0][p];
      if (v > 0) tx = fma2(sl[p], x_m1[p], tx);
      if (v < nv) tx = fma2(su[p], x_p1[p], tx);
      V2 r = b3[p] - tx;
      if (interior[p]) r = r + fma2(uel[p], zy_m1[p], ueh[p] * zy[p]);
      r3[0][p] = fv[p] ? r : V2(FT(0));
    }
    warp_tridiag_n<FT, 1>(lane, Pl, Pd, Pu, r3);
#pragma unroll
    for (int p = 0; p < 2; ++p) x3[0][p] = x3[0][p] + r3[0][p];
  }
  // ---- Δρ, Δρe_tot = A_ee⁻¹(R_ρe − A_e3 x₃), the Newton update
  const V2 x1[2] = K8_P1(x3[0]);
  V2 dre[1][2], nr[2], nre[2], nu[2], nu1[2], n1[2], n2[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) dre[0][p] = cv[p] ? rre[p] - fma2(b1[p], x1[p], b0[p] * x3[0][p]) : V2(FT(0));
  warp_tridiag_n<FT, 1>(lane, el, ed, eu, dre);
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int v = 2 * lane + p;
    nr[p] = (rho[p] + rr[p]) - fma2(a1[p], x1[p], a0[p] * x3[0][p]);
    nre[p] = re[p] - dre[0][p];
    n1[p] = u1[p] - r12[0][p]; n2[p] = u2[p] - r12[1][p];
    nu[p] = interior[p] ? u3[p] - x3[0][p] : V2(FT(0));
    nu1[p] = (v + 1 < nv) ? u3h[p] - x1[p] : V2(FT(0));
  }
  st8(nr, gN, nv, cv[0], cv[1]); st8(n1, gN + cs, nv, cv[0], cv[1]); st8(n2, gN + 2 * cs, nv, cv[0], cv[1]);
  st8(nu, gNf, nf, fv[0], fv[1]);
  // ---- passive tracers: Δ(ρχ) = (dtγ D ⋅ Diag(1/ρ) − I)⁻¹ dtγ D χ with the OLD 1/ρ
  for (int q = 4; q < P.ncf; ++q) {
    V2 rq[2], chi[2], tl[2], td[2], tu[2], zq[1][2];
    ld8(rq, gY + q * cs, nv, cv[0], cv[1], FT(0));
    chi[0] = rq[0] * ir[0]; chi[1] = rq[1] * ir[1];
    const V2 chi_m1[2] = K8_M1(chi), chi_p1[2] = K8_P1(chi);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      tl[p] = tu[p] = zq[0][p] = V2(FT(0)); td[p] = V2(FT(1));
      if (cv[p]) {
        const V2 wh = hi[p] ? w_p1[p] : V2(FT(0)), wlo = lo[p] ? wl[p] : V2(FT(0));
        const V2 l_ = wlo * rmc[p], h_ = wh * rmc[p];
        tl[p] = lo[p] ? l_ * rcpn2(rlo[p]) : V2(FT(0)); td[p] = -(l_ + h_) * ir[p] - FT(1); tu[p] = hi[p] ? h_ * rcpn2(rhi[p]) : V2(FT(0));
        const V2 fl = lo[p] ? wlo * (chi[p] - chi_m1[p]) : V2(FT(0)), fh = hi[p] ? wh * (chi_p1[p] - chi[p]) : V2(FT(0));
        zq[0][p] = (fh - fl) * rmc[p];
      }
    }
    warp_tridiag_n<FT, 1>(lane, tl, td, tu, zq);
    const V2 nq[2] = {rq[0] - zq[0][0], rq[1] - zq[0][1]};
    st8(nq, gN + q * cs, nv, cv[0], cv[1]);
  }
  // ---- cache_imp!(N) and T_post_imp!: (upwinded − centred) enthalpy flux of the updated state
  if (P.upwinding != 0) {
    V2 hn[2], rn[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      hn[p] = V2(FT(0)); rn[p] = cv[p] ? nr[p] : V2(FT(1));
      if (cv[p]) {
        const V2 c1 = fma2(g12, n2[p], g11 * n1[p]), c2 = fma2(g22, n2[p], g12 * n1[p]);
        const V2 K = (fma2(n2[p], c2, n1[p] * c1) * sc2i[p]) * FT(0.5) + (nu[p] * (nu[p] * g33lo[p]) + nu1[p] * (nu1[p] * g33hi[p])) * FT(0.25);
        const V2 etot = nre[p] * rcpn2(nr[p]);
        const V2 T = max2(P.T_min_sgs, fma2(((etot - K) - phi[p]) + P.RT0, V2(P.icv), V2(P.T_0)));
        hn[p] = fma2(T, V2(P.R_d), etot);
      }
    }
    const V2 hn_m1[2] = K8_M1(hn), hn_m2[2] = K8_M2(hn), hn_p1[2] = K8_P1(hn), rn_m1[2] = K8_M1(rn);
    V2 flx[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int v = 2 * lane + p;
      flx[p] = V2(FT(0));
      if (interior[p]) {
        const V2 wv = nu[p] * g33lo[p];
        const V2 mr = fma2(nr[p], V2(mc[p]), rn_m1[p] * mclo[p]) * FT(0.5);
        flx[p] = (mr * wv) * upw_minus_central2(P, wv, hn_m2[p], hn_m1[p], hn[p], hn_p1[p], v, nv);
      }
    }
    const V2 fp[2] = K8_P1(flx);
#pragma unroll
    for (int p = 0; p < 2; ++p) nre[p] = nre[p] + ((-(fp[p] - flx[p])) * rmc[p]) * dtg;
  }
  st8(nre, gN + 3 * cs, nv, cv[0], cv[1]);
}

}  // namespace b200
