// kernels_imp5d.cuh — fused implicit stage WITH implicit vertical diffusion in the packed row layout of kernels_imp5.cuh.
//
// Same algebra as k_imp_stage_diff (kernels_vdiff.cuh: cache_imp! → Wfact incl. update_diffusion_jacobian! → R = dtγ·T_imp(U) incl. the
// diffusion tendency → ldiv! = ApproximateBlockArrowheadIterativeSolve → N = U − ΔU → cache_imp! → T_post_imp!; reference:
// implicit_tendency.jl:36-98,185-339, vertical_diffusion_boundary_layer.jl:64-154, manual_sparse_jacobian.jl:713-870,1031-1261,538-578),
// restructured like k5_imp_stage: one element per CTA, thread (v, j) owns FOUR columns at level v as two f32x2 pairs, a warp is 32
// consecutive levels, vertical neighbours are 64-bit shared loads of pair slabs, and every tridiagonal system — (uₕ,uₕ) with two
// right-hand sides, A_ee = (ρe_tot,ρe_tot), the preconditioner P of the Schur complement, the passive-tracer blocks — is solved by
// parallel cyclic reduction on normalised rows (one reciprocal and one barrier per step).  k_imp_stage_diff (one point per thread,
// ≈ 160 block barriers, IEEE divisions) took ≈ 2 ms per launch at he30/ze63; it stays as the reference implementation of the tests.
#pragma once
#include "kernels_imp5.cuh"
#include "kernels_vdiff.cuh"

namespace b200 {

constexpr int IMP5D_SLABS = 26;
template <class FT> constexpr size_t smem_imp5d() { return (size_t)IMP5D_SLABS * PSLAB * sizeof(P2<FT>); }

// Parallel cyclic reduction of NR systems with the same tridiagonal matrix (l, d, u) at the thread's row v for its two pairs; rows
// beyond the system must come in as identity rows (l = u = 0, d = 1, y = 0).  Double-buffered slabs ba/bc/by (one barrier per step);
// the solution is returned in y and published to out[r] (a slab readable at v ± 1 after the final barrier).
template <class FT, int NR>
__device__ __forceinline__ void pcr5(int nrows, int v, int o0, const P2<FT> (&l)[2], const P2<FT> (&d)[2], const P2<FT> (&u)[2], P2<FT> (&y)[NR][2],
                                     P2<FT>* const (&ba)[2], P2<FT>* const (&bc)[2], P2<FT>* const (&by)[NR][2], P2<FT>* const (&out)[NR]) {
  using V2 = P2<FT>;
  V2 a[2], c[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const V2 rd = rcpn2(d[p]);
    a[p] = l[p] * rd; c[p] = u[p] * rd;
#pragma unroll
    for (int r = 0; r < NR; ++r) y[r][p] = y[r][p] * rd;
  }
  int buf = 0;
  for (int st = 1; st < nrows; st <<= 1, buf ^= 1) {
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int o = o0 + p * PLV;
      ba[buf][o] = a[p]; bc[buf][o] = c[p];
#pragma unroll
      for (int r = 0; r < NR; ++r) by[r][buf][o] = y[r][p];
    }
    __syncthreads();
    const bool hm = v >= st, hp = v + st < LV;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int o = o0 + p * PLV;
      V2 am(FT(0)), cm(FT(0)), ap(FT(0)), cp(FT(0));
      if (hm) { am = ba[buf][o - st]; cm = bc[buf][o - st]; }
      if (hp) { ap = ba[buf][o + st]; cp = bc[buf][o + st]; }
      const V2 rd = rcpn2(V2(FT(1)) - fma2(c[p], ap, a[p] * cm));
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const V2 ym = hm ? by[r][buf][o - st] : V2(FT(0)), yp = hp ? by[r][buf][o + st] : V2(FT(0));
        y[r][p] = (y[r][p] - fma2(c[p], yp, a[p] * ym)) * rd;
      }
      a[p] = -((a[p] * am) * rd);
      c[p] = -((c[p] * cp) * rd);
    }
  }
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int p = 0; p < 2; ++p) out[r][o0 + p * PLV] = y[r][p];
  __syncthreads();
}

template <class FT, int NVC>
__global__ void __launch_bounds__(256, sizeof(FT) == 4 ? 2 : 1)
k5_imp_stage_diff(Par<FT> P, VDiff<FT> D, const FT* __restrict__ hgeo, const VLev<FT>* __restrict__ vlev, const FT* __restrict__ Yc,
                  const FT* __restrict__ Yf, FT* __restrict__ Nc, FT* __restrict__ Nf, FT dtg) {
  using V2 = P2<FT>;
  pdl_launch();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  V2* sb = reinterpret_cast<V2*>(smem_raw);
  int ns = 0;
  auto slab = [&]() { return sb + (ns++) * PSLAB; };
  V2 *s_rho = slab(), *s_u3 = slab(), *s_h = slab(), *s_A = slab(), *s_M = slab(), *s_dp = slab(), *s_Pi = slab(), *s_thv = slab(),
     *s_thp = slab(), *s_phr = slab(), *s_u1 = slab(), *s_u2 = slab(), *s_sd = slab(), *s_kh = slab(), *s_pw = slab();
  V2 *s_x = slab(), *s_z = slab(), *s_r1 = slab(), *s_r2 = slab(), *s_ye = slab();
  V2* const ba[2] = {slab(), slab()};
  V2* const bc[2] = {slab(), slab()};
  V2* const by1[1][2] = {{slab(), slab()}};
  V2* const by2[2][2] = {{by1[0][0], by1[0][1]}, {s_Pi, s_thv}};  // the second right-hand side of the uₕ solve reuses two thermodynamic slabs (consumed by then)
  const int e = blockIdx.x, v = threadIdx.x & 63, j = threadIdx.x >> 6, n0 = j * 4, nv = NVC ? NVC : P.nv, nf = nv + 1;
  const bool cv = v < nv, fv = v < nf, interior = v > 0 && v < nv, lo = v > 0, hi = v < nv - 1;
  const int vm = v > 0 ? v - 1 : 0, vm2 = v > 1 ? v - 2 : 0, vp = v < LV - 1 ? v + 1 : v;
  const int o0 = (2 * j) * PLV + v;
  const FT kap = P.R_d / P.cv_d;
  const int vc = cv ? v : nv - 1, vmc = vm < nv ? vm : nv - 1, vpc = v + 1 < nv ? v + 1 : nv - 1, vf = fv ? v : nv, vf1 = v + 1 <= nv ? v + 1 : nv;
  const FT sc2i = vlev->sc2i[vc], phi = vlev->phic[vc], mc = vlev->mc[vc], mclo = vlev->mc[vmc], rmc = vlev->rmc[vc], rmclo = vlev->rmc[vmc],
           g33lo = vlev->g33f[vf], g33hi = vlev->g33f[vf1], g33m = vlev->g33f[vm], dphif = vlev->dphif[vf],
           beta = P.rayleigh ? vlev->brw[vf] : FT(0);
  // diffusion: face weight factor Δz_f g³³_f s_f² and the 1/s_c of this level and its neighbours (momentum diffusion acts on uₕ/s_c)
  const FT wfac = dtg * (vlev->dzf[vf] * vlev->g33f[vf] / vlev->sf2i[vf]);
  const FT is0 = sqrt(sc2i), isl = sqrt(vlev->sc2i[vmc]), ish = sqrt(vlev->sc2i[vpc]), sc2i_lo = vlev->sc2i[vmc];
  const FT kdec = D.mode == 2 ? D.kdec[vc] : FT(0);
  pdl_wait(Yc, Yf, Nc, Nf);
  const int cs = 16 * nv;
  const FT* gY = Yc + ((size_t)e * P.ncf * 16 * nv + n0 * nv + v);
  const FT* gYf = Yf + ((size_t)e * 16 * nf + n0 * nf + v);
  FT* gN = Nc + ((size_t)e * P.ncf * 16 * nv + n0 * nv + v);
  FT* gNf = Nf + ((size_t)e * 16 * nf + n0 * nf + v);
  V2 rho[2], u1[2], u2[2], re[2], u3[2];
  ld2g(rho, gY, nv, cv, FT(1)); ld2g(u1, gY + cs, nv, cv, FT(0)); ld2g(u2, gY + 2 * cs, nv, cv, FT(0));
  ld2g(re, gY + 3 * cs, nv, cv, FT(0)); ld2g(u3, gYf, nf, interior, FT(0));  // u₃ boundary filter on load
  const FT* hgp = hgeo + (size_t)e * HG_N * 16 + n0;
  V2 g11[2], g12[2], g22[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    g11[p] = ldpair(hgp + HG_GI11 * 16 + 2 * p); g12[p] = ldpair(hgp + HG_GI12 * 16 + 2 * p); g22[p] = ldpair(hgp + HG_GI22 * 16 + 2 * p);
    const int o = o0 + p * PLV;
    s_rho[o] = rho[p]; s_u3[o] = u3[p]; s_u1[o] = u1[p]; s_u2[o] = u2[p];
  }
  __syncthreads();  // (1) state slabs
  // ---- centre thermodynamics, face mass-flux pieces, eddy diffusivity
  V2 u3h[2], rlo[2], h[2], A[2], M[2], ir[2], Tc[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int o = o0 + p * PLV;
    u3h[p] = s_u3[o - v + vp];
    rlo[p] = s_rho[o - v + vm];
    V2 Pi(FT(1)), thv(FT(0)), thp(FT(0)), phr(FT(1)), dp(FT(0)), sd(FT(0)), kh(FT(1));
    h[p] = V2(FT(0)); Tc[p] = V2(FT(0));
    ir[p] = rcpn2(rho[p]);
    if (cv) {
      const V2 c1 = fma2(g12[p], u2[p], g11[p] * u1[p]), c2 = fma2(g22[p], u2[p], g12[p] * u1[p]);
      const V2 K = (fma2(u2[p], c2, u1[p] * c1) * sc2i) * FT(0.5) + (u3[p] * (u3[p] * g33lo) + u3h[p] * (u3h[p] * g33hi)) * FT(0.25);
      const Pt2<FT> t = thermo2(P, rho[p], re[p], K, phi);
      h[p] = t.h; Pi = t.Pi; thv = t.thv; thp = t.thp; phr = pgf_aux2(t); Tc[p] = t.T;
      dp = fma2(t.T, V2(P.R_d - kap * P.cv_d), ((V2(P.T_0 * P.cp_d) - K) - phi) * kap);
      sd = fma2(t.T - P.T_0, V2(P.cp_d), V2(phi));  // dry static energy cp_d (T − T_0) + Φ
      if (D.mode == 2) {
        kh = V2(kdec);
      } else {  // VerticalDiffusion: C_E |uₕ(level 1)| Δz₁/2 below 850 hPa, Gaussian taper in pressure above
        const V2 a = s_u1[o - v], b = s_u2[o - v];
        const V2 n2 = fma2(b, fma2(g22[p], b, g12[p] * a), a * fma2(g12[p], b, g11[p] * a)) * vlev->sc2i[0];
        const V2 KE = V2(sqrt(n2.lo()), sqrt(n2.hi())) * D.ce_za;
        FT kk[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const FT pk = k ? t.p.hi() : t.p.lo(), ke = k ? KE.hi() : KE.lo();
          const FT x = (FT(85000) - pk) / FT(10000);
          kk[k] = pk > FT(85000) ? ke : ke * exp_(-(x * x));
        }
        kh = V2(kk[0], kk[1]);
      }
    }
    A[p] = M[p] = V2(FT(0));
    if (interior) {
      const V2 mr = fma2(rho[p], V2(mc), rlo[p] * mclo) * FT(0.5);
      A[p] = (mr * dtg) * g33lo;
      M[p] = mr * (u3[p] * g33lo);
    }
    s_h[o] = h[p]; s_Pi[o] = Pi; s_thv[o] = thv; s_thp[o] = thp; s_phr[o] = phr; s_dp[o] = dp; s_A[o] = A[p]; s_M[o] = M[p];
    s_sd[o] = sd; s_kh[o] = kh;
  }
  __syncthreads();  // (2)
  // ---- dtγ (J g³³ ᶠρK)/J2 at face v
  V2 wl[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int o = o0 + p * PLV;
    wl[p] = V2(FT(0));
    if (interior) {
      const V2 rf = (rlo[p] + rho[p]) * FT(0.5);
      const V2 ik = (rcpn2(max2(D.eps, s_kh[o - 1])) + rcpn2(max2(D.eps, s_kh[o]))) * FT(0.5);
      wl[p] = (rf * rcpn2(ik)) * wfac;
    }
    s_pw[o] = wl[p];
  }
  __syncthreads();  // (3)
  // ---- residuals, centre-row and face-row coefficients (as k5_imp_stage), diffusion blocks
  V2 rr[2], rre[2], a0[2], a1[2], b0[2], b1[2], el[2], ed[2], eu[2], fc[2], pl[2], pd[2], pu[2], r12[2][2];
  V2 sl[2], sd_[2], su[2], uel[2], ueh[2], b3[2], Pl[2], Pd[2], Pu[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int o = o0 + p * PLV, om = o - v + vm, om2 = o - v + vm2, op = o - v + vp;
    const V2 hl = s_h[om], hm2 = s_h[om2], hp1 = s_h[op];
    const V2 hf0 = v > 0 ? (hl + h[p]) * FT(0.5) : V2(FT(0));
    const V2 hfp = v < nv - 1 ? (h[p] + hp1) * FT(0.5) : V2(FT(0));
    const V2 Ap = s_A[op], Mp = s_M[op];
    const V2 wh = (cv && hi) ? s_pw[op] : V2(FT(0)), wlo = (cv && lo) ? wl[p] : V2(FT(0));
    const V2 irl = rcpn2(rlo[p]), irh = rcpn2(s_rho[op]);
    // advective part: ρ_new = ρ + rr − a0·x[v] − a1·x[v+1] etc.
    rr[p] = ((Mp - M[p]) * (-dtg)) * rmc;
    a0[p] = A[p] * rmc; a1[p] = -(Ap * rmc);
    b0[p] = a0[p] * hf0; b1[p] = a1[p] * hfp;
    rre[p] = ((Mp * hfp - M[p] * hf0) * (-dtg)) * rmc;
    el[p] = eu[p] = pl[p] = pu[p] = V2(FT(0)); ed[p] = pd[p] = V2(FT(1)); fc[p] = V2(FT(0));
    r12[0][p] = r12[1][p] = V2(FT(0));
    if (cv) {
      {  // dry static energy diffusion (vertical_diffusion_boundary_layer.jl:101-102)
        const V2 s0 = s_sd[o];
        const V2 fl = lo ? wlo * (s0 - s_sd[om]) : V2(FT(0)), fh = hi ? wh * (s_sd[op] - s0) : V2(FT(0));
        rre[p] = rre[p] + (fh - fl) * rmc;
      }
      if (D.momentum) {  // uₕ strain-rate form on uₕ/s_c (:91-96)
        const V2 sir = ir[p] * (rmc / is0);
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const V2* su_ = a ? s_u2 : s_u1;
          const V2 c0 = (a ? u2[p] : u1[p]) * is0;
          const V2 fl = lo ? wlo * (c0 - su_[om] * isl) : V2(FT(0)), fh = hi ? wh * (su_[op] * ish - c0) : V2(FT(0));
          r12[a][p] = (fh - fl) * sir;
        }
      }
      const V2 l_ = lo ? wlo * rmc : V2(FT(0)), h_ = hi ? wh * rmc : V2(FT(0));
      const V2 dg = -(l_ + h_);
      const V2 m = dg * (ir[p] * D.cpcv);
      el[p] = lo ? l_ * (irl * D.cpcv) : V2(FT(0));
      ed[p] = m - FT(1);
      eu[p] = hi ? h_ * (irh * D.cpcv) : V2(FT(0));
      fc[p] = m * rcpn2(m - FT(1));
      pl[p] = l_ * ir[p]; pd[p] = dg * ir[p] - FT(1); pu[p] = h_ * ir[p];
    }
    // face row v (manual_sparse_jacobian.jl:746-868): Schur tridiagonal T of the dry blocks, couplings, R₃ = dtγ T_imp(u₃)
    sl[p] = su[p] = uel[p] = ueh[p] = V2(FT(0)); sd_[p] = V2(dtg * (-beta) - FT(1)); b3[p] = V2(FT(0));
    Pl[p] = Pu[p] = V2(FT(0)); Pd[p] = sd_[p];
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
      const V2 ue_lo = (irf * dtg) * kap, ue_hi = -ue_lo;
      const V2 x_lo = irf * (rlo[p] * (-kap)), x_hi = -(irf * (rho[p] * (-kap)));
      const V2 k0 = u3[p] * (FT(0.5) * g33lo);
      V2 l = (x_lo * (u3m * (FT(0.5) * g33m))) * dtg;
      V2 d = (fma2(x_hi, k0, x_lo * k0) - beta) * dtg - FT(1);
      V2 u = (x_hi * (u3h[p] * (FT(0.5) * g33hi))) * dtg;
      const V2 ru_lo_a = Am * rmclo, ru_hi_a = -(A[p] * rmclo), ru_lo_b = a0[p], ru_hi_b = a1[p];
      const V2 eu_lo_a = ru_lo_a * hfm, eu_hi_a = ru_hi_a * hf0;
      l = l + fma2(ue_lo, eu_lo_a, ur_lo * ru_lo_a);
      d = d + (fma2(ur_hi, ru_lo_b, ur_lo * ru_hi_a) + fma2(ue_hi, b0[p], ue_lo * eu_hi_a));
      u = u + fma2(ue_hi, b1[p], ur_hi * ru_hi_b);
      sl[p] = l; sd_[p] = d; su[p] = u; uel[p] = ue_lo; ueh[p] = ue_hi;
      const V2 rr_a = ((M[p] - Mm) * (-dtg)) * rmclo;
      const V2 tf = -((V2(dphif) - dphr) + (((thpl + thp) * FT(0.5)) * P.cp_d) * dPi) - u3[p] * beta;
      // Schur right-hand side: R₃ + A₃ρ R_ρ (A_ρρ = −I) − A₃uₕ Δuₕ is added after the uₕ solve, − A₃e A_ee⁻¹ R_ρe after the A_ee solve
      b3[p] = fma2(tf, V2(dtg), fma2(ur_lo, rr_a, ur_hi * rr[p]));
      // preconditioner: the Schur complement with A_ee replaced by its main diagonal, T − A₃e Diag(m/(m − 1)) A_e3  (m = d_ee + 1)
      const V2 wl2 = s_pw[om];  // face v − 1
      const V2 l_lo = v > 1 ? wl2 * rmclo : V2(FT(0)), h_lo = wl[p] * rmclo;
      const V2 m_lo = (-(l_lo + h_lo)) * (irl * D.cpcv);
      const V2 fc_lo = m_lo * rcpn2(m_lo - FT(1));
      const V2 ca = ue_lo * fc_lo, cb = ue_hi * fc[p];
      Pl[p] = l - ca * eu_lo_a;
      Pd[p] = d - (ca * eu_hi_a + cb * b0[p]);
      Pu[p] = u - cb * b1[p];
    }
  }
  // ---- Δuₕ: exact tridiagonal solves (two right-hand sides), or −R = 0 without momentum diffusion
  V2* const o12[2] = {s_r1, s_r2};
  if (D.momentum) {
    __syncthreads();  // by2 aliases s_Pi / s_thv: every thread has read its thermodynamic neighbours
    pcr5<FT, 2>(nv, v, o0, pl, pd, pu, r12, ba, bc, by2, o12);
  }
  // ---- y_e = A_ee⁻¹ R_ρe
  V2 ye[1][2] = {{rre[0], rre[1]}};
  V2* const oye[1] = {s_ye};
  if (!cv) { ye[0][0] = ye[0][1] = V2(FT(0)); }
  pcr5<FT, 1>(nv, v, o0, el, ed, eu, ye, ba, bc, by1, oye);
  // ---- Schur right-hand side, x₃ = P⁻¹ b
  V2 x3[1][2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int o = o0 + p * PLV;
    if (interior) {
      b3[p] = b3[p] - fma2(uel[p], s_ye[o - 1], ueh[p] * ye[0][p]);
      if (D.momentum) {  // − A₃uₕ Δuₕ: (u₃, uₕ) = dtγ ᶠp_grad ⋅ Diag(−κρ) ⋅ CT12(uₕ) at centres v − 1 and v (manual_sparse_jacobian.jl:855-868)
        const V2 irf = rcpn2((rlo[p] + rho[p]) * FT(0.5));
        const V2 xl = (irf * (rlo[p] * (-kap))) * dtg, xh = (-(irf * (rho[p] * (-kap)))) * dtg;
        const V2 u1l_ = s_u1[o - 1], u2l_ = s_u2[o - 1];
        const V2 ck1 = fma2(g12[p], u2[p], g11[p] * u1[p]) * sc2i, ck2 = fma2(g22[p], u2[p], g12[p] * u1[p]) * sc2i;
        const V2 ck1l = fma2(g12[p], u2l_, g11[p] * u1l_) * sc2i_lo, ck2l = fma2(g22[p], u2l_, g12[p] * u1l_) * sc2i_lo;
        b3[p] = b3[p] - (fma2(xl * ck1l, s_r1[o - 1], (xh * ck1) * r12[0][p]) + fma2(xl * ck2l, s_r2[o - 1], (xh * ck2) * r12[1][p]));
      }
    }
    x3[0][p] = b3[p];
  }
  V2* const ox[1] = {s_x};
  pcr5<FT, 1>(nf, v, o0, Pl, Pd, Pu, x3, ba, bc, by1, ox);
  // ---- stationary iteration x ← x + P⁻¹(b − S x),  S x = T x − A₃e A_ee⁻¹ A_e3 x  (A₃e(A_ee⁻¹ + I) form as in k_ldiv_diff)
  for (int it = 0; it < D.n_iters; ++it) {
    V2 z[1][2], yv[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int o = o0 + p * PLV;
      yv[p] = cv ? fma2(b1[p], s_x[o - v + vp], b0[p] * x3[0][p]) : V2(FT(0));
      z[0][p] = yv[p];
    }
    V2* const oz[1] = {s_z};
    pcr5<FT, 1>(nv, v, o0, el, ed, eu, z, ba, bc, by1, oz);
    // publish z + y for the face rows
#pragma unroll
    for (int p = 0; p < 2; ++p) s_ye[o0 + p * PLV] = z[0][p] + yv[p];
    __syncthreads();
    V2 r3[1][2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int o = o0 + p * PLV;
      V2 tx = sd_[p] * x3[0][p];
      if (v > 0) tx = fma2(sl[p], s_x[o - 1], tx);
      if (v < nv) tx = fma2(su[p], s_x[o - v + vp], tx);
      V2 r = b3[p] - tx;
      if (interior) r = r + fma2(uel[p], s_ye[o - 1], ueh[p] * s_ye[o]);
      r3[0][p] = fv ? r : V2(FT(0));
    }
    V2* const orr[1] = {s_z};
    pcr5<FT, 1>(nf, v, o0, Pl, Pd, Pu, r3, ba, bc, by1, orr);
#pragma unroll
    for (int p = 0; p < 2; ++p) { x3[0][p] = x3[0][p] + r3[0][p]; s_x[o0 + p * PLV] = x3[0][p]; }
    __syncthreads();
  }
  // ---- Δρ, Δρe_tot = A_ee⁻¹(R_ρe − A_e3 x₃), the Newton update
  V2 x1[2], dre[1][2], nr[2], nre[2], nu[2], nu1[2], n1[2], n2[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    x1[p] = s_x[o0 + p * PLV - v + vp];
    dre[0][p] = cv ? rre[p] - fma2(b1[p], x1[p], b0[p] * x3[0][p]) : V2(FT(0));
  }
  V2* const ode[1] = {s_ye};
  pcr5<FT, 1>(nv, v, o0, el, ed, eu, dre, ba, bc, by1, ode);
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    nr[p] = (rho[p] + rr[p]) - fma2(a1[p], x1[p], a0[p] * x3[0][p]);  // ρ − Δρ, Δρ = a0 x[v] + a1 x[v+1] − rr
    nre[p] = re[p] - dre[0][p];
    n1[p] = u1[p] - r12[0][p]; n2[p] = u2[p] - r12[1][p];
    nu[p] = interior ? u3[p] - x3[0][p] : V2(FT(0));
    nu1[p] = (v + 1 < nv) ? u3h[p] - x1[p] : V2(FT(0));
  }
  if (cv) { st2g(nr, gN, nv); st2g(n1, gN + cs, nv); st2g(n2, gN + 2 * cs, nv); }
  if (fv) st2g(nu, gNf, nf);
  // ---- passive tracers: Δ(ρχ) = (dtγ D ⋅ Diag(1/ρ) − I)⁻¹ dtγ D χ with the OLD 1/ρ
  for (int q = 4; q < P.ncf; ++q) {
    V2 rq[2], tl[2], td[2], tu[2], zq[1][2];
    ld2g(rq, gY + q * cs, nv, cv, FT(0));
#pragma unroll
    for (int p = 0; p < 2; ++p) s_z[o0 + p * PLV] = rq[p] * ir[p];  // χ
    __syncthreads();
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int o = o0 + p * PLV, om = o - v + vm, op = o - v + vp;
      tl[p] = tu[p] = zq[0][p] = V2(FT(0)); td[p] = V2(FT(1));
      if (cv) {
        const V2 wh = hi ? s_pw[op] : V2(FT(0)), wlo = lo ? wl[p] : V2(FT(0));
        const V2 l_ = wlo * rmc, h_ = wh * rmc;
        tl[p] = lo ? l_ * rcpn2(rlo[p]) : V2(FT(0)); td[p] = -(l_ + h_) * ir[p] - FT(1); tu[p] = hi ? h_ * rcpn2(s_rho[op]) : V2(FT(0));
        const V2 c0 = s_z[o];
        const V2 fl = lo ? wlo * (c0 - s_z[om]) : V2(FT(0)), fh = hi ? wh * (s_z[op] - c0) : V2(FT(0));
        zq[0][p] = (fh - fl) * rmc;
      }
    }
    V2* const oq[1] = {s_r1};
    __syncthreads();  // s_z is read above and PCR's output slab differs, but keep the phases apart
    pcr5<FT, 1>(nv, v, o0, tl, td, tu, zq, ba, bc, by1, oq);
    V2 nq[2] = {rq[0] - zq[0][0], rq[1] - zq[0][1]};
    if (cv) st2g(nq, gN + q * cs, nv);
  }
  // ---- cache_imp!(N) and T_post_imp!: (upwinded − centred) enthalpy flux of the updated state
  if (P.upwinding != 0) {
    V2 hn[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      hn[p] = V2(FT(0));
      if (cv) {
        const V2 c1 = fma2(g12[p], n2[p], g11[p] * n1[p]), c2 = fma2(g22[p], n2[p], g12[p] * n1[p]);
        const V2 K = (fma2(n2[p], c2, n1[p] * c1) * sc2i) * FT(0.5) + (nu[p] * (nu[p] * g33lo) + nu1[p] * (nu1[p] * g33hi)) * FT(0.25);
        const V2 etot = nre[p] * rcpn2(nr[p]);
        const V2 T = max2(P.T_min_sgs, fma2(((etot - K) - phi) + P.RT0, V2(P.icv), V2(P.T_0)));
        hn[p] = fma2(T, V2(P.R_d), etot);
      }
      s_h[o0 + p * PLV] = hn[p]; s_rho[o0 + p * PLV] = cv ? nr[p] : V2(FT(1));
    }
    __syncthreads();
    V2 flx[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      flx[p] = V2(FT(0));
      if (interior) {
        const int o = o0 + p * PLV, om = o - 1, om2 = o - v + vm2, op = o - v + vp;
        const V2 w = nu[p] * g33lo;
        const V2 mr = fma2(nr[p], V2(mc), s_rho[om] * mclo) * FT(0.5);
        flx[p] = (mr * w) * upw_minus_central2(P, w, s_h[om2], s_h[om], hn[p], s_h[op], v, nv);
      }
      s_M[o0 + p * PLV] = flx[p];
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const V2 fp = s_M[o0 + p * PLV - v + vp];
      nre[p] = nre[p] + ((-(fp - flx[p])) * rmc) * dtg;
    }
  }
  if (cv) st2g(nre, gN + 3 * cs, nv);
}

}  // namespace b200
