// kernels_vdiff.cuh — vertical (K-theory) diffusion and the approximate arrowhead solve that implicit diffusion switches on
// (SURVEY.md §8f n2), hook level, one element (16 columns) per CTA in the slab layout of kernels_implicit.cuh.
//
//   k_vdiff_tend2  vertical_diffusion_boundary_layer_tendency!  (src/prognostic_equations/vertical_diffusion_boundary_layer.jl:64-154,
//                  dry branch + passive tracers) — ADDS to Yₜ.c; appended to T_exp (diff_mode Explicit, remaining_tendency.jl:185-195)
//                  or to T_imp (diff_mode Implicit, implicit/implicit_tendency.jl:69-78)
//   k_vdiff_jac    update_diffusion_jacobian!  (implicit/manual_sparse_jacobian.jl:1031-1261, dry non-EDMF branch): two planes per
//                  element, dtγ·(J g³³ ᶠρK)/J2 on faces and 1/ρ at centres — every tridiagonal block of the reference is a row or
//                  column scaling of ᶜadvdivᵥ_matrix ⋅ Diag(ᶠρK) ⋅ ᶠgradᵥ_matrix, so the blocks themselves are never stored
//   k_ldiv_diff    ldiv! with ApproximateBlockArrowheadIterativeSolve(ρ, ρe_tot; alg₁ = BlockLowerTriangularSolve(ρ),
//                  alg₂ = BlockLowerTriangularSolve(uₕ), P_alg₁ = MainDiagonalPreconditioner(), n_iters)
//                  (manual_sparse_jacobian.jl:538-578; ClimaCore MatrixFields field_matrix_solver.jl [UPSTREAM-RECALL])
//
// Index of this file (every kernel is parity-green on a B200, tests/test_gpu_vertical_diffusion.py; round-2 validation record in
// profiles/r2_opt_in_validation.md):
//   k_vdiff_tend2                                diffusion tendency, quarter element per CTA, no state slabs
//   k_vdiff_jac, k_ldiv_diff                     diffusion Jacobian planes; approximate arrowhead solve (16-lane Thomas sweeps)
//   k_imp_stage_diff                             fused implicit stage with implicit diffusion (b200_implicit_stage, the fused stepper)
//   k_lim_vborrow                                lim!: vertical mass-borrowing limiter (off unless configured)
//   k_t_imp2, k_wfact2, k_t_post_imp2            the dry hook kernels (quarter element per CTA); ldiv! is k5_imp_stage<…, LDIV> (kernels_imp5.cuh)
// Removed in round 2 after the A/B record was committed: the element-slab first generation (k_vdiff_tend, k_t_imp, k_wfact, k_ldiv,
// k_t_post_imp: slower) and the PCR variants of the diffusion solve (k_vdiff_jac2 + k_ldiv_diff2: 15.2 vs 10.5 ms/step).
//
// Eddy diffusivity (src/cache/eddy_diffusivity_coefficient.jl:16-42, src/cache/precomputed_quantities.jl:652-676): K_u = K_h;
// DecayWithHeightDiffusion K = D₀ exp(−(z − z_sfc)/H) (host table per level), VerticalDiffusion K = C_E |uₕ(level 1)| Δz₁/2 below
// 850 hPa, Gaussian taper in pressure above.  Face value: harmonic mean ᶠinterp(ρ)/ᶠinterp(1/max(K, ε)).
//
// Metric factors: J_f g³³_f / J2 = s_f²/Δz_f and J_c/J2 = s_c² Δz_c (VLev::mc), the horizontal Jacobian cancels; the physical
// wind is A⁻ᵀuₕ with A = s_c·A₂D, so on the flat deep shell the momentum diffusion acts on uₕ/s_c and A₂D cancels as well.
#pragma once
#include "kernels_implicit.cuh"

namespace b200 {

enum { JD_PW = 0, JD_IRHO, JD_N };  // planes written by k_vdiff_jac, each [16][Nv+1] per element

template <class FT>
struct VDiff {
  int mode;      // 1 VerticalDiffusion, 2 DecayWithHeightDiffusion
  int momentum;  // !disable_momentum_vertical_diffusion
  int n_iters;   // approximate_linear_solve_iters
  FT ce_za;      // C_E · Δz(level 1)/2
  FT eps;        // eps(FT)
  FT cpcv;       // cp_d / cv_d  (∂s_d/∂e_tot, manual_sparse_jacobian.jl:1129-1131 with cv_m = cv_d)
  const FT* kdec;  // [LV] D₀ exp(−(z_c − z_sfc)/H)
};

// K_h at the centres of the element → kh; needs S.rho, S.u1, S.u2, S.T.
template <class FT>
__device__ __forceinline__ void vdiff_kh(const Par<FT>& P, const VDiff<FT>& D, const FT* hg, const VLev<FT>& V,
                                         const ImpSlabs<FT>& S, FT* kh) {
  const int nv = P.nv;
  for (int idx = threadIdx.x; idx < NN * LV; idx += NT) {
    int n = idx >> 6, v = idx & 63;
    if (v >= nv) continue;
    int o = n * LVP + v;
    FT K;
    if (D.mode == 2) {
      K = D.kdec[v];
    } else {
      FT a = S.u1[n * LVP], b = S.u2[n * LVP];  // Fields.level(ᶜuₕ, 1)
      FT g11 = hg[HG_GI11 * 16 + n], g12 = hg[HG_GI12 * 16 + n], g22 = hg[HG_GI22 * 16 + n];
      FT nrm = sqrt((a * (g11 * a + g12 * b) + b * (g12 * a + g22 * b)) * V.sc2i[0]);
      FT KE = D.ce_za * nrm;
      FT p = S.rho[o] * P.R_d * S.T[o];
      FT x = (FT(85000) - p) / FT(10000);
      K = p > FT(85000) ? KE : KE * exp_(-(x * x));
    }
    kh[o] = K;
  }
}

// scale · (J g³³ ᶠρK)/J2 on the faces of the element → pw (zero on the boundary faces: ᶜdiffdivᵥ / ᶠgradᵥ boundary conditions)
template <class FT>
__device__ __forceinline__ void vdiff_pw(const Par<FT>& P, const VDiff<FT>& D, const VLev<FT>& V, const ImpSlabs<FT>& S,
                                         const FT* kh, FT* pw, FT scale) {
  const int nv = P.nv, nf = nv + 1;
  for (int idx = threadIdx.x; idx < NN * LV; idx += NT) {
    int n = idx >> 6, f = idx & 63;
    if (f >= nf) continue;
    int o = n * LVP + f;
    FT w = FT(0);
    if (f > 0 && f < nv) {
      FT rf = FT(0.5) * (S.rho[o - 1] + S.rho[o]);
      FT ik = FT(0.5) * (FT(1) / fmax_(kh[o - 1], D.eps) + FT(1) / fmax_(kh[o], D.eps));
      w = scale * (V.dzf[f] * V.g33f[f] / V.sf2i[f]) * (rf / ik);
    }
    pw[o] = w;
  }
}


// Second generation of k_vdiff_tend: a quarter element (4 columns × 64 levels) per CTA, no state slabs — each thread loads its own
// point, computes T (and K_h) once, and only the six per-column profiles the vertical differences need (ρ, K_h, s_d, uₕ/s_c, χ) and the
// face weights go through 7.4 KB of shared memory, so the kernel is HBM-bound (read Y, read-modify-write Yₜ) instead of latency-bound.
// Same operations in the same order as k_vdiff_tend: bitwise identical in the CPU emulator (tests/test_kernels_cpu_emulation.py);
// on the GPU the two differ by FMA contraction only (both ≤ 1.2e-13 from the Float64 oracle).  Measured on B200, he30/ze63 Float32:
// 77 µs per launch against 267 µs for k_vdiff_tend (profiles/r1_vdiff_timing.jsonl).
constexpr int VD2_ST = 66;  // column stride
constexpr int VD2_ARR = 7;
template <class FT>
__global__ void __launch_bounds__(NT) k_vdiff_tend2(Par<FT> P, VDiff<FT> D, const FT* __restrict__ hgeo,
                                                    const VLev<FT>* __restrict__ vlev, const FT* __restrict__ Yc,
                                                    const FT* __restrict__ Yf, FT* Ytc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<FT> sm(smem_raw);
  FT* s_rho = sm.take(4 * VD2_ST); FT* s_kh = sm.take(4 * VD2_ST); FT* s_pw = sm.take(4 * VD2_ST); FT* s_sd = sm.take(4 * VD2_ST);
  FT* s_a = sm.take(4 * VD2_ST); FT* s_b = sm.take(4 * VD2_ST); FT* s_c = sm.take(4 * VD2_ST);
  const int h = blockIdx.x >> 2, nl = threadIdx.x >> 6, n = (blockIdx.x & 3) * 4 + nl, v = threadIdx.x & 63, nv = P.nv;
  const VLev<FT>& V = *vlev;
  const FT* hg = hgeo + (size_t)h * HG_N * 16;
  const FT* gY = Yc + (size_t)h * P.ncf * 16 * nv;
  FT* gT = Ytc + (size_t)h * P.ncf * 16 * nv;
  const bool act = v < nv;
  const int o = nl * VD2_ST + v;
  FT rho = FT(1), is0 = FT(0);
  if (act) {
    rho = gY[(0 * 16 + n) * nv + v];
    const FT u1 = gY[(1 * 16 + n) * nv + v], u2 = gY[(2 * 16 + n) * nv + v], re = gY[(3 * 16 + n) * nv + v];
    const FT* gf = Yf + (size_t)h * 16 * (nv + 1) + (size_t)n * (nv + 1);
    const FT K = kinetic(hg, V, u1, u2, gf[v], gf[v + 1], n, v);
    const Pt<FT> t = thermo(P, rho, re, K, V.phic[v]);
    FT kh;
    if (D.mode == 2) {
      kh = D.kdec[v];
    } else {
      const FT a = gY[(1 * 16 + n) * nv], b = gY[(2 * 16 + n) * nv];  // Fields.level(ᶜuₕ, 1)
      const FT g11 = hg[HG_GI11 * 16 + n], g12 = hg[HG_GI12 * 16 + n], g22 = hg[HG_GI22 * 16 + n];
      const FT nrm = sqrt((a * (g11 * a + g12 * b) + b * (g12 * a + g22 * b)) * V.sc2i[0]);
      const FT KE = D.ce_za * nrm;
      const FT p = rho * P.R_d * t.T;
      const FT x = (FT(85000) - p) / FT(10000);
      kh = p > FT(85000) ? KE : KE * exp_(-(x * x));
    }
    is0 = sqrt(V.sc2i[v]);
    s_rho[o] = rho; s_kh[o] = kh; s_sd[o] = P.cp_d * (t.T - P.T_0) + V.phic[v];
    s_a[o] = u1 * is0; s_b[o] = u2 * is0;
  }
  __syncthreads();
  {  // weight of the lower face of level v; faces 0 and nv carry no flux
    FT w = FT(0);
    if (act && v > 0) {
      const FT rf = FT(0.5) * (s_rho[o - 1] + s_rho[o]);
      const FT ik = FT(0.5) * (FT(1) / fmax_(s_kh[o - 1], D.eps) + FT(1) / fmax_(s_kh[o], D.eps));
      w = FT(1) * (V.dzf[v] * V.g33f[v] / V.sf2i[v]) * (rf / ik);
    }
    if (v <= nv) s_pw[o] = w;
  }
  __syncthreads();
  const bool lo = v > 0, hi = v < nv - 1;
  const FT wl = (act && lo) ? s_pw[o] : FT(0), wh = (act && hi) ? s_pw[o + 1] : FT(0);
  const FT rm = act ? V.rmc[v] : FT(0);
  if (act) {
    {
      const FT s0 = s_sd[o];
      const FT fl = lo ? wl * (s0 - s_sd[o - 1]) : FT(0), fh = hi ? wh * (s_sd[o + 1] - s0) : FT(0);
      gT[(3 * 16 + n) * nv + v] += (fh - fl) * rm;
    }
    if (D.momentum) {
      const FT sir = rm / (is0 * rho);
      {
        const FT c0 = s_a[o];
        const FT fl = lo ? wl * (c0 - s_a[o - 1]) : FT(0), fh = hi ? wh * (s_a[o + 1] - c0) : FT(0);
        gT[(1 * 16 + n) * nv + v] += (fh - fl) * sir;
      }
      {
        const FT c0 = s_b[o];
        const FT fl = lo ? wl * (c0 - s_b[o - 1]) : FT(0), fh = hi ? wh * (s_b[o + 1] - c0) : FT(0);
        gT[(2 * 16 + n) * nv + v] += (fh - fl) * sir;
      }
    }
  }
  for (int q = 4; q < P.ncf; ++q) {
    __syncthreads();
    if (act) s_c[o] = gY[(size_t)(q * 16 + n) * nv + v] / rho;
    __syncthreads();
    if (act) {
      const FT c0 = s_c[o];
      const FT fl = lo ? wl * (c0 - s_c[o - 1]) : FT(0), fh = hi ? wh * (s_c[o + 1] - c0) : FT(0);
      gT[(size_t)(q * 16 + n) * nv + v] += (fh - fl) * rm;
    }
  }
}

template <class FT>
__global__ void __launch_bounds__(NT) k_vdiff_jac(Par<FT> P, VDiff<FT> D, const FT* __restrict__ hgeo,
                                                  const VLev<FT>* __restrict__ vlev, const FT* __restrict__ Yc,
                                                  const FT* __restrict__ Yf, FT dtg, FT* jacd) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<FT> sm(smem_raw);
  VLev<FT>& V = *reinterpret_cast<VLev<FT>*>(sm.take(sizeof(VLev<FT>) / sizeof(FT)));
  FT* hg = sm.take(HG_ELEM * 16);
  ImpSlabs<FT> S; imp_carve(sm, S);
  FT* kh = sm.take(SLAB); FT* pw = sm.take(SLAB);
  const int h = blockIdx.x, nv = P.nv, nf = nv + 1;
  load_vlev(&V, vlev); load_hgeo(hg, hgeo, h); imp_load_state(S, Yc, Yf, h, nv, P.ncf);
  __syncthreads();
  imp_thermo(P, hg, V, S);
  __syncthreads();
  vdiff_kh(P, D, hg, V, S, kh);
  __syncthreads();
  vdiff_pw(P, D, V, S, kh, pw, dtg);
  __syncthreads();
  FT* gj = jacd + (size_t)h * JD_N * 16 * nf;
  const size_t pl = (size_t)16 * nf;
  for (int idx = threadIdx.x; idx < NN * LV; idx += NT) {
    int n = idx >> 6, v = idx & 63;
    if (v >= nf) continue;
    size_t o = (size_t)n * nf + v;
    gj[JD_PW * pl + o] = pw[n * LVP + v];
    gj[JD_IRHO * pl + o] = v < nv ? FT(1) / S.rho[n * LVP + v] : FT(0);
  }
}


// Thomas algorithm that leaves the matrix untouched: c' goes to cpw, the solution overwrites x (same operation order as the
// oracle's _tri_solve).
template <class FT>
__device__ __forceinline__ void thomas_nd(const FT* l, const FT* d, const FT* u, FT* cpw, FT* x, int n) {
  FT cp = u[0] / d[0], dp = x[0] / d[0];
  cpw[0] = cp; x[0] = dp;
  for (int i = 1; i < n; ++i) {
    FT li = l[i];
    FT den = d[i] - li * cp;
    cp = u[i] / den;
    dp = (x[i] - li * dp) / den;
    cpw[i] = cp; x[i] = dp;
  }
  FT xx = dp;
  for (int i = n - 2; i >= 0; --i) {
    xx = x[i] - cpw[i] * xx;
    x[i] = xx;
  }
}

#define VD_FOR_POINTS(n, v) \
  for (int idx_ = threadIdx.x, n = idx_ >> 6, v = idx_ & 63; idx_ < NN * LV; idx_ += NT, n = idx_ >> 6, v = idx_ & 63)

// The Schur complement stored by k_wfact (JC_L/D/U) is T = A₃₃ + A₃ρA_ρ3 + A₃eA_e3, i.e. the exact one for A_ρρ = A_ee = −I.
// With implicit diffusion A_ee is tridiagonal (dtγ·D·Diag(cp_d/(cv_d ρ)) − I), so
//     S x = T x − A₃e (A_ee⁻¹ + I) A_e3 x           (the u₃ Schur complement of the full Jacobian)
//     P   = T − A₃e Diag(1 + 1/d_ee) A_e3           (A_ee replaced by its main diagonal: tridiagonal preconditioner)
// and 1 + 1/d_ee = m/(m − 1) with m = dtγ·D_kk·cp_d/(cv_d ρ_k) ≤ 0 evaluated without cancellation.
template <class FT>
__global__ void __launch_bounds__(NT) k_ldiv_diff(Par<FT> P, VDiff<FT> D, const VLev<FT>* __restrict__ vlev,
                                                  const FT* __restrict__ jac, const FT* __restrict__ jacd,
                                                  const FT* __restrict__ Rc, const FT* __restrict__ Rf, FT* dYc, FT* dYf) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<FT> sm(smem_raw);
  FT* sl = sm.take(SLAB); FT* sd = sm.take(SLAB); FT* su = sm.take(SLAB);   // T (faces)
  FT* pl_ = sm.take(SLAB); FT* pd = sm.take(SLAB); FT* pu = sm.take(SLAB);  // (uₕ,uₕ) / tracer blocks, then P (faces)
  FT* el = sm.take(SLAB); FT* ed = sm.take(SLAB); FT* eu = sm.take(SLAB);   // A_ee (centres)
  FT* fac = sm.take(SLAB);                                                  // m/(m − 1) (centres)
  FT* scp = sm.take(SLAB); FT* zz = sm.take(SLAB);                          // Thomas scratch / centre scratch
  FT* rr = sm.take(SLAB); FT* r1 = sm.take(SLAB); FT* r2 = sm.take(SLAB); FT* re = sm.take(SLAB);
  FT* ye = sm.take(SLAB); FT* b3 = sm.take(SLAB); FT* x3 = sm.take(SLAB); FT* r3 = sm.take(SLAB);
  FT* pw = sm.take(SLAB); FT* ir = sm.take(SLAB);
  FT* s_rmc = sm.take(LV);
  const int h = blockIdx.x, nv = P.nv, nf = nv + 1;
  const FT* gj = jac + (size_t)h * JC_N * 16 * nf;
  const FT* gd = jacd + (size_t)h * JD_N * 16 * nf;
  const size_t pl = (size_t)16 * nf;
  const FT* gRc = Rc + (size_t)h * P.ncf * 16 * nv;
  const FT* gRf = Rf + (size_t)h * 16 * nf;
  FT* gdc = dYc + (size_t)h * P.ncf * 16 * nv;
  FT* gdf = dYf + (size_t)h * 16 * nf;
  if (threadIdx.x < LV) s_rmc[threadIdx.x] = vlev->rmc[threadIdx.x];
  load_slab(sl, gj + JC_L * pl, nf); load_slab(sd, gj + JC_D * pl, nf); load_slab(su, gj + JC_U * pl, nf);
  load_slab(pw, gd + JD_PW * pl, nf); load_slab(ir, gd + JD_IRHO * pl, nf);
  load_slab(rr, gRc, nv); load_slab(r1, gRc + 16 * nv, nv); load_slab(r2, gRc + 32 * nv, nv); load_slab(re, gRc + 48 * nv, nv);
  __syncthreads();
  // ---- A_ee, the factor 1 + 1/d_ee, and the (uₕ,uₕ) block dtγ Diag(1/ρ)⋅D − I (:1252-1258)
  VD_FOR_POINTS(n, v) {
    if (v >= nv) continue;
    const int o = n * LVP + v;
    const FT lo = v > 0 ? pw[o] * s_rmc[v] : FT(0), hi = v < nv - 1 ? pw[o + 1] * s_rmc[v] : FT(0);
    const FT dg = -(lo + hi);
    const FT m = dg * (D.cpcv * ir[o]);
    el[o] = v > 0 ? lo * (D.cpcv * ir[o - 1]) : FT(0);
    ed[o] = m - FT(1);
    eu[o] = v < nv - 1 ? hi * (D.cpcv * ir[o + 1]) : FT(0);
    fac[o] = m / (m - FT(1));
    pl_[o] = lo * ir[o]; pd[o] = dg * ir[o] - FT(1); pu[o] = hi * ir[o];
    ye[o] = re[o];
  }
  __syncthreads();
  // ---- uₕ (exact tridiagonal solves or the −I fallback) and y_e = A_ee⁻¹ R_ρe
  if (D.momentum) {
    if (threadIdx.x < 16) thomas_nd(pl_ + threadIdx.x * LVP, pd + threadIdx.x * LVP, pu + threadIdx.x * LVP, scp + threadIdx.x * LVP, r1 + threadIdx.x * LVP, nv);
    else if (threadIdx.x >= 32 && threadIdx.x < 48) { int n = threadIdx.x - 32; thomas_nd(pl_ + n * LVP, pd + n * LVP, pu + n * LVP, zz + n * LVP, r2 + n * LVP, nv); }
    else if (threadIdx.x >= 64 && threadIdx.x < 80) { int n = threadIdx.x - 64; thomas_nd(el + n * LVP, ed + n * LVP, eu + n * LVP, b3 + n * LVP, ye + n * LVP, nv); }
  } else {
    VD_FOR_POINTS(n, v) { if (v < nv) { int o = n * LVP + v; r1[o] = -r1[o]; r2[o] = -r2[o]; } }
    if (threadIdx.x >= 64 && threadIdx.x < 80) { int n = threadIdx.x - 64; thomas_nd(el + n * LVP, ed + n * LVP, eu + n * LVP, b3 + n * LVP, ye + n * LVP, nv); }
  }
  __syncthreads();
  VD_FOR_POINTS(n, v) {
    if (v < nv) { gdc[(1 * 16 + n) * nv + v] = r1[n * LVP + v]; gdc[(2 * 16 + n) * nv + v] = r2[n * LVP + v]; }
  }
  // ---- passive tracers: (ρχ,ρχ) = dtγ D⋅Diag(1/ρ) − I (:1190-1195), exact tridiagonal solve
  for (int q = 4; q < P.ncf; ++q) {
    __syncthreads();
    VD_FOR_POINTS(n, v) {
      if (v >= nv) continue;
      const int o = n * LVP + v;
      const FT lo = v > 0 ? pw[o] * s_rmc[v] : FT(0), hi = v < nv - 1 ? pw[o + 1] * s_rmc[v] : FT(0);
      pl_[o] = v > 0 ? lo * ir[o - 1] : FT(0); pd[o] = -(lo + hi) * ir[o] - FT(1); pu[o] = v < nv - 1 ? hi * ir[o + 1] : FT(0);
      zz[o] = gRc[(size_t)(q * 16 + n) * nv + v];
    }
    __syncthreads();
    if (threadIdx.x < 16) thomas_nd(pl_ + threadIdx.x * LVP, pd + threadIdx.x * LVP, pu + threadIdx.x * LVP, scp + threadIdx.x * LVP, zz + threadIdx.x * LVP, nv);
    __syncthreads();
    VD_FOR_POINTS(n, v) { if (v < nv) gdc[(size_t)(q * 16 + n) * nv + v] = zz[n * LVP + v]; }
  }
  __syncthreads();
  // ---- Schur right-hand side b₃ = R₃ + A₃ρR_ρ − A₃e y_e − A₃ₕ xₕ  and the preconditioner P
  VD_FOR_POINTS(n, f) {
    if (f >= nf) continue;
    const size_t o = (size_t)n * nf + f;
    const int s = n * LVP + f;
    FT rhs = gRf[o];
    FT l = sl[s], d = sd[s], u = su[s];
    if (f > 0 && f < nv) {
      const FT ue_lo = gj[JC_UE_LO * pl + o], ue_hi = gj[JC_UE_HI * pl + o];
      rhs += gj[JC_UR_LO * pl + o] * rr[s - 1] + gj[JC_UR_HI * pl + o] * rr[s];
      rhs -= ue_lo * ye[s - 1] + ue_hi * ye[s];
      rhs -= gj[JC_U1_LO * pl + o] * r1[s - 1] + gj[JC_U1_HI * pl + o] * r1[s];
      rhs -= gj[JC_U2_LO * pl + o] * r2[s - 1] + gj[JC_U2_HI * pl + o] * r2[s];
      const FT a = ue_lo * fac[s - 1], b = ue_hi * fac[s];
      l -= a * gj[JC_EU_LO * pl + o - 1];
      d -= a * gj[JC_EU_HI * pl + o - 1] + b * gj[JC_EU_LO * pl + o];
      u -= b * gj[JC_EU_HI * pl + o];
    }
    b3[s] = rhs; x3[s] = rhs;
    pl_[s] = l; pd[s] = d; pu[s] = u;
  }
  __syncthreads();
  if (threadIdx.x < 16) thomas_nd(pl_ + threadIdx.x * LVP, pd + threadIdx.x * LVP, pu + threadIdx.x * LVP, scp + threadIdx.x * LVP, x3 + threadIdx.x * LVP, nf);  // x[0] = P⁻¹ b
  __syncthreads();
  for (int it = 0; it < D.n_iters; ++it) {
    VD_FOR_POINTS(n, v) {  // y = A_e3 x
      if (v >= nv) continue;
      const size_t o = (size_t)n * nf + v;
      const int s = n * LVP + v;
      FT y = gj[JC_EU_LO * pl + o] * x3[s] + gj[JC_EU_HI * pl + o] * x3[s + 1];
      ye[s] = y; zz[s] = y;
    }
    __syncthreads();
    if (threadIdx.x < 16) thomas_nd(el + threadIdx.x * LVP, ed + threadIdx.x * LVP, eu + threadIdx.x * LVP, scp + threadIdx.x * LVP, zz + threadIdx.x * LVP, nv);
    __syncthreads();
    VD_FOR_POINTS(n, f) {  // r = b − S x
      if (f >= nf) continue;
      const size_t o = (size_t)n * nf + f;
      const int s = n * LVP + f;
      FT tx = sd[s] * x3[s];
      if (f > 0) tx += sl[s] * x3[s - 1];
      if (f < nv) tx += su[s] * x3[s + 1];
      FT r = b3[s] - tx;
      if (f > 0 && f < nv) r += gj[JC_UE_LO * pl + o] * (zz[s - 1] + ye[s - 1]) + gj[JC_UE_HI * pl + o] * (zz[s] + ye[s]);
      r3[s] = r;
    }
    __syncthreads();
    if (threadIdx.x < 16) thomas_nd(pl_ + threadIdx.x * LVP, pd + threadIdx.x * LVP, pu + threadIdx.x * LVP, scp + threadIdx.x * LVP, r3 + threadIdx.x * LVP, nf);
    __syncthreads();
    VD_FOR_POINTS(n, f) { if (f < nf) x3[n * LVP + f] += r3[n * LVP + f]; }
    __syncthreads();
  }
  // ---- x₁ = A₁₁⁻¹(b₁ − A₁₂x₂)
  VD_FOR_POINTS(n, v) {
    const int s = n * LVP + v;
    const size_t o = (size_t)n * nf + v;
    if (v < nf) gdf[o] = x3[s];
    if (v < nv) {
      const FT x0 = x3[s], x1 = x3[s + 1];
      gdc[(0 * 16 + n) * nv + v] = gj[JC_RU_LO * pl + o] * x0 + gj[JC_RU_HI * pl + o] * x1 - rr[s];
      ye[s] = re[s] - (gj[JC_EU_LO * pl + o] * x0 + gj[JC_EU_HI * pl + o] * x1);
    }
  }
  __syncthreads();
  if (threadIdx.x < 16) thomas_nd(el + threadIdx.x * LVP, ed + threadIdx.x * LVP, eu + threadIdx.x * LVP, scp + threadIdx.x * LVP, ye + threadIdx.x * LVP, nv);
  __syncthreads();
  VD_FOR_POINTS(n, v) { if (v < nv) gdc[(3 * 16 + n) * nv + v] = ye[n * LVP + v]; }
}



// ---------------------------------------------------------------------------------------------
// lim!, second branch (src/prognostic_equations/limited_tendencies.jl:95-121): ClimaCore Limiters.VerticalMassBorrowingLimiter((0,))
// applied to χ = ρχ/ρ of every tracer, then ρχ = χ·ρ [UPSTREAM-RECALL ClimaCore 0.15.1 src/Limiters/vertical_mass_borrowing_limiter.jl,
// after E3SM's massborrow]: per column, sweep level 1 → Nv carrying the mass deficit (weights ρ·Δz), then Nv → 1 while a deficit
// remains.  One (element, tracer) per CTA: the two slabs are staged with coalesced loads, 16 lanes do the two sweeps.
template <class FT>
__global__ void __launch_bounds__(NT) k_lim_vborrow(const VLev<FT>* __restrict__ vlev, FT* Yc, int ncf, int nv, FT qmin) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<FT> sm(smem_raw);
  FT* sq = sm.take(SLAB); FT* sr = sm.take(SLAB);
  const int h = blockIdx.x, t = blockIdx.y;
  FT* gY = Yc + (size_t)h * ncf * 16 * nv;
  FT* gq = gY + (size_t)(4 + t) * 16 * nv;
  load_slab(sr, gY, nv); load_slab(sq, gq, nv);
  __syncthreads();
  if (threadIdx.x < 16) {
    FT* q = sq + threadIdx.x * LVP;
    const FT* r = sr + threadIdx.x * LVP;
    FT bmass = FT(0);
    for (int v = 0; v < nv; ++v) {
      const FT m = r[v] * vlev->dzc[v];
      const FT nmass = q[v] / r[v] + bmass / m;
      if (nmass > qmin) { q[v] = nmass; bmass = FT(0); }
      else { bmass = (nmass - qmin) * m; q[v] = qmin; }
    }
    for (int v = nv - 1; v >= 0; --v) {
      if (bmass < FT(0)) {
        const FT m = r[v] * vlev->dzc[v];
        const FT nmass = q[v] + bmass / m;
        if (nmass > qmin) { q[v] = nmass; bmass = FT(0); }
        else { bmass = (nmass - qmin) * m; q[v] = qmin; }
      }
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < NN * LV; idx += NT) {
    int n = idx >> 6, v = idx & 63;
    if (v < nv) gq[n * nv + v] = sq[n * LVP + v] * sr[n * LVP + v];
  }
}

// ---------------------------------------------------------------------------------------------
// Second generation of the dry hook kernels (B200_HOOK_KERNELS=2; matched against the first generation bit for bit in the CPU CTA
// emulator, not yet run on a B200): a QUARTER element (4 columns × 64 levels) per CTA, one point per thread, and only the column
// profiles that the vertical stencils need in shared memory (12 arrays of 4·LVP words = 12.5 KB in Float32, so 8 CTAs/SM instead of
// the 2 that the 12 whole-element slabs of the first generation allow).  The per-point device functions of kernels_implicit.cuh
// (timp_center, timp_face, face_coef, center_coef, upwind_minus_central) are reused unchanged: they address a profile as
// p[n·LVP + v] with the element node n, so the ImpSlabs pointers are shifted by −4·quarter·LVP.
template <class FT, bool MOIST = false>
__device__ __forceinline__ void q_prepare(const Par<FT>& P, const VLev<FT>& V, const FT* hg, const FT* __restrict__ Yc,
                                          const FT* __restrict__ Yf, int h, int quarter, FT* base, ImpSlabs<FT>& S) {
  const int nl = threadIdx.x >> 6, n = quarter * 4 + nl, v = threadIdx.x & 63, nv = P.nv, nf = nv + 1;
  FT** f[12] = {&S.rho, &S.u1, &S.u2, &S.re, &S.u3, &S.K, &S.h, &S.Pi, &S.thv, &S.thp, &S.phr, &S.T};
  for (int k = 0; k < 12; ++k) *f[k] = base + k * 4 * LVP - quarter * 4 * LVP;
  const FT* gY = Yc + (size_t)h * P.ncf * 16 * nv;
  const int o = n * LVP + v;
  FT* qprof = base + 13 * 4 * LVP - quarter * 4 * LVP;  // moist (0M) contexts: q_tot = ρq_tot/ρ (profile 13; 12 and 14 are scratch)
  if (v < nv) {
    S.rho[o] = gY[(0 * 16 + n) * nv + v]; S.u1[o] = gY[(1 * 16 + n) * nv + v];
    S.u2[o] = gY[(2 * 16 + n) * nv + v]; S.re[o] = gY[(3 * 16 + n) * nv + v];
  }
  if (v < nf) S.u3[o] = Yf[(size_t)h * 16 * nf + (size_t)n * nf + v];
  __syncthreads();
  if (v < nv) {
    FT K = kinetic(hg, V, S.u1[o], S.u2[o], S.u3[o], S.u3[o + 1], n, v);
    Pt<FT> t;
    if constexpr (MOIST) {
      Mst<FT> m;
      const FT rq = gY[(4 * 16 + n) * nv + v];
      t = thermo_m(P, S.rho[o], S.re[o], rq, K, V.phic[v], m);
      qprof[o] = rq / S.rho[o];
    } else {
      t = thermo(P, S.rho[o], S.re[o], K, V.phic[v]);
    }
    S.K[o] = K; S.h[o] = t.h; S.Pi[o] = t.Pi; S.thv[o] = t.thv; S.thp[o] = t.thp; S.phr[o] = pgf_aux(t); S.T[o] = t.T;
  }
  __syncthreads();
}
constexpr int Q_WORDS = 15 * 4 * LVP;  // 12 profiles + scratch + (moist) q_tot + scratch
constexpr int Q_WORDS_DRY = 13 * 4 * LVP;  // dry contexts: 12 profiles + scratch

template <class FT, bool MOIST = false>
__global__ void __launch_bounds__(NT) k_t_imp2(Par<FT> P, const FT* __restrict__ hgeo, const VLev<FT>* __restrict__ vlev,
                                               const FT* __restrict__ Yc, const FT* __restrict__ Yf, FT* Ytc, FT* Ytf) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int h = blockIdx.x >> 2, quarter = blockIdx.x & 3, n = quarter * 4 + (threadIdx.x >> 6), v = threadIdx.x & 63;
  const int nv = P.nv, nf = nv + 1;
  const VLev<FT>& V = *vlev;
  const FT* hg = hgeo + (size_t)h * HG_N * 16;
  ImpSlabs<FT> S;
  q_prepare<FT, MOIST>(P, V, hg, Yc, Yf, h, quarter, reinterpret_cast<FT*>(smem_raw), S);
  FT* gT = Ytc + (size_t)h * P.ncf * 16 * nv;
  FT* gF = Ytf + (size_t)h * 16 * nf;
  if (v < nv) {
    FT rt, et; timp_center(V, S, n, v, nv, rt, et);
    gT[(0 * 16 + n) * nv + v] = rt; gT[(1 * 16 + n) * nv + v] = FT(0);
    gT[(2 * 16 + n) * nv + v] = FT(0); gT[(3 * 16 + n) * nv + v] = et;
    for (int q = 4; q < P.ncf; ++q) gT[(q * 16 + n) * nv + v] = FT(0);
    if constexpr (MOIST) {  // central transport of the active tracer ρq_tot (implicit_tendency.jl:210-214): timp_center with q_tot for h_tot
      ImpSlabs<FT> Sq = S;
      Sq.h = reinterpret_cast<FT*>(smem_raw) + 13 * 4 * LVP - quarter * 4 * LVP;
      FT rt2, qt; timp_center(V, Sq, n, v, nv, rt2, qt);
      gT[(4 * 16 + n) * nv + v] = qt;
    }
  }
  if (v < nf) gF[n * nf + v] = timp_face(P, V, S, n, v, nv);
}

template <class FT>
__global__ void __launch_bounds__(NT) k_wfact2(Par<FT> P, const FT* __restrict__ hgeo, const VLev<FT>* __restrict__ vlev,
                                               const FT* __restrict__ Yc, const FT* __restrict__ Yf, FT dtg, FT* jac) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int h = blockIdx.x >> 2, quarter = blockIdx.x & 3, n = quarter * 4 + (threadIdx.x >> 6), v = threadIdx.x & 63;
  const int nv = P.nv, nf = nv + 1;
  const VLev<FT>& V = *vlev;
  const FT* hg = hgeo + (size_t)h * HG_N * 16;
  ImpSlabs<FT> S;
  q_prepare(P, V, hg, Yc, Yf, h, quarter, reinterpret_cast<FT*>(smem_raw), S);
  if (v >= nf) return;
  FT* gj = jac + (size_t)h * JC_N * 16 * nf;
  FaceCoef<FT> c = face_coef(P, hg, V, S, dtg, n, v, nv);
  const size_t o = (size_t)n * nf + v, pl = (size_t)16 * nf;
  gj[JC_L * pl + o] = c.l; gj[JC_D * pl + o] = c.d; gj[JC_U * pl + o] = c.u;
  gj[JC_UR_LO * pl + o] = c.ur_lo; gj[JC_UR_HI * pl + o] = c.ur_hi;
  gj[JC_UE_LO * pl + o] = c.ue_lo; gj[JC_UE_HI * pl + o] = c.ue_hi;
  gj[JC_U1_LO * pl + o] = c.u1_lo; gj[JC_U1_HI * pl + o] = c.u1_hi;
  gj[JC_U2_LO * pl + o] = c.u2_lo; gj[JC_U2_HI * pl + o] = c.u2_hi;
  FT a = FT(0), b = FT(0), cc = FT(0), dd = FT(0);
  if (v < nv) center_coef(V, S, dtg, n, v, nv, a, b, cc, dd);
  gj[JC_RU_LO * pl + o] = a; gj[JC_RU_HI * pl + o] = b; gj[JC_EU_LO * pl + o] = cc; gj[JC_EU_HI * pl + o] = dd;
}

template <class FT, bool MOIST = false>
__global__ void __launch_bounds__(NT) k_t_post_imp2(Par<FT> P, const FT* __restrict__ hgeo, const VLev<FT>* __restrict__ vlev,
                                                    const FT* __restrict__ Yc, const FT* __restrict__ Yf, FT* Ytc, FT* Ytf) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int h = blockIdx.x >> 2, quarter = blockIdx.x & 3, n = quarter * 4 + (threadIdx.x >> 6), v = threadIdx.x & 63;
  const int nv = P.nv, nf = nv + 1;
  const VLev<FT>& V = *vlev;
  const FT* hg = hgeo + (size_t)h * HG_N * 16;
  ImpSlabs<FT> S;
  FT* base = reinterpret_cast<FT*>(smem_raw);
  q_prepare<FT, MOIST>(P, V, hg, Yc, Yf, h, quarter, base, S);
  FT* flx = base + 12 * 4 * LVP - quarter * 4 * LVP;
  FT* qprof = base + 13 * 4 * LVP - quarter * 4 * LVP;
  FT* flq = base + 14 * 4 * LVP - quarter * 4 * LVP;
  const int o = n * LVP + v;
  if (v < nf) {
    FT r = FT(0), rq = FT(0);
    if (v > 0 && v < nv) {
      FT w = V.g33f[v] * S.u3[o];
      r = rho_mface(V, S.rho, o, v) * w * upwind_minus_central(P, S.h, o, v, nv, w);
      if (MOIST) rq = rho_mface(V, S.rho, o, v) * w * upwind_minus_central(P, qprof, o, v, nv, w);
    }
    flx[o] = r;
    if (MOIST) flq[o] = rq;
  }
  __syncthreads();
  FT* gT = Ytc + (size_t)h * P.ncf * 16 * nv;
  FT* gF = Ytf + (size_t)h * 16 * nf;
  if (v < nv) {
    gT[(0 * 16 + n) * nv + v] = FT(0); gT[(1 * 16 + n) * nv + v] = FT(0); gT[(2 * 16 + n) * nv + v] = FT(0);
    gT[(3 * 16 + n) * nv + v] = -(flx[o + 1] - flx[o]) / V.mc[v];
    for (int q = 4; q < P.ncf; ++q) gT[(q * 16 + n) * nv + v] = FT(0);
    if (MOIST) gT[(4 * 16 + n) * nv + v] = -(flq[o + 1] - flq[o]) / V.mc[v];  // implicit_tendency.jl:333-338
  }
  if (v < nf) gF[n * nf + v] = FT(0);
}

// ---------------------------------------------------------------------------------------------
// Fused implicit stage WITH implicit vertical diffusion (the default for implicit_diffusion since round 2; parity-green on a B200):
// what b200_implicit_stage / k5_imp_stage do for the dry Jacobian, extended by the diffusion tendency, the diffusion blocks and the
// approximate arrowhead iteration — cache_imp! (u₃ filter) → Wfact → R = dtγ·T_imp(U) → ldiv! → N = U − ΔU → cache_imp! → T_post_imp! —
// in ONE kernel, out of place.  Quarter element per CTA, one point per thread; every coefficient profile lives in shared memory
// (48 profiles of 4·LVP words: 50 KB in Float32), tridiagonal solves by parallel cyclic reduction with one row per thread.  Algebra
// as in k_ldiv_diff (T-based Schur operator and preconditioner).
constexpr int QD_PROFILES = 48;

// PCR over the 4 columns of the CTA, one row per thread: thread (column, v) owns row v at index o.
template <class FT>
__device__ __forceinline__ void pcr_q(const FT* l, const FT* d, const FT* u, FT* x, int n, int o, int v, FT* wa, FT* wb, FT* wc) {
  __syncthreads();
  const bool act = v < n;
  if (act) { wa[o] = l[o]; wb[o] = d[o]; wc[o] = u[o]; }
  __syncthreads();
  for (int s = 1; s < n; s <<= 1) {
    FT A = FT(0), B = FT(1), C = FT(0), Dd = FT(0);
    if (act) {
      B = wb[o]; Dd = x[o];
      if (v - s >= 0) { const FT r = wa[o] / wb[o - s]; A = -r * wa[o - s]; B -= r * wc[o - s]; Dd -= r * x[o - s]; }
      if (v + s < n) { const FT r = wc[o] / wb[o + s]; C = -r * wc[o + s]; B -= r * wa[o + s]; Dd -= r * x[o + s]; }
    }
    __syncthreads();
    if (act) { wa[o] = A; wb[o] = B; wc[o] = C; x[o] = Dd; }
    __syncthreads();
  }
  if (act) x[o] = x[o] / wb[o];
  __syncthreads();
}

template <class FT>
__global__ void __launch_bounds__(NT) k_imp_stage_diff(Par<FT> P, VDiff<FT> D, const FT* __restrict__ hgeo,
                                                       const VLev<FT>* __restrict__ vlev, const FT* __restrict__ Uc,
                                                       const FT* __restrict__ Uf, FT* Nc, FT* Nf, FT dtg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int h = blockIdx.x >> 2, quarter = blockIdx.x & 3, n = quarter * 4 + (threadIdx.x >> 6), v = threadIdx.x & 63;
  const int nv = P.nv, nf = nv + 1, o = n * LVP + v;
  const bool cen = v < nv, fac_ = v < nf, inner = v > 0 && v < nv;
  const VLev<FT>& V = *vlev;
  const FT* hg = hgeo + (size_t)h * HG_N * 16;
  FT* base = reinterpret_cast<FT*>(smem_raw);
  int np = 0;
  auto take = [&]() { FT* p = base + (np++) * 4 * LVP - quarter * 4 * LVP; return p; };
  ImpSlabs<FT> S;
  S.rho = take(); S.u1 = take(); S.u2 = take(); S.re = take(); S.u3 = take(); S.K = take(); S.h = take(); S.Pi = take();
  S.thv = take(); S.thp = take(); S.phr = take(); S.T = take();
  FT *pw = take(), *ir = take(), *rr = take(), *rre = take(), *r1 = take(), *r2 = take();
  FT *sl = take(), *sd = take(), *su = take();
  FT *url = take(), *urh = take(), *uel = take(), *ueh = take(), *u1l = take(), *u1h = take(), *u2l = take(), *u2h = take();
  FT *rul = take(), *ruh = take(), *eul = take(), *euh = take();
  FT *pl_ = take(), *pd = take(), *pu = take(), *el = take(), *ed = take(), *eu = take(), *fc = take();
  FT *wa = take(), *wb = take(), *wc = take(), *ye = take(), *zz = take(), *b3 = take(), *x3 = take(), *r3 = take();
  // ---- cache_imp!(U): stage the columns, u₃ boundary filter on load, thermodynamics
  const FT* gU = Uc + (size_t)h * P.ncf * 16 * nv;
  FT* gN = Nc + (size_t)h * P.ncf * 16 * nv;
  if (cen) {
    S.rho[o] = gU[(0 * 16 + n) * nv + v]; S.u1[o] = gU[(1 * 16 + n) * nv + v];
    S.u2[o] = gU[(2 * 16 + n) * nv + v]; S.re[o] = gU[(3 * 16 + n) * nv + v];
  }
  if (fac_) S.u3[o] = inner ? Uf[(size_t)h * 16 * nf + (size_t)n * nf + v] : FT(0);
  __syncthreads();
  if (cen) {
    FT K = kinetic(hg, V, S.u1[o], S.u2[o], S.u3[o], S.u3[o + 1], n, v);
    Pt<FT> t = thermo(P, S.rho[o], S.re[o], K, V.phic[v]);
    S.K[o] = K; S.h[o] = t.h; S.Pi[o] = t.Pi; S.thv[o] = t.thv; S.thp[o] = t.thp; S.phr[o] = pgf_aux(t); S.T[o] = t.T;
  }
  __syncthreads();
  // ---- eddy diffusivity → zz
  if (cen) {
    FT kh;
    if (D.mode == 2) {
      kh = D.kdec[v];
    } else {
      const FT a = S.u1[n * LVP], b = S.u2[n * LVP];
      const FT g11 = hg[HG_GI11 * 16 + n], g12 = hg[HG_GI12 * 16 + n], g22 = hg[HG_GI22 * 16 + n];
      const FT nrm = sqrt((a * (g11 * a + g12 * b) + b * (g12 * a + g22 * b)) * V.sc2i[0]);
      const FT KE = D.ce_za * nrm;
      const FT p = S.rho[o] * P.R_d * S.T[o];
      const FT x = (FT(85000) - p) / FT(10000);
      kh = p > FT(85000) ? KE : KE * exp_(-(x * x));
    }
    zz[o] = kh;
  }
  __syncthreads();
  // ---- dtγ·(J g³³ ᶠρK)/J2, 1/ρ; advective residuals and centre-row coefficients
  FT rre0 = FT(0);
  if (fac_) {
    FT w = FT(0);
    if (inner) {
      const FT rf = FT(0.5) * (S.rho[o - 1] + S.rho[o]);
      const FT ik = FT(0.5) * (FT(1) / fmax_(zz[o - 1], D.eps) + FT(1) / fmax_(zz[o], D.eps));
      w = dtg * (V.dzf[v] * V.g33f[v] / V.sf2i[v]) * (rf / ik);
    }
    pw[o] = w;
  }
  if (cen) {
    ir[o] = FT(1) / S.rho[o];
    FT rt, et; timp_center(V, S, n, v, nv, rt, et);
    rr[o] = dtg * rt; rre0 = dtg * et;
    FT a, b, c, d; center_coef(V, S, dtg, n, v, nv, a, b, c, d);
    rul[o] = a; ruh[o] = b; eul[o] = c; euh[o] = d;
  }
  __syncthreads();
  // ---- diffusive residuals (pw carries dtγ), face-row coefficients, A_ee, (uₕ,uₕ)
  const bool lo = v > 0, hi = v < nv - 1;
  const FT wl = (cen && lo) ? pw[o] : FT(0), wh = (cen && hi) ? pw[o + 1] : FT(0);
  const FT rm = cen ? V.rmc[v] : FT(0);
  if (cen) {
    {
      const FT s0 = P.cp_d * (S.T[o] - P.T_0) + V.phic[v];
      const FT fl = lo ? wl * (s0 - (P.cp_d * (S.T[o - 1] - P.T_0) + V.phic[v - 1])) : FT(0);
      const FT fh = hi ? wh * ((P.cp_d * (S.T[o + 1] - P.T_0) + V.phic[v + 1]) - s0) : FT(0);
      rre[o] = rre0 + (fh - fl) * rm;
    }
    FT q1 = FT(0), q2 = FT(0);
    if (D.momentum) {
      const FT is0 = sqrt(V.sc2i[v]);
      const FT isl = lo ? sqrt(V.sc2i[v - 1]) : FT(0), ish = hi ? sqrt(V.sc2i[v + 1]) : FT(0);
      const FT sir = rm / (is0 * S.rho[o]);
      { const FT c0 = S.u1[o] * is0; const FT fl = lo ? wl * (c0 - S.u1[o - 1] * isl) : FT(0), fh = hi ? wh * (S.u1[o + 1] * ish - c0) : FT(0); q1 = (fh - fl) * sir; }
      { const FT c0 = S.u2[o] * is0; const FT fl = lo ? wl * (c0 - S.u2[o - 1] * isl) : FT(0), fh = hi ? wh * (S.u2[o + 1] * ish - c0) : FT(0); q2 = (fh - fl) * sir; }
    }
    r1[o] = q1; r2[o] = q2;
    const FT l_ = lo ? pw[o] * rm : FT(0), h_ = hi ? pw[o + 1] * rm : FT(0);
    const FT dg = -(l_ + h_);
    const FT m = dg * (D.cpcv * ir[o]);
    el[o] = lo ? l_ * (D.cpcv * ir[o - 1]) : FT(0);
    ed[o] = m - FT(1);
    eu[o] = hi ? h_ * (D.cpcv * ir[o + 1]) : FT(0);
    fc[o] = m / (m - FT(1));
    pl_[o] = l_ * ir[o]; pd[o] = dg * ir[o] - FT(1); pu[o] = h_ * ir[o];
    ye[o] = rre[o];
  }
  if (fac_) {
    FaceCoef<FT> c = face_coef(P, hg, V, S, dtg, n, v, nv);
    sl[o] = c.l; sd[o] = c.d; su[o] = c.u;
    url[o] = c.ur_lo; urh[o] = c.ur_hi; uel[o] = c.ue_lo; ueh[o] = c.ue_hi;
    u1l[o] = c.u1_lo; u1h[o] = c.u1_hi; u2l[o] = c.u2_lo; u2h[o] = c.u2_hi;
    b3[o] = dtg * timp_face(P, V, S, n, v, nv);
  }
  // ---- Δuₕ (exact tridiagonal solves, or −R = 0 without momentum diffusion) and y_e = A_ee⁻¹ R_ρe
  if (D.momentum) {
    pcr_q(pl_, pd, pu, r1, nv, o, v, wa, wb, wc);
    pcr_q(pl_, pd, pu, r2, nv, o, v, wa, wb, wc);
  }
  pcr_q(el, ed, eu, ye, nv, o, v, wa, wb, wc);
  // ---- Schur right-hand side and preconditioner
  if (fac_) {
    FT rhs = b3[o], l = sl[o], d = sd[o], u = su[o];
    if (inner) {
      rhs += url[o] * rr[o - 1] + urh[o] * rr[o];
      rhs -= uel[o] * ye[o - 1] + ueh[o] * ye[o];
      rhs -= u1l[o] * r1[o - 1] + u1h[o] * r1[o];
      rhs -= u2l[o] * r2[o - 1] + u2h[o] * r2[o];
      const FT a = uel[o] * fc[o - 1], b = ueh[o] * fc[o];
      l -= a * eul[o - 1];
      d -= a * euh[o - 1] + b * eul[o];
      u -= b * euh[o];
    }
    b3[o] = rhs; x3[o] = rhs;
    pl_[o] = l; pd[o] = d; pu[o] = u;
  }
  pcr_q(pl_, pd, pu, x3, nf, o, v, wa, wb, wc);
  for (int it = 0; it < D.n_iters; ++it) {
    if (cen) { const FT y = eul[o] * x3[o] + euh[o] * x3[o + 1]; ye[o] = y; zz[o] = y; }
    pcr_q(el, ed, eu, zz, nv, o, v, wa, wb, wc);
    if (fac_) {
      FT tx = sd[o] * x3[o];
      if (v > 0) tx += sl[o] * x3[o - 1];
      if (v < nv) tx += su[o] * x3[o + 1];
      FT r = b3[o] - tx;
      if (inner) r += uel[o] * (zz[o - 1] + ye[o - 1]) + ueh[o] * (zz[o] + ye[o]);
      r3[o] = r;
    }
    pcr_q(pl_, pd, pu, r3, nf, o, v, wa, wb, wc);
    if (fac_) x3[o] += r3[o];
    __syncthreads();
  }
  // ---- Δρ, Δρe_tot and the Newton update N = U − ΔU (profiles of S become N)
  FT drho = FT(0);
  if (cen) {
    const FT x0 = x3[o], x1 = x3[o + 1];
    drho = rul[o] * x0 + ruh[o] * x1 - rr[o];
    ye[o] = rre[o] - (eul[o] * x0 + euh[o] * x1);
  }
  pcr_q(el, ed, eu, ye, nv, o, v, wa, wb, wc);
  // passive tracers: Δ(ρχ) = (dtγ D⋅Diag(1/ρ) − I)⁻¹ dtγ D χ with the OLD 1/ρ (ir) — before S.rho is overwritten nothing else needs it
  for (int q = 4; q < P.ncf; ++q) {
    __syncthreads();
    if (cen) {
      const FT l_ = lo ? pw[o] * rm : FT(0), h_ = hi ? pw[o + 1] * rm : FT(0);
      pl_[o] = lo ? l_ * ir[o - 1] : FT(0); pd[o] = -(l_ + h_) * ir[o] - FT(1); pu[o] = hi ? h_ * ir[o + 1] : FT(0);
      r3[o] = gU[(size_t)(q * 16 + n) * nv + v] * ir[o];  // χ
    }
    __syncthreads();
    if (cen) {
      const FT c0 = r3[o];
      const FT fl = lo ? wl * (c0 - r3[o - 1]) : FT(0), fh = hi ? wh * (r3[o + 1] - c0) : FT(0);
      zz[o] = (fh - fl) * rm;
    }
    pcr_q(pl_, pd, pu, zz, nv, o, v, wa, wb, wc);
    if (cen) gN[(size_t)(q * 16 + n) * nv + v] = gU[(size_t)(q * 16 + n) * nv + v] - zz[o];
  }
  __syncthreads();
  if (cen) {
    S.rho[o] = S.rho[o] - drho; S.re[o] = S.re[o] - ye[o];
    S.u1[o] = S.u1[o] - r1[o]; S.u2[o] = S.u2[o] - r2[o];
  }
  if (fac_) S.u3[o] = inner ? S.u3[o] - x3[o] : FT(0);
  __syncthreads();
  // ---- cache_imp!(N), T_post_imp!
  FT e_new = cen ? S.re[o] : FT(0);
  if (P.upwinding != 0) {
    if (cen) {
      FT K = kinetic(hg, V, S.u1[o], S.u2[o], S.u3[o], S.u3[o + 1], n, v);
      Pt<FT> t = thermo(P, S.rho[o], S.re[o], K, V.phic[v]);
      S.h[o] = t.h;
    }
    __syncthreads();
    if (fac_) {
      FT r = FT(0);
      if (inner) {
        const FT w = V.g33f[v] * S.u3[o];
        r = rho_mface(V, S.rho, o, v) * w * upwind_minus_central(P, S.h, o, v, nv, w);
      }
      b3[o] = r;
    }
    __syncthreads();
    if (cen) e_new += dtg * (-(b3[o + 1] - b3[o]) / V.mc[v]);
  }
  if (cen) {
    gN[(0 * 16 + n) * nv + v] = S.rho[o]; gN[(1 * 16 + n) * nv + v] = S.u1[o];
    gN[(2 * 16 + n) * nv + v] = S.u2[o]; gN[(3 * 16 + n) * nv + v] = e_new;
  }
  if (fac_) Nf[(size_t)h * 16 * nf + (size_t)n * nf + v] = S.u3[o];
}

}  // namespace b200
