// CPU emulation of single CTAs running the vertical-diffusion kernels of climaatmos.jl_b200/csrc/kernels_vdiff.cuh (and k_wfact2, whose
// planes k_ldiv_diff consumes) — the kernel source is compiled unchanged by g++ against the stub cuda_runtime.h in this directory;
// the threads of a block are the fibers of the CTA emulator in cuda_runtime.h.  Test infrastructure only (tests/test_kernels_cpu_emulation.py): it checks
// indexing, barriers-as-phases and arithmetic of the kernels against the oracle when no GPU is at hand.  It is NOT a product path.
#include <vector>

// The product library exports host stubs with the kernels' mangled names; when both are loaded in one process the dynamic linker
// would bind the emulator's calls to those stubs.  A private namespace keeps the two apart.
#define b200 b200_emu
#include "cuda_runtime.h"
thread_local uint3_emu threadIdx, blockIdx;
namespace b200 { alignas(16) unsigned char smem_raw[256 * 1024]; }

#include "kernels_vdiff.cuh"
#include "kernels_limiter.cuh"

using namespace b200;
#ifndef EMU_FT
#define EMU_FT double
#endif
typedef EMU_FT FT;  // -DEMU_FT=float builds the Float32 instantiations (arrays are FT, the scalar block sc stays double)
#include "emu_moist.h"

template <class F>
static void run_grid(int nblocks, F&& body) {
  const std::function<void()> fn = body;
  for (int b = 0; b < nblocks; ++b) {  // one CTA after the other: the next block reuses the shared memory
    blockIdx = {(unsigned)b, 0, 0};
    emu::run_cta(NT, fn);
  }
}

// sc: R_d, cp_d, cv_d, T_0, p_ref_theta, T_surf_ref, T_min_ref, T_min_sgs, dt, rayleigh(0/1), mode, momentum, n_iters, C_E·Δz₁/2, dtγ, [15]-[17] unused here
// vl: [11][64] = sc2i, sf2i, sf, dzc, dzf, mc, rmc, g33f, phic, dphif, brw ; hgeo: [nh][HG_N][16] ; kdec [64]
extern "C" __attribute__((visibility("default"))) int emu_vdiff(int nh, int nv, int ncf, const double* sc, const FT* vl, const FT* hgeo, const FT* kdec,
                         const FT* Yc, const FT* Yf, const FT* Rc, const FT* Rf, FT* Ytc, FT* jac,
                         FT* jacd, FT* dYc, FT* dYf) {
  Par<FT> P;
  memset(&P, 0, sizeof(P));
  P.R_d = sc[0]; P.cp_d = sc[1]; P.cv_d = sc[2]; P.T_0 = sc[3]; P.p0 = sc[4]; P.kappa = sc[0] / sc[1]; P.Ts_ref = sc[5];
  P.Tmin_ref = sc[6]; P.T_min_sgs = sc[7]; P.dt = sc[8]; P.icv = 1.0 / sc[2]; P.ip0 = 1.0 / sc[4]; P.dTs7 = (sc[5] - sc[6]) / 7.0;
  P.RT0 = sc[0] * sc[3]; P.nh = nh; P.nv = nv; P.ncf = ncf; P.rayleigh = (int)sc[9];
  VDiff<FT> D;
  D.mode = (int)sc[10]; D.momentum = (int)sc[11]; D.n_iters = (int)sc[12]; D.ce_za = sc[13]; D.eps = sizeof(FT) == 4 ? (FT)1.1920928955078125e-7 : (FT)2.220446049250313e-16;
  D.cpcv = sc[1] / sc[2]; D.kdec = kdec;
  const FT dtg = sc[14];
  static VLev<FT> V;
  memset(&V, 0, sizeof(V));
  FT* dst[11] = {V.sc2i, V.sf2i, V.sf, V.dzc, V.dzf, V.mc, V.rmc, V.g33f, V.phic, V.dphif, V.brw};
  for (int a = 0; a < 11; ++a) memcpy(dst[a], vl + a * 64, 64 * sizeof(FT));
  run_grid(nh * 4, [&] { k_vdiff_tend2<FT>(P, D, hgeo, &V, Yc, Yf, Ytc); });
  run_grid(nh * 4, [&] { k_wfact2<FT>(P, hgeo, &V, Yc, Yf, dtg, jac); });
  run_grid(nh, [&] { k_vdiff_jac<FT>(P, D, hgeo, &V, Yc, Yf, dtg, jacd); });
  run_grid(nh, [&] { k_ldiv_diff<FT>(P, D, &V, jac, jacd, Rc, Rf, dYc, dYf); });
  return 0;
}

// k_lim_vborrow on every (element, tracer): Yc in place; dzc [64]
extern "C" __attribute__((visibility("default"))) int emu_vborrow(int nh, int nv, int ncf, const FT* dzc, FT* Yc) {
  static VLev<FT> V;
  memset(&V, 0, sizeof(V));
  memcpy(V.dzc, dzc, 64 * sizeof(FT));
  const std::function<void()> fn = [&] { k_lim_vborrow<FT>(&V, Yc, ncf, nv, (FT)0); };
  for (int t = 0; t < ncf - 4; ++t)
    for (int b = 0; b < nh; ++b) {
      blockIdx = {(unsigned)b, (unsigned)t, 0};
      emu::run_cta(NT, fn);
    }
  return 0;
}

// The dry hook kernels (k_cache_imp of kernels_implicit.cuh; k_t_imp2, k_wfact2, k_t_post_imp2 of kernels_vdiff.cuh; validated
// on the B200, emulated here so that the CPU test tier exercises the product source too): cache_imp! → T_imp! → Wfact → ldiv! → T_post_imp!.
// sc as in emu_vdiff plus sc[16] = energy upwinding (0 | 1 | 3).  Sc/Sf are unused (kept for the call signature).  Outputs: Yf is filtered in place; Tc/pc/hc/Kc [nh][16][nv];
// Ytc/Ytf (T_imp), dYc/dYf (ldiv of Rc/Rf), Ypc (T_post_imp centres), Sc/Sf (state after the fused stage).
extern "C" __attribute__((visibility("default"))) int emu_hooks(int nh, int nv, int ncf, const double* sc, const FT* vl,
                                                                const FT* hgeo, const FT* Yc, FT* Yf, const FT* Rc,
                                                                const FT* Rf, FT* Kc, FT* Tc, FT* pc, FT* hc,
                                                                FT* Ytc, FT* Ytf, FT* jac, FT* dYc, FT* dYf,
                                                                FT* Ypc, FT* Ypf, FT* Sc, FT* Sf) {
  Par<FT> P;
  memset(&P, 0, sizeof(P));
  P.R_d = sc[0]; P.cp_d = sc[1]; P.cv_d = sc[2]; P.T_0 = sc[3]; P.p0 = sc[4]; P.kappa = sc[0] / sc[1]; P.Ts_ref = sc[5];
  P.Tmin_ref = sc[6]; P.T_min_sgs = sc[7]; P.dt = sc[8]; P.icv = 1.0 / sc[2]; P.ip0 = 1.0 / sc[4]; P.dTs7 = (sc[5] - sc[6]) / 7.0;
  P.RT0 = sc[0] * sc[3]; P.nh = nh; P.nv = nv; P.ncf = ncf; P.rayleigh = (int)sc[9]; P.upwinding = (int)sc[16];
  const FT dtg = sc[14];
  static VLev<FT> V;
  memset(&V, 0, sizeof(V));
  FT* dst[11] = {V.sc2i, V.sf2i, V.sf, V.dzc, V.dzf, V.mc, V.rmc, V.g33f, V.phic, V.dphif, V.brw};
  for (int a = 0; a < 11; ++a) memcpy(dst[a], vl + a * 64, 64 * sizeof(FT));
  emu_apply_moist(P, sc[3]);
  run_grid(nh, [&] { k_cache_imp<FT>(P, hgeo, &V, Yc, Yf, (FT*)nullptr, (FT*)nullptr, Kc, Tc, pc, hc); });
  if (g_moist_on) run_grid(nh * 4, [&] { k_t_imp2<FT, true>(P, hgeo, &V, Yc, Yf, Ytc, Ytf); });
  else run_grid(nh * 4, [&] { k_t_imp2<FT>(P, hgeo, &V, Yc, Yf, Ytc, Ytf); });
  if (!g_moist_on) run_grid(nh * 4, [&] { k_wfact2<FT>(P, hgeo, &V, Yc, Yf, dtg, jac); });  // debug planes: dry only
  (void)Rc; (void)Rf; (void)dYc; (void)dYf;  // ldiv! of the dry path is k5_imp_stage<…, LDIV>: emu_imp5.cpp (emu_ldiv5)
  if (g_moist_on) run_grid(nh * 4, [&] { k_t_post_imp2<FT, true>(P, hgeo, &V, Yc, Yf, Ypc, Ypf); });
  else run_grid(nh * 4, [&] { k_t_post_imp2<FT>(P, hgeo, &V, Yc, Yf, Ypc, Ypf); });
  (void)Sc; (void)Sf;
  return 0;
}

// k_imp_stage_diff: N ← fused implicit stage with implicit vertical diffusion of U (sc as in emu_vdiff, sc[16] = energy upwinding)
extern "C" __attribute__((visibility("default"))) int emu_stage_diff(int nh, int nv, int ncf, const double* sc, const FT* vl,
                                                                     const FT* hgeo, const FT* kdec, const FT* Uc,
                                                                     const FT* Uf, FT* Nc, FT* Nf) {
  Par<FT> P;
  memset(&P, 0, sizeof(P));
  P.R_d = sc[0]; P.cp_d = sc[1]; P.cv_d = sc[2]; P.T_0 = sc[3]; P.p0 = sc[4]; P.kappa = sc[0] / sc[1]; P.Ts_ref = sc[5];
  P.Tmin_ref = sc[6]; P.T_min_sgs = sc[7]; P.dt = sc[8]; P.icv = 1.0 / sc[2]; P.ip0 = 1.0 / sc[4]; P.dTs7 = (sc[5] - sc[6]) / 7.0;
  P.RT0 = sc[0] * sc[3]; P.nh = nh; P.nv = nv; P.ncf = ncf; P.rayleigh = (int)sc[9]; P.upwinding = (int)sc[16];
  VDiff<FT> D;
  D.mode = (int)sc[10]; D.momentum = (int)sc[11]; D.n_iters = (int)sc[12]; D.ce_za = sc[13]; D.eps = sizeof(FT) == 4 ? (FT)1.1920928955078125e-7 : (FT)2.220446049250313e-16;
  D.cpcv = sc[1] / sc[2]; D.kdec = kdec;
  static VLev<FT> V;
  memset(&V, 0, sizeof(V));
  FT* dst[11] = {V.sc2i, V.sf2i, V.sf, V.dzc, V.dzf, V.mc, V.rmc, V.g33f, V.phic, V.dphif, V.brw};
  for (int a = 0; a < 11; ++a) memcpy(dst[a], vl + a * 64, 64 * sizeof(FT));
  run_grid(nh * 4, [&] { k_imp_stage_diff<FT>(P, D, hgeo, &V, Uc, Uf, Nc, Nf, (FT)sc[14]); });
  return 0;
}

// lim!, first branch: k_lim_bounds then k_lim_apply (kernels_limiter.cuh; no barriers — the threads are simply run one after the other).
// nbr_off/nbr: vertex neighbours of every element (CSR); hgeo with HG_WJ filled; bounds from ref_c, limiter applied to Yc in place.
extern "C" __attribute__((visibility("default"))) int emu_sem_limiter(int nh, int nv, int ncf, const int* nbr_off, const int* nbr,
                                                                      const FT* hgeo, const FT* ref_c, FT* Yc) {
  const int ntr = ncf - 4;
  std::vector<FT> bnd((size_t)ntr * nh * 2 * LV, FT(0));
  for (int pass = 0; pass < 2; ++pass)
    for (int t = 0; t < ntr; ++t)
      for (int e = 0; e < nh; ++e)
        for (int v = 0; v < 64; ++v) {
          threadIdx = {(unsigned)v, 0, 0};
          blockIdx = {(unsigned)e, (unsigned)t, 0};
          if (pass == 0) k_lim_bounds<FT>(ref_c, ncf, nv, nh, bnd.data(), (FT*)nullptr, ntr);
          else k_lim_apply<FT>(Yc, ncf, nv, nh, bnd.data(), nbr_off, nbr, hgeo, (const FT*)nullptr, 0ll, (const int*)nullptr, (const int*)nullptr, ntr);
        }
  return 0;
}
