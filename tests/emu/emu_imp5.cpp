// CPU emulation of the DEFAULT fused implicit-stage kernel k5_imp_stage (kernels_imp5.cuh; Float64 instantiation, PCR column solve) —
// the second-largest kernel of the benchmarked step — with the same CTA emulator as emu_vdiff.cpp.  The kernel source is compiled
// unchanged; griddepcontrol.* assembles to nothing (empty assembler macros), and the packed-Float32 PTX belongs to specialisations
// that are not instantiated here.  Test infrastructure only.
#include <vector>
#define b200 b200_emu5
#define EMU_WARP_INTRINSICS
#include "cuda_runtime.h"
thread_local uint3_emu threadIdx, blockIdx;
namespace b200 { alignas(16) unsigned char smem_raw[256 * 1024]; }
#define __constant__
#define FULL_MASK_EMU 0xffffffffu
// griddepcontrol.* (programmatic dependent launch) is a no-op for a single emulated grid: teach the assembler two empty macros so that
// the inline PTX statements of common.cuh assemble to nothing
__asm__(".macro griddepcontrol.launch_dependents\n.endm\n.macro griddepcontrol.wait\n.endm");
#include "kernels_imp5.cuh"
#include "kernels_imp8.cuh"
#include "kernels_imp8d.cuh"

using namespace b200;
typedef double FT;
#include "emu_moist.h"

template <class F>
static void run_grid(int nblocks, F&& body) {
  const std::function<void()> fn = body;
  for (int b = 0; b < nblocks; ++b) {
    blockIdx = {(unsigned)b, 0, 0};
    emu::run_cta(256, fn);
  }
}

// sc: R_d, cp_d, cv_d, T_0, p_ref_theta, T_surf_ref, T_min_ref, T_min_sgs, dt, rayleigh, dtγ, energy upwinding, ncf
extern "C" __attribute__((visibility("default"))) int emu_imp5(int nh, int nv, const double* sc, const double* vl, const double* hgeo,
                                                               const double* Uc, const double* Uf, double* Nc, double* Nf) {
  Par<FT> P;
  memset(&P, 0, sizeof(P));
  P.R_d = sc[0]; P.cp_d = sc[1]; P.cv_d = sc[2]; P.T_0 = sc[3]; P.p0 = sc[4]; P.kappa = sc[0] / sc[1]; P.Ts_ref = sc[5];
  P.Tmin_ref = sc[6]; P.T_min_sgs = sc[7]; P.dt = sc[8]; P.icv = 1.0 / sc[2]; P.ip0 = 1.0 / sc[4]; P.dTs7 = (sc[5] - sc[6]) / 7.0;
  P.RT0 = sc[0] * sc[3]; P.nh = nh; P.nv = nv; P.ncf = (int)sc[12]; P.rayleigh = (int)sc[9]; P.upwinding = (int)sc[11];
  static VLev<FT> V;
  memset(&V, 0, sizeof(V));
  FT* dst[11] = {V.sc2i, V.sf2i, V.sf, V.dzc, V.dzf, V.mc, V.rmc, V.g33f, V.phic, V.dphif, V.brw};
  for (int a = 0; a < 11; ++a) memcpy(dst[a], vl + a * 64, 64 * sizeof(FT));
  emu_apply_moist(P, sc[3]);
  if (g_moist_on) run_grid(nh, [&] { k5_imp_stage<FT, 0, false, true>(P, hgeo, &V, Uc, Uf, Nc, Nf, (FT)sc[10]); });
  else run_grid(nh, [&] { k5_imp_stage<FT, 0>(P, hgeo, &V, Uc, Uf, Nc, Nf, (FT)sc[10]); });
  return 0;
}

// ldiv!(ΔY, J(Y, dtγ), R) of the hook path: k5_imp_stage<…, LDIV = true> on the state Wfact was called with (same sc / vl / hgeo)
extern "C" __attribute__((visibility("default"))) int emu_ldiv5(int nh, int nv, const double* sc, const double* vl, const double* hgeo,
                                                                const double* Yc, const double* Yf, const double* Rc, const double* Rf,
                                                                double* dYc, double* dYf) {
  Par<FT> P;
  memset(&P, 0, sizeof(P));
  P.R_d = sc[0]; P.cp_d = sc[1]; P.cv_d = sc[2]; P.T_0 = sc[3]; P.p0 = sc[4]; P.kappa = sc[0] / sc[1]; P.Ts_ref = sc[5];
  P.Tmin_ref = sc[6]; P.T_min_sgs = sc[7]; P.dt = sc[8]; P.icv = 1.0 / sc[2]; P.ip0 = 1.0 / sc[4]; P.dTs7 = (sc[5] - sc[6]) / 7.0;
  P.RT0 = sc[0] * sc[3]; P.nh = nh; P.nv = nv; P.ncf = (int)sc[12]; P.rayleigh = (int)sc[9]; P.upwinding = (int)sc[11];
  static VLev<FT> V;
  memset(&V, 0, sizeof(V));
  FT* dst[11] = {V.sc2i, V.sf2i, V.sf, V.dzc, V.dzf, V.mc, V.rmc, V.g33f, V.phic, V.dphif, V.brw};
  for (int a = 0; a < 11; ++a) memcpy(dst[a], vl + a * 64, 64 * sizeof(FT));
  emu_apply_moist(P, sc[3]);
  if (g_moist_on) run_grid(nh, [&] { k5_imp_stage<FT, 0, true, true>(P, hgeo, &V, Yc, Yf, dYc, dYf, (FT)sc[10], Rc, Rf); });
  else run_grid(nh, [&] { k5_imp_stage<FT, 0, true>(P, hgeo, &V, Yc, Yf, dYc, dYf, (FT)sc[10], Rc, Rf); });
  return 0;
}

// k8_imp_stage_diff (kernels_imp8d.cuh): the fused implicit stage with implicit vertical diffusion, warp per column pair.
// sc as emu_imp5 plus sc[13] = diffusion mode (1 | 2), sc[14] = momentum diffusion, sc[15] = n_iters, sc[16] = C_E·Δz₁/2; kdec [64]
extern "C" __attribute__((visibility("default"))) int emu_imp5d(int nh, int nv, const double* sc, const double* vl, const double* hgeo, const double* kdec,
                                                                const double* Uc, const double* Uf, double* Nc, double* Nf) {
  Par<FT> P;
  memset(&P, 0, sizeof(P));
  P.R_d = sc[0]; P.cp_d = sc[1]; P.cv_d = sc[2]; P.T_0 = sc[3]; P.p0 = sc[4]; P.kappa = sc[0] / sc[1]; P.Ts_ref = sc[5];
  P.Tmin_ref = sc[6]; P.T_min_sgs = sc[7]; P.dt = sc[8]; P.icv = 1.0 / sc[2]; P.ip0 = 1.0 / sc[4]; P.dTs7 = (sc[5] - sc[6]) / 7.0;
  P.RT0 = sc[0] * sc[3]; P.nh = nh; P.nv = nv; P.ncf = (int)sc[12]; P.rayleigh = (int)sc[9]; P.upwinding = (int)sc[11];
  VDiff<FT> D;
  D.mode = (int)sc[13]; D.momentum = (int)sc[14]; D.n_iters = (int)sc[15]; D.ce_za = sc[16]; D.eps = 2.220446049250313e-16; D.cpcv = sc[1] / sc[2];
  D.kdec = kdec;
  static VLev<FT> V;
  memset(&V, 0, sizeof(V));
  FT* dst[11] = {V.sc2i, V.sf2i, V.sf, V.dzc, V.dzf, V.mc, V.rmc, V.g33f, V.phic, V.dphif, V.brw};
  for (int a = 0; a < 11; ++a) memcpy(dst[a], vl + a * 64, 64 * sizeof(FT));
  run_grid(nh, [&] { k8_imp_stage_diff<FT, 0>(P, D, hgeo, &V, Uc, Uf, Nc, Nf, (FT)sc[10]); });
  return 0;
}

// k8_imp_stage (kernels_imp8.cuh): the dry fused implicit stage with a warp per column pair (no shared memory, shuffle PCR); sc as emu_imp5
extern "C" __attribute__((visibility("default"))) int emu_imp8(int nh, int nv, const double* sc, const double* vl, const double* hgeo,
                                                               const double* Uc, const double* Uf, double* Nc, double* Nf) {
  Par<FT> P;
  memset(&P, 0, sizeof(P));
  P.R_d = sc[0]; P.cp_d = sc[1]; P.cv_d = sc[2]; P.T_0 = sc[3]; P.p0 = sc[4]; P.kappa = sc[0] / sc[1]; P.Ts_ref = sc[5];
  P.Tmin_ref = sc[6]; P.T_min_sgs = sc[7]; P.dt = sc[8]; P.icv = 1.0 / sc[2]; P.ip0 = 1.0 / sc[4]; P.dTs7 = (sc[5] - sc[6]) / 7.0;
  P.RT0 = sc[0] * sc[3]; P.nh = nh; P.nv = nv; P.ncf = (int)sc[12]; P.rayleigh = (int)sc[9]; P.upwinding = (int)sc[11];
  static VLev<FT> V;
  memset(&V, 0, sizeof(V));
  FT* dst[11] = {V.sc2i, V.sf2i, V.sf, V.dzc, V.dzf, V.mc, V.rmc, V.g33f, V.phic, V.dphif, V.brw};
  for (int a = 0; a < 11; ++a) memcpy(dst[a], vl + a * 64, 64 * sizeof(FT));
  emu_apply_moist(P, sc[3]);
  if (g_moist_on) run_grid(nh, [&] { k8_imp_stage<FT, 0, false, true>(P, hgeo, &V, Uc, Uf, Nc, Nf, (FT)sc[10]); });
  else run_grid(nh, [&] { k8_imp_stage<FT, 0>(P, hgeo, &V, Uc, Uf, Nc, Nf, (FT)sc[10]); });
  return 0;
}
// ldiv! in the same layout (k8_imp_stage<…, LDIV>)
extern "C" __attribute__((visibility("default"))) int emu_ldiv8(int nh, int nv, const double* sc, const double* vl, const double* hgeo,
                                                                const double* Yc, const double* Yf, const double* Rc, const double* Rf,
                                                                double* dYc, double* dYf) {
  Par<FT> P;
  memset(&P, 0, sizeof(P));
  P.R_d = sc[0]; P.cp_d = sc[1]; P.cv_d = sc[2]; P.T_0 = sc[3]; P.p0 = sc[4]; P.kappa = sc[0] / sc[1]; P.Ts_ref = sc[5];
  P.Tmin_ref = sc[6]; P.T_min_sgs = sc[7]; P.dt = sc[8]; P.icv = 1.0 / sc[2]; P.ip0 = 1.0 / sc[4]; P.dTs7 = (sc[5] - sc[6]) / 7.0;
  P.RT0 = sc[0] * sc[3]; P.nh = nh; P.nv = nv; P.ncf = (int)sc[12]; P.rayleigh = (int)sc[9]; P.upwinding = (int)sc[11];
  static VLev<FT> V;
  memset(&V, 0, sizeof(V));
  FT* dst[11] = {V.sc2i, V.sf2i, V.sf, V.dzc, V.dzf, V.mc, V.rmc, V.g33f, V.phic, V.dphif, V.brw};
  for (int a = 0; a < 11; ++a) memcpy(dst[a], vl + a * 64, 64 * sizeof(FT));
  emu_apply_moist(P, sc[3]);
  if (g_moist_on) run_grid(nh, [&] { k8_imp_stage<FT, 0, true, true>(P, hgeo, &V, Yc, Yf, dYc, dYf, (FT)sc[10], Rc, Rf); });
  else run_grid(nh, [&] { k8_imp_stage<FT, 0, true>(P, hgeo, &V, Yc, Yf, dYc, dYf, (FT)sc[10], Rc, Rf); });
  return 0;
}

// k8_t_imp / k8_t_post_imp (kernels_imp8.cuh): T_imp! and T_post_imp! of the hook path in the warp-per-column-pair layout; sc as emu_imp5
extern "C" __attribute__((visibility("default"))) int emu_hooks8(int nh, int nv, const double* sc, const double* vl, const double* hgeo,
                                                                 const double* Yc, const double* Yf, double* Ytc, double* Ytf, double* Ypc, double* Ypf) {
  Par<FT> P;
  memset(&P, 0, sizeof(P));
  P.R_d = sc[0]; P.cp_d = sc[1]; P.cv_d = sc[2]; P.T_0 = sc[3]; P.p0 = sc[4]; P.kappa = sc[0] / sc[1]; P.Ts_ref = sc[5];
  P.Tmin_ref = sc[6]; P.T_min_sgs = sc[7]; P.dt = sc[8]; P.icv = 1.0 / sc[2]; P.ip0 = 1.0 / sc[4]; P.dTs7 = (sc[5] - sc[6]) / 7.0;
  P.RT0 = sc[0] * sc[3]; P.nh = nh; P.nv = nv; P.ncf = (int)sc[12]; P.rayleigh = (int)sc[9]; P.upwinding = (int)sc[11];
  static VLev<FT> V;
  memset(&V, 0, sizeof(V));
  FT* dst[11] = {V.sc2i, V.sf2i, V.sf, V.dzc, V.dzf, V.mc, V.rmc, V.g33f, V.phic, V.dphif, V.brw};
  for (int a = 0; a < 11; ++a) memcpy(dst[a], vl + a * 64, 64 * sizeof(FT));
  emu_apply_moist(P, sc[3]);
  if (g_moist_on) {
    run_grid(nh, [&] { k8_t_imp<FT, 0, true>(P, hgeo, &V, Yc, Yf, Ytc, Ytf); });
    run_grid(nh, [&] { k8_t_post_imp<FT, 0, true>(P, hgeo, &V, Yc, Yf, Ypc, Ypf); });
  } else {
    run_grid(nh, [&] { k8_t_imp<FT, 0>(P, hgeo, &V, Yc, Yf, Ytc, Ytf); });
    run_grid(nh, [&] { k8_t_post_imp<FT, 0>(P, hgeo, &V, Yc, Yf, Ypc, Ypf); });
  }
  return 0;
}
