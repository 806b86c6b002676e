// CPU emulation of the explicit-tendency kernels of the benchmarked step, k5_exp_a (kernels_pair.cuh) and k7_exp_c (kernels_lvl.cuh; Float64
// instantiations), from their unchanged source, on the fiber CTA emulator of cuda_runtime.h with its emulated warp shuffles (the 32
// lanes of a warp meet at a per-warp barrier, publish their value, and read the source lane's).  griddepcontrol.* assembles to nothing; the packed-Float32 PTX is not instantiated.
// Test infrastructure only (tests/test_kernels_cpu_emulation.py).
#include <vector>
#define b200 b200_emux
#define EMU_WARP_INTRINSICS
#include "cuda_runtime.h"
thread_local uint3_emu threadIdx, blockIdx;
namespace b200 { alignas(128) unsigned char smem_raw[256 * 1024]; }
#define __constant__
__asm__(".macro griddepcontrol.launch_dependents\n.endm\n.macro griddepcontrol.wait\n.endm");
#include "kernels_implicit.cuh"
#include "kernels_row.cuh"
#include "kernels_row.cuh"
#include "kernels_pair.cuh"
#include "kernels_lvl.cuh"

using namespace b200;
typedef double FT;
#include "emu_moist.h"

template <class F>
static void run_grid(int nx, int ny, F&& body) {
  const std::function<void()> fn = body;
  for (int y = 0; y < ny; ++y)
    for (int b = 0; b < nx; ++b) {
      blockIdx = {(unsigned)b, (unsigned)y, 0};
      emu::run_cta(256, fn);
    }
}

// sc: R_d, cp_d, cv_d, T_0, p_ref_theta, T_surf_ref, T_min_ref, T_min_sgs, dt, ν₄ᵥ, ν₄ₛ, divergence damping factor, hyperdiff, rayleigh, viscous,
//     energy upwinding, ncf, tracer upwinding ; vl: [14][64] = sc2i, sf2i, sf, dzc, dzf, mc, rmc, g33f, phic, dphif, brw, bruh, bvc, bvf ; D [16], w [4]
// which: 0 = k5_exp_a (writes Ytc, Ytf, H), 1 = k7_exp_c (kernels_lvl.cuh: reads H, updates Ytc, Ytf; one thread per (element, level), 4 elements per CTA),
//        2 / 3 = k5_tracer_a / the tracer parts of k7_exp_c
extern "C" __attribute__((visibility("default"))) int emu_exp5(int which, int nh, int nv, const double* sc, const double* vl, const double* Dm,
                                                               const double* w, const double* hgeo, const double* Yc, const double* Yf,
                                                               double* Ytc, double* Ytf, double* H, double* Ylc) {
  Par<FT> P;
  memset(&P, 0, sizeof(P));
  P.R_d = sc[0]; P.cp_d = sc[1]; P.cv_d = sc[2]; P.T_0 = sc[3]; P.p0 = sc[4]; P.kappa = sc[0] / sc[1]; P.Ts_ref = sc[5];
  P.Tmin_ref = sc[6]; P.T_min_sgs = sc[7]; P.dt = sc[8]; P.icv = 1.0 / sc[2]; P.ip0 = 1.0 / sc[4]; P.dTs7 = (sc[5] - sc[6]) / 7.0;
  P.RT0 = sc[0] * sc[3]; P.nu4v = sc[9]; P.nu4s = sc[10]; P.ddf = sc[11]; P.hyperdiff = (int)sc[12]; P.rayleigh = (int)sc[13];
  P.viscous = (int)sc[14]; P.upwinding = (int)sc[15]; P.tupw = 3; P.nh = nh; P.nv = nv; P.ncf = (int)sc[16];
  static VLev<FT> V;
  memset(&V, 0, sizeof(V));
  FT* dst[14] = {V.sc2i, V.sf2i, V.sf, V.dzc, V.dzf, V.mc, V.rmc, V.g33f, V.phic, V.dphif, V.brw, V.bruh, V.bvc, V.bvf};
  for (int a = 0; a < 14; ++a) memcpy(dst[a], vl + a * 64, 64 * sizeof(FT));
  // derivative matrices as capi.cu:create_geo stores them: D, the weak form Dw[i][k] = −D[k][i] w_k / w_i, and the paired layout
  double md[32];
  for (int i = 0; i < 4; ++i)
    for (int k = 0; k < 4; ++k) {
      V.D[i * 4 + k] = Dm[i * 4 + k];
      V.Dw[i * 4 + k] = -Dm[k * 4 + i] * w[k] / w[i];
    }
  for (int k = 0; k < 16; ++k) { md[k] = V.D[k]; md[16 + k] = V.Dw[k]; c_Dd[k] = md[k]; c_Dd[16 + k] = md[16 + k]; }
  for (int ww = 0; ww < 2; ++ww)
    for (int k = 0; k < 4; ++k)
      for (int pp = 0; pp < 2; ++pp) {
        c_Pd[(ww * 4 + k) * 2 + pp].x = md[ww * 16 + (2 * pp) * 4 + k];
        c_Pd[(ww * 4 + k) * 2 + pp].y = md[ww * 16 + (2 * pp + 1) * 4 + k];
      }
  P.tupw = (int)sc[17];
  emu_apply_moist(P, sc[3]);
  const int ntr = P.ncf - 4;
  if (which == 0 && g_moist_on) run_grid(nh, 1, [&] { k5_exp_a<FT, 0, true>(P, hgeo, &V, Yc, Yf, Ytc, Ytf, H, (FT*)g_moist_Hw, Ylc); });
  // (moist: part 1 of k7_exp_c also applies the water terms — reads H[4] and Hw; updates ρe_tot of Ytc and ρ, ρq_tot of Ylc, or of Ytc when Ylc is NULL)
  else if (which == 0) run_grid(nh, 1, [&] { k5_exp_a<FT, 0>(P, hgeo, &V, Yc, Yf, Ytc, Ytf, H); });
  else if (which == 1) run_grid((nh + LVL_EPB - 1) / LVL_EPB, 3, [&] { k7_exp_c<FT, 0>(P, hgeo, &V, Yc, H, Ytc, Ytf, Ylc ? Ylc : Ytc, (const FT*)g_moist_Hw); });
  // passive tracers (grid = elements × tracers): 2 = k5_tracer_a (Yₜ, Yₜ_lim, ∇²χ → H), 3 = parts 3.. of k7_exp_c (tracer hyperdiffusion → Yₜ_lim)
  else if (which == 2) { if (ntr - g_moist_on > 0) run_grid(nh, ntr - g_moist_on, [&] { k5_tracer_a<FT>(P, hgeo, &V, Yc, Yf, Ytc, Ylc, H); }); }
  else if (ntr - g_moist_on > 0) {  // parts 3.. of k7_exp_c only (run_grid's y index is offset by the three dry parts)
    const int nb = (nh + LVL_EPB - 1) / LVL_EPB;
    const std::function<void()> fn = [&] { k7_exp_c<FT, 0>(P, hgeo, &V, Yc, H, Ytc, Ytf, Ylc); };
    for (int y = 3; y < 3 + ntr - g_moist_on; ++y)
      for (int b = 0; b < nb; ++b) { blockIdx = {(unsigned)b, (unsigned)y, 0}; emu::run_cta(256, fn); }
  }
  return 0;
}
