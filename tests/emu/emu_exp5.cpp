// CPU emulation of the explicit-tendency kernels of the benchmarked step, k5_exp_a (kernels_pair.cuh) and k7_exp_c (kernels_lvl.cuh; Float64
// instantiations), from their unchanged source.  On top of the CTA emulator of emu_vdiff.cpp this one emulates warp shuffles: the 32
// host threads of a warp meet at a per-warp barrier, publish their value, and read the source lane's (all shuffles of these kernels are
// executed by full, converged warps).  griddepcontrol.* assembles to nothing; the packed-Float32 PTX is not instantiated.
// Test infrastructure only (tests/test_kernels_cpu_emulation.py).
#include <thread>
#include <vector>
#define b200 b200_emux
#include "cuda_runtime.h"
thread_local uint3_emu threadIdx, blockIdx;
std::barrier<>* g_cta_barrier = nullptr;
namespace b200 { alignas(128) unsigned char smem_raw[256 * 1024]; }
#define __constant__
struct WarpX { std::barrier<> bar{32}; alignas(16) unsigned char buf[32][16]; };
static WarpX g_warp[8];
template <class T> inline T shfl_emu(T v, int src_lane) {
  WarpX& w = g_warp[threadIdx.x >> 5];
  const int l = threadIdx.x & 31;
  memcpy(w.buf[l], &v, sizeof(T));
  w.bar.arrive_and_wait();
  T r; memcpy(&r, w.buf[src_lane & 31], sizeof(T));
  w.bar.arrive_and_wait();
  return r;
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) { return shfl_emu(v, (int)(threadIdx.x & 31) ^ m); }
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return shfl_emu(v, src); }
template <class T> inline T __shfl_up_sync(unsigned, T v, int d) { const int l = threadIdx.x & 31; return shfl_emu(v, l >= d ? l - d : l); }
inline void __syncwarp(unsigned = 0xffffffffu) { g_warp[threadIdx.x >> 5].bar.arrive_and_wait(); }
inline int __any_sync(unsigned, int pred) {
  WarpX& w = g_warp[threadIdx.x >> 5];
  const int l = threadIdx.x & 31;
  memcpy(w.buf[l], &pred, sizeof(int));
  w.bar.arrive_and_wait();
  int any = 0;
  for (int k = 0; k < 32; ++k) { int p; memcpy(&p, w.buf[k], sizeof(int)); any |= (p != 0); }
  w.bar.arrive_and_wait();
  return any;
}
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
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
  std::barrier<> bar(256);
  g_cta_barrier = &bar;
  std::vector<std::thread> th;
  for (int t = 0; t < 256; ++t)
    th.emplace_back([&, t] {
      for (int y = 0; y < ny; ++y)
        for (int b = 0; b < nx; ++b) {
          threadIdx = {(unsigned)t, 0, 0};
          blockIdx = {(unsigned)b, (unsigned)y, 0};
          body();
          bar.arrive_and_wait();
        }
    });
  for (auto& x : th) x.join();
}

// sc: R_d, cp_d, cv_d, T_0, p_ref_theta, T_surf_ref, T_min_ref, T_min_sgs, dt, ν₄ᵥ, ν₄ₛ, divergence damping factor, hyperdiff, rayleigh, viscous,
//     energy upwinding, ncf, tracer upwinding ; vl: [14][64] = sc2i, sf2i, sf, dzc, dzf, mc, rmc, g33f, phic, dphif, brw, bruh, bvc, bvf ; D [16], w [4]
// which: 0 = k5_exp_a (writes Ytc, Ytf, H), 1 = k7_exp_c (kernels_lvl.cuh: reads H, updates Ytc, Ytf; one thread per (element, level), 4 elements per CTA),
//        2 / 3 = k5_tracer_a / k5_tracer_c
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
  // 4 = k_moist_c: water part of the hyperdiffusion apply (reads H[4] and Hw; updates ρe_tot of Ytc and ρ, ρq_tot of Ylc, or of Ytc when Ylc is NULL)
  else if (which == 4) run_grid(nh, 1, [&] { k_moist_c<FT>(P, hgeo, &V, Yc, H, (const FT*)g_moist_Hw, Ytc, Ylc ? Ylc : Ytc); });
  else if (which == 0) run_grid(nh, 1, [&] { k5_exp_a<FT, 0>(P, hgeo, &V, Yc, Yf, Ytc, Ytf, H); });
  else if (which == 1) run_grid((nh + LVL_EPB - 1) / LVL_EPB, 3, [&] { k7_exp_c<FT, 0>(P, hgeo, &V, Yc, H, Ytc, Ytf); });
  // passive tracers (grid = elements × tracers): 2 = k5_tracer_a (Yₜ, Yₜ_lim, ∇²χ → H), 3 = k5_tracer_c (tracer hyperdiffusion → Yₜ_lim)
  else if (which == 2) { if (ntr - g_moist_on > 0) run_grid(nh, ntr - g_moist_on, [&] { k5_tracer_a<FT>(P, hgeo, &V, Yc, Yf, Ytc, Ylc, H); }); }
  else if (ntr - g_moist_on > 0) run_grid(nh, ntr - g_moist_on, [&] { k5_tracer_c<FT>(P, hgeo, &V, Yc, H, Ylc); });
  return 0;
}
