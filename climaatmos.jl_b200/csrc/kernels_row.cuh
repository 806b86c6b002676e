// kernels_row.cuh — the "one GLL row per thread" layout shared by the explicit-tendency kernels (kernels_pair.cuh).
//
// One element per CTA of 256 threads.  Thread (v, j) owns the four nodes i = 0..3 of row j at level v in
// registers; a warp holds 8 consecutive levels × 4 rows (lane = vl + 8·j).  On B200 this gives
//   * ξ¹-derivatives: thread-local 4×4 contractions with D/Dw entries from the constant bank;
//   * ξ²-derivatives: warp shuffles between the lanes vl, vl+8, vl+16, vl+24 of a level with per-thread matrix
//     rows/columns — no shared-memory slabs and no block barriers for horizontal work;
//   * vertical neighbours (k±1): a shared-memory column exchange, 3 block barriers per launch;
//   * global accesses: 32-byte segments of 8 consecutive levels (full sector efficiency).
// History (profiles/r1_ncu_summary.md): whole-slab shared-memory kernels (500 LDS/point, LSU/latency-bound) and a
// one-thread-per-level variant (16 nodes/thread, 255 registers, instruction-fetch-bound) preceded this layout; their
// code was removed in round 2 once the A/B record was committed under profiles/.
#pragma once
#include "common.cuh"

namespace b200 {

// derivative matrices in the constant bank: D[16] then Dw[16] (strong / weak), filled by b200_create
__constant__ float c_Df[32];   // D[16] then Dw[16]
__constant__ double c_Dd[32];
template <class FT> __device__ __forceinline__ FT cM(int k);
template <> __device__ __forceinline__ float cM<float>(int k) { return c_Df[k]; }
template <> __device__ __forceinline__ double cM<double>(int k) { return c_Dd[k]; }

constexpr unsigned FULLM = 0xffffffffu;

template <class FT>
struct Lev {  // per-thread level constants
  FT sc, mc, rmc, phi, g33lo, g33hi, bruh, bvc;   // centre v
  FT sf, sf2i, dzf, mclo, sclo, bvf;              // face v (mclo/sclo: centre v-1)
};
template <class FT>
__device__ __forceinline__ Lev<FT> load_lev(const VLev<FT>* __restrict__ V, int v, int nv) {
  Lev<FT> L;
  const int vc = v < nv ? v : nv - 1, vm = v > 0 ? v - 1 : 0, vf = v <= nv ? v : nv, vf1 = v + 1 <= nv ? v + 1 : nv;
  L.sc = V->sc2i[vc]; L.mc = V->mc[vc]; L.rmc = V->rmc[vc]; L.phi = V->phic[vc]; L.g33lo = V->g33f[vf]; L.g33hi = V->g33f[vf1];
  L.bruh = V->bruh[vc]; L.bvc = V->bvc[vc];
  L.sf = V->sf[vf]; L.sf2i = V->sf2i[vf]; L.dzf = V->dzf[vf]; L.mclo = V->mc[vm < nv ? vm : nv - 1];
  L.sclo = V->sc2i[vm < nv ? vm : nv - 1]; L.bvf = V->bvf[vf];
  return L;
}

constexpr int CT = 256;



#define B200_ROW_PROLOGUE                                                                             \
  const int e = blockIdx.x, lane = threadIdx.x & 31, vl = lane & 7, j = lane >> 3;                    \
  const int v = (threadIdx.x >> 5) * 8 + vl, nv = P.nv, nf = nv + 1;                                  \
  const bool cv = v < nv, fv = v < nf;                                                                \
  for (int k = threadIdx.x; k < HG_ELEM * 16; k += CT) hg[k] = hgeo[(size_t)e * HG_N * 16 + k];       \
  FT md[4], mw[4];                                                                                    \
  _Pragma("unroll") for (int k = 0; k < 4; ++k) { md[k] = cM<FT>(j * 4 + k); mw[k] = cM<FT>(16 + j * 4 + k); } \
  const Lev<FT> L = load_lev(vlev, v, nv);                                                            \
  const int n0 = j * 4;

#define B200_ROW_PROLOGUE_NV(NVC_)                                                                            \
  const int e = blockIdx.x, lane = threadIdx.x & 31, vl = lane & 7, j = lane >> 3;                    \
  const int v = (threadIdx.x >> 5) * 8 + vl, nv = (NVC_) ? (NVC_) : P.nv, nf = nv + 1;                                  \
  const bool cv = v < nv, fv = v < nf;                                                                \
  for (int k = threadIdx.x; k < HG_ELEM * 16; k += CT) hg[k] = hgeo[(size_t)e * HG_N * 16 + k];       \
  FT md[4], mw[4];                                                                                    \
  _Pragma("unroll") for (int k = 0; k < 4; ++k) { md[k] = cM<FT>(j * 4 + k); mw[k] = cM<FT>(16 + j * 4 + k); } \
  const Lev<FT> L = load_lev(vlev, v, nv);                                                            \
  const int n0 = j * 4;

}  // namespace b200
