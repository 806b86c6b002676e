// kernels_limiter.cuh — lim!(Y, p, t, ref_Y): the SEM quasi-monotone tracer limiter
// (src/prognostic_equations/limited_tendencies.jl:64-122 → ClimaCore Limiters.QuasiMonotoneLimiter:
//  compute_bounds! = compute_element_bounds! + compute_neighbor_bounds_local!, apply_limiter! = apply_limit_slab! per
//  (element, level) slab [UPSTREAM-RECALL ClimaCore 0.15.1 src/Limiters/quasimonotone.jl]).
//
// One element per CTA, ONE THREAD PER LEVEL (level is the fastest index: coalesced), the 16 nodes of the slab in registers.
//   k_lim_bounds  per element and level: min / max of q = ρq/ρ over the 16 nodes of the REFERENCE state
//   k_lim_apply   bounds widened over the element's vertex neighbours; clip ρq to [ρ q_min, ρ q_max] (bounds relaxed to contain
//                 the slab mean) and redistribute the clipped mass over the nodes with room in proportion to ρ·WJ; at most 16
//                 iterations, stop when |Δmass| ≤ eps·|mass|.  Sums run over the nodes in the order n = 4j + i, like the oracle.
// grid = (elements, tracers).  The vertical Jacobian factor of WJ is constant over a slab and cancels, so the horizontal W·J2 is used.
#pragma once
#include "common.cuh"
#include "moist.cuh"  // eps_

namespace b200 {


// E (multi-rank only): the bounds again, in the shape of a centre field [h][2·ntr][16][nv] (every node column of the element holds
// the element's bound), so the ordinary peer-memory halo (k_pack_p2p: the node columns shared with each neighbour) carries them.
template <class FT>
__global__ void __launch_bounds__(64) k_lim_bounds(const FT* __restrict__ ref_c, int ncf, int nv, int nh, FT* __restrict__ bnd,
                                                   FT* __restrict__ E, int ntr) {
  const int e = blockIdx.x, t = blockIdx.y, v = threadIdx.x;
  if (v >= nv) return;
  const FT* r = ref_c + (size_t)e * ncf * 16 * nv + v;
  const FT* x = r + (size_t)(4 + t) * 16 * nv;
  FT lo = x[0] / r[0], hi = lo;
#pragma unroll
  for (int n = 1; n < 16; ++n) {
    const FT q = x[n * nv] / r[n * nv];
    lo = fmin_(lo, q); hi = fmax_(hi, q);
  }
  FT* b = bnd + ((size_t)(t * nh + e) * 2) * LV;
  b[v] = lo; b[LV + v] = hi;
  if (E) {
    FT* d = E + ((size_t)e * 2 * ntr + 2 * t) * 16 * nv + v;
#pragma unroll
    for (int n = 0; n < 16; ++n) { d[n * nv] = lo; d[(16 + n) * nv] = hi; }
  }
}

template <class FT>
__global__ void __launch_bounds__(64) k_lim_apply(FT* __restrict__ Yc, int ncf, int nv, int nh, const FT* __restrict__ bnd,
                                                  const int* __restrict__ nbr_off, const int* __restrict__ nbr_list,
                                                  const FT* __restrict__ hgeo, const FT* __restrict__ ghostE, long long gpar,
                                                  const int* __restrict__ seq, const int* __restrict__ ghost_node, int ntr) {
  const int e = blockIdx.x, t = blockIdx.y, v = threadIdx.x;
  if (v >= nv) return;
  const FT* bt = bnd + (size_t)t * nh * 2 * LV;
  FT qmin = bt[(size_t)e * 2 * LV + v], qmax = bt[(size_t)e * 2 * LV + LV + v];
  const FT* gE = ghostE ? reinterpret_cast<const FT*>(reinterpret_cast<const char*>(ghostE) + (size_t)(*seq & 1) * (size_t)gpar) : nullptr;
  for (int k = nbr_off[e]; k < nbr_off[e + 1]; ++k) {
    const int nb = nbr_list[k];
    if (nb < nh) {
      qmin = fmin_(qmin, bt[(size_t)nb * 2 * LV + v]);
      qmax = fmax_(qmax, bt[(size_t)nb * 2 * LV + LV + v]);
    } else {  // ghost neighbour: its bounds arrived in the node column the owner shares with this rank
      const int g = nb - nh;
      const FT* p = gE + ((size_t)g * 2 * ntr + 2 * t) * 16 * nv + ghost_node[g] * nv + v;
      qmin = fmin_(qmin, p[0]);
      qmax = fmax_(qmax, p[(size_t)16 * nv]);
    }
  }
  const FT* r = Yc + (size_t)e * ncf * 16 * nv + v;
  FT* xg = Yc + (size_t)e * ncf * 16 * nv + (size_t)(4 + t) * 16 * nv + v;
  FT rho[16], x[16], w[16];
  FT total_mass = FT(0), tracer_mass = FT(0);
#pragma unroll
  for (int n = 0; n < 16; ++n) {
    rho[n] = r[n * nv]; x[n] = xg[n * nv]; w[n] = hgeo[((size_t)e * HG_N + HG_WJ) * 16 + n];
    total_mass += rho[n] * w[n];
    tracer_mass += x[n] * w[n];
  }
  const FT q_avg = tracer_mass / total_mass;
  qmin = fmin_(qmin, q_avg); qmax = fmax_(qmax, q_avg);
  const FT rtol = eps_<FT>();
  for (int it = 0; it < 16; ++it) {
    FT dm = FT(0);
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const FT xmax = rho[n] * qmax, xmin = rho[n] * qmin;
      if (x[n] > xmax) { dm += (x[n] - xmax) * w[n]; x[n] = xmax; }
      else if (x[n] < xmin) { dm += (x[n] - xmin) * w[n]; x[n] = xmin; }
    }
    if (abs_(dm) <= rtol * abs_(tracer_mass)) break;
    const bool add = dm > FT(0);
    FT mass_at = FT(0);
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const bool room = add ? (x[n] < rho[n] * qmax) : (x[n] > rho[n] * qmin);
      if (room) mass_at += rho[n] * w[n];
    }
    const FT dq = dm / mass_at;
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const bool room = add ? (x[n] < rho[n] * qmax) : (x[n] > rho[n] * qmin);
      if (room) x[n] += rho[n] * dq;
    }
  }
#pragma unroll
  for (int n = 0; n < 16; ++n) xg[n * nv] = x[n];
}

}  // namespace b200
