// kernels_dss.cuh — weighted direct stiffness summation and fused stage increments.
//
//   k_dss     Spaces.weighted_dss! (src/prognostic_equations/constrain_state.jl:59-64,
//             remaining_tendency.jl:18-21; docs/src/discretization.md:139-154): ONE gather–scatter
//             launch for a whole list of fields.  For every unique perimeter node (CSR built from
//             ClimaCore's Topology2D tables, bit-exact) and level: weight × value of each
//             collocated element node is summed in a fixed order (ascending global element id,
//             so results do not depend on the rank count) and written back to every local member.
//             Covariant12 pairs are summed in the local (east,north) basis using the per-node
//             ∂x/∂ξ and ∂ξ/∂x matrices copied from the grid (pole-safe: no analytic basis).
//   k_pack    copy whole element slabs of the send elements into the halo send buffer.
//   k_axpy_n  U = u + Σ_j c_j T_j  (ClimaTimeSteppers fused_increment!).
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int DSS_MAX_ITEMS = 8;
struct DssItem {
  void* p0;   // first (or only) component plane, local elements
  void* p1;   // second component of a Covariant12 pair, or nullptr
  void* g0;   // same planes for ghost elements (element index - nh), or nullptr
  void* g1;
  int nlev;     // levels per node
  int estride;  // element stride of this plane in values (local field)
  int gstride;  // element stride in the ghost buffer
};
struct DssArgs {
  DssItem it[DSS_MAX_ITEMS];
  int n;
};

template <class FT>
__global__ void __launch_bounds__(256) k_dss(DssArgs A, const int* __restrict__ off, const int* __restrict__ mem,
                                            const FT* __restrict__ hgeo, int nnodes, int nh) {
  const int v = threadIdx.x;
  const int node = blockIdx.x * 4 + threadIdx.y;
  if (node >= nnodes) return;
  const int b = off[node];
  const int cnt = off[node + 1] - b;
  int el[4], nd[4];
  FT w[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (q < cnt) {
      int m = mem[b + q];
      el[q] = m >> 4; nd[q] = m & 15;
      w[q] = hgeo[((size_t)el[q] * HG_N + HG_DSSW) * 16 + nd[q]];
    }
  }
  for (int k = 0; k < A.n; ++k) {
    const DssItem I = A.it[k];
    if (v >= I.nlev) continue;
    FT* p0 = reinterpret_cast<FT*>(I.p0);
    FT* p1 = reinterpret_cast<FT*>(I.p1);
    const FT* g0 = reinterpret_cast<const FT*>(I.g0);
    const FT* g1 = reinterpret_cast<const FT*>(I.g1);
    if (!p1) {
      FT s = FT(0);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (q < cnt) {
          FT a = el[q] < nh ? p0[(size_t)el[q] * I.estride + nd[q] * I.nlev + v]
                            : g0[(size_t)(el[q] - nh) * I.gstride + nd[q] * I.nlev + v];
          s += w[q] * a;
        }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (q < cnt && el[q] < nh) p0[(size_t)el[q] * I.estride + nd[q] * I.nlev + v] = s;
    } else {
      FT su = FT(0), sv = FT(0);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (q < cnt) {
          FT a1, a2;
          if (el[q] < nh) {
            size_t o = (size_t)el[q] * I.estride + nd[q] * I.nlev + v;
            a1 = p0[o]; a2 = p1[o];
          } else {
            size_t o = (size_t)(el[q] - nh) * I.gstride + nd[q] * I.nlev + v;
            a1 = g0[o]; a2 = g1[o];
          }
          const FT* hg = hgeo + (size_t)el[q] * HG_N * 16 + nd[q];
          FT uu = hg[HG_AI00 * 16] * a1 + hg[HG_AI10 * 16] * a2;
          FT vv = hg[HG_AI01 * 16] * a1 + hg[HG_AI11 * 16] * a2;
          su += w[q] * uu; sv += w[q] * vv;
        }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (q < cnt && el[q] < nh) {
          const FT* hg = hgeo + (size_t)el[q] * HG_N * 16 + nd[q];
          size_t o = (size_t)el[q] * I.estride + nd[q] * I.nlev + v;
          p0[o] = hg[HG_A00 * 16] * su + hg[HG_A10 * 16] * sv;
          p1[o] = hg[HG_A01 * 16] * su + hg[HG_A11 * 16] * sv;
        }
    }
  }
}

// copy element slabs (all components of one field) of the listed elements into a packed buffer
template <class FT>
__global__ void k_pack(const FT* __restrict__ src, FT* __restrict__ dst, const int* __restrict__ elems, int slab) {
  const int e = elems[blockIdx.x];
  const FT* s = src + (size_t)e * slab;
  FT* d = dst + (size_t)blockIdx.x * slab;
  for (int i = threadIdx.x; i < slab; i += blockDim.x) d[i] = s[i];
}

constexpr int AXPY_MAX = 8;
template <class FT>
struct AxpyArgs {
  const FT* T[AXPY_MAX];
  FT c[AXPY_MAX];
  int n;
};

// `nlev` > 0 marks a face field with nlev levels per column whose first and last level are forced to
// zero (the u₃ impenetrability filter of cache_imp!, folded into the increment by the native stepper).
template <class FT, int VEC>
__global__ void __launch_bounds__(256) k_axpy_n(FT* out, const FT* base, AxpyArgs<FT> A, size_t nvec, int nlev) {
  struct alignas(sizeof(FT) * VEC) Vt { FT x[VEC]; };
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    Vt r = reinterpret_cast<const Vt*>(base)[i];
    for (int k = 0; k < A.n; ++k) {
      Vt t = reinterpret_cast<const Vt*>(A.T[k])[i];
#pragma unroll
      for (int q = 0; q < VEC; ++q) r.x[q] += A.c[k] * t.x[q];
    }
    if (nlev > 0) {
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        int lev = (int)((i * VEC + q) % (size_t)nlev);
        if (lev == 0 || lev == nlev - 1) r.x[q] = FT(0);
      }
    }
    reinterpret_cast<Vt*>(out)[i] = r;
  }
}

// out = (a - b) * s   (T_imp[i] = (U - temp)/dtγ)
template <class FT, int VEC>
__global__ void __launch_bounds__(256) k_diff_scale(FT* out, const FT* a, const FT* b, FT s, size_t nvec) {
  struct alignas(sizeof(FT) * VEC) Vt { FT x[VEC]; };
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    Vt x = reinterpret_cast<const Vt*>(a)[i], y = reinterpret_cast<const Vt*>(b)[i], r;
#pragma unroll
    for (int q = 0; q < VEC; ++q) r.x[q] = (x.x[q] - y.x[q]) / s;
    reinterpret_cast<Vt*>(out)[i] = r;
  }
}

}  // namespace b200
