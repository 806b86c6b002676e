// CPU emulation of the weighted-DSS kernel of the benchmarked step, k_dss2 (kernels_dss.cuh; Float64 instantiation, single-rank path),
// from its unchanged source, with the node records built like capi.cu:create_geo builds them.  Test infrastructure only.
#include <vector>
#define b200 b200_emud
#include "cuda_runtime.h"
thread_local uint3_emu threadIdx, blockIdx;
// static __shared__ arrays of these kernels: one instance per kernel instantiation, shared by the 256 fibers of the emulated CTA
#undef __shared__
#define __shared__ static
struct dim3_emu { unsigned x, y, z; };
static dim3_emu gridDim{1, 1, 1}, blockDim{64, 4, 1};
inline void __threadfence_system() {}
inline void __threadfence() {}
inline void __syncwarp(unsigned = 0xffffffffu) {}
inline long long clock64() { return 0; }
inline void __nanosleep(unsigned) {}
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline void __trap() { __builtin_trap(); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
__asm__(".macro griddepcontrol.launch_dependents\n.endm\n.macro griddepcontrol.wait\n.endm");
#include "kernels_dss.cuh"

using namespace b200;
typedef double FT;
static uint3_emu tid64x4(int t) { return uint3_emu{(unsigned)(t & 63), (unsigned)(t >> 6), 0}; }  // blockDim = (64, 4)

// state DSS (ρ, (uₕ₁,uₕ₂) as a Covariant12 pair, ρe_tot, u₃): off/mem = CSR of the unique perimeter nodes (mem = elem·16 + node),
// hgeo [nh][HG_N][16] with HG_DSSW, HG_A**, HG_AI** filled
extern "C" __attribute__((visibility("default"))) int emu_dss_state(int nh, int nv, int nnodes, const int* off, const int* mem,
                                                                    const double* hgeo, double* Yc, double* Yf) {
  std::vector<DssNode<FT>> rec((size_t)nnodes);
  memset(rec.data(), 0, rec.size() * sizeof(DssNode<FT>));
  for (int nd = 0; nd < nnodes; ++nd) {
    DssNode<FT>& R = rec[nd];
    R.cnt = off[nd + 1] - off[nd];
    for (int q = 0; q < R.cnt; ++q) {
      const int m = mem[off[nd] + q];
      const FT* o = hgeo + (size_t)(m >> 4) * HG_N * 16 + (m & 15);
      R.mem[q] = m;
      R.w[q] = o[HG_DSSW * 16];
      R.ai[q][0] = o[HG_AI00 * 16]; R.ai[q][1] = o[HG_AI10 * 16]; R.ai[q][2] = o[HG_AI01 * 16]; R.ai[q][3] = o[HG_AI11 * 16];
      R.a[q][0] = o[HG_A00 * 16]; R.a[q][1] = o[HG_A10 * 16]; R.a[q][2] = o[HG_A01 * 16]; R.a[q][3] = o[HG_A11 * 16];
    }
  }
  DssArgs A;
  memset(&A, 0, sizeof(A));
  const int cs = 16 * nv;
  A.n = 4;
  A.it[0] = DssItem{Yc, nullptr, nullptr, nullptr, nv, 4 * cs, 0};
  A.it[1] = DssItem{Yc + cs, Yc + 2 * cs, nullptr, nullptr, nv, 4 * cs, 0};
  A.it[2] = DssItem{Yc + 3 * cs, nullptr, nullptr, nullptr, nv, 4 * cs, 0};
  A.it[3] = DssItem{Yf, nullptr, nullptr, nullptr, nv + 1, 16 * (nv + 1), 0};
  P2PWait W{nullptr, nullptr, nullptr, 0};
  const int nblocks = (nnodes + 3) / 4;
  const std::function<void()> fn = [&] { k_dss2<FT, 4, 0x2, false, false>(A, rec.data(), 0, nnodes, nh, W); };
  for (int b = 0; b < nblocks; ++b) {
    blockIdx = {(unsigned)b, 0, 0};
    emu::run_cta(256, tid64x4, fn);
  }
  return 0;
}

static std::vector<DssNode<FT>> build_records(int nnodes, const int* off, const int* mem, const double* hgeo) {
  std::vector<DssNode<FT>> rec((size_t)nnodes);
  memset(rec.data(), 0, rec.size() * sizeof(DssNode<FT>));
  for (int nd = 0; nd < nnodes; ++nd) {
    DssNode<FT>& R = rec[nd];
    R.cnt = off[nd + 1] - off[nd];
    for (int q = 0; q < R.cnt; ++q) {
      const int m = mem[off[nd] + q];
      const FT* o = hgeo + (size_t)(m >> 4) * HG_N * 16 + (m & 15);
      R.mem[q] = m;
      R.w[q] = o[HG_DSSW * 16];
      R.ai[q][0] = o[HG_AI00 * 16]; R.ai[q][1] = o[HG_AI10 * 16]; R.ai[q][2] = o[HG_AI01 * 16]; R.ai[q][3] = o[HG_AI11 * 16];
      R.a[q][0] = o[HG_A00 * 16]; R.a[q][1] = o[HG_A10 * 16]; R.a[q][2] = o[HG_A01 * 16]; R.a[q][3] = o[HG_A11 * 16];
    }
  }
  return rec;
}

// k_axpy_dss<FT, 3>: out = dss!(filter(base + Σ_{k<3} c_k T_k)) — the stage increment fused with the state DSS (single rank): node blocks
// interleaved with one interior block per element.  T: three (Tc, Tf) pairs; dmask as in AxDssArgs.
extern "C" __attribute__((visibility("default"))) int emu_axpy_dss3(int nh, int nv, int ncf, int nnodes, const int* off, const int* mem,
                                                                    const double* hgeo, const double* base_c, const double* base_f,
                                                                    const double* T0c, const double* T0f, const double* T1c, const double* T1f,
                                                                    const double* T2c, const double* T2f, const double* coef, unsigned dmask,
                                                                    double* out_c, double* out_f) {
  std::vector<DssNode<FT>> rec = build_records(nnodes, off, mem, hgeo);
  AxDssArgs<FT> A;
  memset(&A, 0, sizeof(A));
  A.out_c = out_c; A.out_f = out_f; A.base_c = base_c; A.base_f = base_f;
  A.Tc[0] = T0c; A.Tf[0] = T0f; A.Tc[1] = T1c; A.Tf[1] = T1f; A.Tc[2] = T2c; A.Tf[2] = T2f;
  for (int k = 0; k < 3; ++k) A.c[k] = coef[k];
  A.dmask = dmask; A.ncf = ncf; A.nv = nv; A.nh = nh;
  P2PWait W{nullptr, nullptr, nullptr, 0};
  const int nbn = (nnodes + 3) / 4, nint = nh, nblocks = nbn + nint;
  const std::function<void()> fn = [&] { k_axpy_dss<FT, 3, false, false>(A, rec.data(), 0, nnodes, nbn, nint, W); };
  for (int b = 0; b < nblocks; ++b) {
    blockIdx = {(unsigned)b, 0, 0};
    emu::run_cta(256, tid64x4, fn);
  }
  return 0;
}

// DSS of the ∇² fields inside T_exp (remaining_tendency.jl:18-21): (∇²u₁, ∇²u₂) as a Covariant12 pair, ∇²u₃ and ∇²s_d as scalars —
// k_dss2<FT, 3, 0x1> on H [nh][ncf][16][nv]
extern "C" __attribute__((visibility("default"))) int emu_dss_h(int nh, int nv, int ncf, int nnodes, const int* off, const int* mem,
                                                                const double* hgeo, double* H) {
  std::vector<DssNode<FT>> rec = build_records(nnodes, off, mem, hgeo);
  DssArgs A;
  memset(&A, 0, sizeof(A));
  const int cs = 16 * nv;
  A.n = 3;
  A.it[0] = DssItem{H, H + cs, nullptr, nullptr, nv, ncf * cs, 0};
  A.it[1] = DssItem{H + 2 * cs, nullptr, nullptr, nullptr, nv, ncf * cs, 0};
  A.it[2] = DssItem{H + 3 * cs, nullptr, nullptr, nullptr, nv, ncf * cs, 0};
  P2PWait W{nullptr, nullptr, nullptr, 0};
  const int nblocks = (nnodes + 3) / 4;
  const std::function<void()> fn = [&] { k_dss2<FT, 3, 0x1, false, false>(A, rec.data(), 0, nnodes, nh, W); };
  for (int b = 0; b < nblocks; ++b) {
    blockIdx = {(unsigned)b, 0, 0};
    emu::run_cta(256, tid64x4, fn);
  }
  return 0;
}

// k_axpy_dss with n = 1..6 terms (pointer arrays Tc[n], Tf[n]); otherwise as emu_axpy_dss3
template <int N>
static void run_axdss(const AxDssArgs<FT>& A, const std::vector<DssNode<FT>>& rec, int nnodes, int nh) {
  P2PWait W{nullptr, nullptr, nullptr, 0};
  const int nbn = (nnodes + 3) / 4, nint = nh, nblocks = nbn + nint;
  const std::function<void()> fn = [&] { k_axpy_dss<FT, N, false, false>(A, rec.data(), 0, nnodes, nbn, nint, W); };
  for (int b = 0; b < nblocks; ++b) {
    blockIdx = {(unsigned)b, 0, 0};
    emu::run_cta(256, tid64x4, fn);
  }
}
extern "C" __attribute__((visibility("default"))) int emu_axpy_dss_n(int n, int nh, int nv, int ncf, int nnodes, const int* off, const int* mem,
                                                                     const double* hgeo, const double* base_c, const double* base_f,
                                                                     const double* const* Tc, const double* const* Tf, const double* coef,
                                                                     unsigned dmask, double* out_c, double* out_f) {
  std::vector<DssNode<FT>> rec = build_records(nnodes, off, mem, hgeo);
  AxDssArgs<FT> A;
  memset(&A, 0, sizeof(A));
  A.out_c = out_c; A.out_f = out_f; A.base_c = base_c; A.base_f = base_f;
  for (int k = 0; k < n; ++k) { A.Tc[k] = Tc[k]; A.Tf[k] = Tf[k]; A.c[k] = coef[k]; }
  A.dmask = dmask; A.ncf = ncf; A.nv = nv; A.nh = nh;
  switch (n) {
    case 1: run_axdss<1>(A, rec, nnodes, nh); break;
    case 2: run_axdss<2>(A, rec, nnodes, nh); break;
    case 3: run_axdss<3>(A, rec, nnodes, nh); break;
    case 4: run_axdss<4>(A, rec, nnodes, nh); break;
    case 5: run_axdss<5>(A, rec, nnodes, nh); break;
    case 6: run_axdss<6>(A, rec, nnodes, nh); break;
    default: return -1;
  }
  return 0;
}

// The generic k_dss (one item per blockIdx.y, CSR + hgeo instead of node records): the fallback of impl_dss for states with more than two
// tracers.  State items: ρ, (uₕ₁,uₕ₂), ρe_tot, ρχ₁…, u₃.  No barriers: the threads are run one after the other.
extern "C" __attribute__((visibility("default"))) int emu_dss_generic(int nh, int nv, int ncf, int nnodes, const int* off, const int* mem,
                                                                      const double* hgeo, double* Yc, double* Yf) {
  DssArgs A;
  memset(&A, 0, sizeof(A));
  const int cs = 16 * nv, es = ncf * cs;
  int n = 0;
  A.it[n++] = DssItem{Yc, nullptr, nullptr, nullptr, nv, es, 0};
  A.it[n++] = DssItem{Yc + cs, Yc + 2 * cs, nullptr, nullptr, nv, es, 0};
  for (int q = 3; q < ncf; ++q) A.it[n++] = DssItem{Yc + q * cs, nullptr, nullptr, nullptr, nv, es, 0};
  A.it[n++] = DssItem{Yf, nullptr, nullptr, nullptr, nv + 1, 16 * (nv + 1), 0};
  A.n = n;
  if (n > DSS_MAX_ITEMS) return -1;
  const int nblocks = (nnodes + 3) / 4;
  for (int k = 0; k < n; ++k)
    for (int b = 0; b < nblocks; ++b)
      for (int t = 0; t < 256; ++t) {
        threadIdx = {(unsigned)(t & 63), (unsigned)(t >> 6), 0};
        blockIdx = {(unsigned)b, (unsigned)k, 0};
        k_dss<FT>(A, off, mem, hgeo, nnodes, nh);
      }
  return 0;
}
