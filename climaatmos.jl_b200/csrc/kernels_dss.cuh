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
  // peer-memory halo: the ghost planes g0/g1 point into parity block 0; the block in use is (*seq & 1), gpar bytes further on
  const int* seq;
  long long gpar;
};

// wait (inside a consumer kernel) until every neighbour has raised its flag to the current exchange number: thread 0 of the
// block polls the flags in my own memory, the block then proceeds
// A rank that waits longer than g_p2p_spin SM clocks (≈ 30 s by default) for a neighbour neither hangs the GPU nor traps (a trap
// would destroy the CUDA context): it records (neighbour rank + 1) | exchange number << 8 in the context's host-mapped error word and
// goes on with whatever the ghost block holds; every C-ABI entry point that uses the halo checks that word first and returns an
// error code with a message (capi.cu: halo_failed) — the state is invalid from then on, the process is not.
__device__ int* g_p2p_err = nullptr;               // host-mapped error word (set by b200_halo_import), 0 = ok
__device__ long long g_p2p_spin = 60000000000ll;   // B200_P2P_SPIN_LIMIT overrides (tests)
__device__ __forceinline__ bool p2p_timed_out(long long t0, int nbr, int value) {
  if (clock64() - t0 <= g_p2p_spin) return false;
  if (g_p2p_err) { *reinterpret_cast<volatile int*>(g_p2p_err) = (nbr + 1) | (value << 8); __threadfence_system(); }
  return true;
}
struct P2PWait { const int* flags; const int* nbr_rank; const int* seq; int nn; };
__device__ __forceinline__ void p2p_block_wait(const P2PWait& W) {
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    const int value = *reinterpret_cast<const volatile int*>(W.seq);
    const long long t0 = clock64();
    for (int q = 0; q < W.nn; ++q) {
      const volatile int* f = W.flags + W.nbr_rank[q];
      while (*f < value) {
        __nanosleep(40);
        if (p2p_timed_out(t0, W.nbr_rank[q], value)) break;
      }
    }
    __threadfence_system();
  }
  __syncthreads();
}

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
  // one item per blockIdx.y: every (node, level, item) is an independent thread, so the member loads of all
  // items are in flight together (the first version looped over items per thread and was latency-bound at
  // ≈1.7 TB/s, profiles/r1_ncu_summary.md)
  {
    const int k = blockIdx.y;
    const DssItem I = A.it[k];
    if (v >= I.nlev) return;
    FT* p0 = reinterpret_cast<FT*>(I.p0);
    FT* p1 = reinterpret_cast<FT*>(I.p1);
    const size_t gp = A.seq ? (size_t)(*A.seq & 1) * (size_t)A.gpar : 0;  // parity block of the peer-memory halo
    const FT* g0 = reinterpret_cast<const FT*>(reinterpret_cast<const char*>(I.g0) + gp);
    const FT* g1 = reinterpret_cast<const FT*>(reinterpret_cast<const char*>(I.g1) + gp);
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

// ---- peer-memory halo (NVLink P2P): the sender writes its boundary-element slabs straight into the
// neighbour's ghost buffer (cudaIpc-mapped pointer) and then raises a flag there; no NCCL call, no staging copy.
struct P2PField { const void* src; int slab; long long goff; int ncomp, nlev; };  // goff: offset of this field's ghost block per unit nh_ghost; slab = ncomp·16·nlev
struct P2PArgs {
  P2PField f[4];
  int nfields;
};
// The exchange number lives in DEVICE memory (*seq, incremented by k_p2p_signal), so none of these kernels takes a per-call value:
// a captured CUDA graph of the step replays correctly.  Ghost blocks are double-buffered by the parity of the exchange number
// (a block is rewritten two exchanges later, after the neighbour's next flag has proved it finished reading).
// Signalling is part of the pack kernels: every block bumps a device counter when its slab is written; the last one to finish
// raises my flag in every neighbour's memory and advances the exchange number (P2PSig).  k_p2p_signal remains for ranks with
// nothing to send.
struct P2PSig {
  int* const* peer_flags;  // [nn] address of my flag in neighbour q
  int* seq;                // exchange number (device memory)
  int* done;               // blocks finished in the current pack launch
  int nn;
};
__device__ __forceinline__ void p2p_block_done(const P2PSig& S, int nblocks, int value) {
  __shared__ int s_last;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  __threadfence_system();  // this thread's peer writes are visible system-wide before the block reports
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(S.done, 1) == nblocks - 1);
  __syncthreads();
  if (s_last) {
    if (tid < S.nn) {
      __threadfence_system();
      *reinterpret_cast<volatile int*>(S.peer_flags[tid]) = value;
    }
    __syncthreads();
    if (tid == 0) { *S.done = 0; *S.seq = value; }
  }
}
// Only the node columns the neighbour actually sums are sent: slot_mask[slot] has bit n set when node n of the send element is
// collocated with a node of an element owned by that neighbour (4 of 16 columns across an edge, 1 across a vertex).
__device__ __forceinline__ int p2p_nodes(unsigned mask, int* nodes /* shared, 16 */) {
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    int c = 0;
    for (int n = 0; n < 16; ++n) if (mask >> n & 1) nodes[c++] = n;
  }
  __syncthreads();
  return __popc(mask);
}
// The send plan of a rank.  The pack work rides in the FIRST nblocks blocks of the kernel that sums the ghost-free nodes (k_dss2 /
// k_axpy_dss with PACK): the slabs fly while the rest of that grid works, and the packed (ghost-touching) node columns are
// disjoint from the columns the ghost-free nodes update in place.
struct P2PPlan {
  const int* send_elems; const int* slot_nbr; const int* slot_dst; const int* slot_mask;
  void* const* dst;          // [2][nn] neighbour ghost blocks per parity
  const int* nbr_nh_ghost;
  P2PSig S;
  int nblocks;               // send slots
};
// one block per (send slot): copy the shared node columns of the element into the neighbour's ghost block, then report
template <class FT>
__device__ __forceinline__ void p2p_pack_body(const P2PArgs& A, const P2PPlan& Q, int slot, int* nodes) {
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
  const int e = Q.send_elems[slot], q = Q.slot_nbr[slot], g = Q.slot_dst[slot];
  const int cnt = p2p_nodes((unsigned)Q.slot_mask[slot], nodes);
  const int value = *Q.S.seq + 1;
  FT* base = reinterpret_cast<FT*>(Q.dst[(value & 1) * Q.S.nn + q]);
  for (int k = 0; k < A.nfields; ++k) {
    const FT* s = reinterpret_cast<const FT*>(A.f[k].src) + (size_t)e * A.f[k].slab;
    FT* d = base + (size_t)A.f[k].goff * Q.nbr_nh_ghost[q] + (size_t)g * A.f[k].slab;
    const int nlev = A.f[k].nlev, tot = A.f[k].ncomp * cnt * nlev;
    for (int i = tid; i < tot; i += nt) {
      const int lev = i % nlev, t = i / nlev, o = ((t / cnt) * 16 + nodes[t % cnt]) * nlev + lev;
      d[o] = s[o];
    }
  }
  p2p_block_done(Q.S, Q.nblocks, value);
}
template <class FT>
__global__ void k_pack_p2p(P2PArgs A, P2PPlan Q) {
  __shared__ int nodes[16];
  pdl_wait();
  p2p_pack_body<FT>(A, Q, blockIdx.x, nodes);
}
// raise my flag in every neighbour's memory (after the pack kernel has completed) and advance the exchange number
__global__ void k_p2p_signal(int* const* __restrict__ peer_flags, int n, int* seq) {
  const int q = threadIdx.x;
  const int value = *seq + 1;
  if (q < n) {
    __threadfence_system();
    *reinterpret_cast<volatile int*>(peer_flags[q]) = value;
  }
  __syncwarp();
  if (q == 0) *seq = value;
}
// wait until every neighbour has raised its flag to the current exchange number in my memory
__global__ void k_p2p_wait(const int* __restrict__ flags, const int* __restrict__ nbr_rank, int n, const int* seq) {
  const int q = threadIdx.x;
  const int value = *seq;
  if (q < n) {
    const volatile int* f = flags + nbr_rank[q];
    const long long t0 = clock64();
    while (*f < value) {
      __nanosleep(50);
      if (p2p_timed_out(t0, nbr_rank[q], value)) break;
    }
    __threadfence_system();
  }
}

// Second-generation DSS: one compact record per unique node (members, weights and the 2×2 basis-change
// matrices of every member) so that a thread issues the record read and then ALL member loads of ALL items
// back to back — two dependent memory latencies instead of four per item (off → mem → weight → data).
template <class FT>
struct DssNode {
  int32_t cnt;
  int32_t mem[4];       // elem*16 + node
  FT w[4];              // DSS weight of the member
  FT ai[4][4];          // w·(∂ξ/∂x)ᵀ rows: (ai00, ai10, ai01, ai11) of the member (covariant → weighted physical)
  FT a[4][4];           // (a00, a10, a01, a11) of the member (physical → covariant)
};

// Body for one (node, level) with a compile-time member bound CNT (2 for face-interior nodes — 80 % of the
// nodes — 4 for vertices), compile-time item kinds (bit k of PAIRS ⇒ item k is a Covariant12 pair) and
// compile-time halo flag: the generic fully predicated version issued ≈800 instructions per thread
// (ncu: issue-active 77 %, i.e. instruction-bound, profiles/r1_ncu_summary.md).
template <class FT, int NI, int PAIRS, int CNT, bool HALO>
__device__ __forceinline__ void dss_body(const DssArgs& A, const DssNode<FT>& R, int cnt, int v, int nh) {
  FT x0[NI][CNT], x1[NI][CNT];
  int off[NI][CNT];
  const size_t gp = (HALO && A.seq) ? (size_t)(*A.seq & 1) * (size_t)A.gpar : 0;
#pragma unroll
  for (int k = 0; k < NI; ++k) {
    const DssItem& I = A.it[k];
#pragma unroll
    for (int q = 0; q < CNT; ++q) {
      x0[k][q] = FT(0); x1[k][q] = FT(0); off[k][q] = -1;
      if ((CNT == 2 || q < cnt) && v < I.nlev) {
        const int el = R.mem[q] >> 4, nd = R.mem[q] & 15;
        if (!HALO || el < nh) {
          const int o = el * I.estride + nd * I.nlev + v;
          off[k][q] = o;
          x0[k][q] = reinterpret_cast<const FT*>(I.p0)[o];
          if (PAIRS & (1 << k)) x1[k][q] = reinterpret_cast<const FT*>(I.p1)[o];
        } else {
          const int o = (el - nh) * I.gstride + nd * I.nlev + v;
          x0[k][q] = reinterpret_cast<const FT*>(reinterpret_cast<const char*>(I.g0) + gp)[o];
          if (PAIRS & (1 << k)) x1[k][q] = reinterpret_cast<const FT*>(reinterpret_cast<const char*>(I.g1) + gp)[o];
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NI; ++k) {
    const DssItem& I = A.it[k];
    FT* p0 = reinterpret_cast<FT*>(I.p0);
    FT* p1 = reinterpret_cast<FT*>(I.p1);
    if (!(PAIRS & (1 << k))) {
      FT s = FT(0);
#pragma unroll
      for (int q = 0; q < CNT; ++q)
        if (CNT == 2 || q < cnt) s += R.w[q] * x0[k][q];
#pragma unroll
      for (int q = 0; q < CNT; ++q)
        if (off[k][q] >= 0) p0[off[k][q]] = s;
    } else {
      FT su = FT(0), sv = FT(0);
#pragma unroll
      for (int q = 0; q < CNT; ++q)
        if (CNT == 2 || q < cnt) {
          FT uu = R.ai[q][0] * x0[k][q] + R.ai[q][1] * x1[k][q];
          FT vv = R.ai[q][2] * x0[k][q] + R.ai[q][3] * x1[k][q];
          su += R.w[q] * uu; sv += R.w[q] * vv;
        }
#pragma unroll
      for (int q = 0; q < CNT; ++q)
        if (off[k][q] >= 0) {
          p0[off[k][q]] = R.a[q][0] * su + R.a[q][1] * sv;
          p1[off[k][q]] = R.a[q][2] * su + R.a[q][3] * sv;
        }
    }
  }
}

template <class FT, int NI, int PAIRS, bool HALO, bool PACK = false>
__global__ void __launch_bounds__(256) k_dss2(DssArgs A, const DssNode<FT>* __restrict__ rec, int node0, int nnodes, int nh, P2PWait W,
                                              P2PArgs PA = P2PArgs(), P2PPlan Q = P2PPlan()) {
  __shared__ DssNode<FT> sr[4];
  __shared__ int pk_nodes[16];
  pdl_launch();
  const int v = threadIdx.x;
  const int npack = PACK ? Q.nblocks : 0;
  if (PACK && (int)blockIdx.x < npack) {  // the first blocks send this rank's boundary columns
    pdl_wait();
    p2p_pack_body<FT>(PA, Q, blockIdx.x, pk_nodes);
    return;
  }
  const int node = node0 + ((int)blockIdx.x - npack) * 4 + threadIdx.y;  // records [node0, nnodes)
  constexpr int RW = sizeof(DssNode<FT>) / 4;
  if (node < nnodes && v < RW) reinterpret_cast<uint32_t*>(&sr[threadIdx.y])[v] = reinterpret_cast<const uint32_t*>(&rec[node])[v];
  if (RW > 64 && node < nnodes && v + 64 < RW)
    reinterpret_cast<uint32_t*>(&sr[threadIdx.y])[v + 64] = reinterpret_cast<const uint32_t*>(&rec[node])[v + 64];
  __syncthreads();
  pdl_wait();                             // earlier kernels of this stream (incl. the pack kernel that advanced *seq) are complete
  if (HALO && W.seq) p2p_block_wait(W);   // … and now the neighbours' slabs of this exchange have landed
  if (node >= nnodes) return;
  const DssNode<FT>& R = sr[threadIdx.y];
  const int cnt = R.cnt;  // uniform over the two warps of a node
  if (cnt == 2) dss_body<FT, NI, PAIRS, 2, HALO>(A, R, cnt, v, nh);
  else dss_body<FT, NI, PAIRS, 4, HALO>(A, R, cnt, v, nh);
}

// copy element slabs (all components of one field) of the listed elements into a packed buffer
template <class FT>
__global__ void k_pack(const FT* __restrict__ src, FT* __restrict__ dst, const int* __restrict__ elems, int slab) {
  const int e = elems[blockIdx.x];
  const FT* s = src + (size_t)e * slab;
  FT* d = dst + (size_t)blockIdx.x * slab;
  for (int i = threadIdx.x; i < slab; i += blockDim.x) d[i] = s[i];
}

// u₃ boundary filter of cache_imp! (precomputed_quantities.jl:486-554, flat surface: ᶠuₕ³ = 0 ⇒ Y.f.u₃ = 0 on faces ½ and Nv+½), in place:
// the fused stepper applies it to the incoming state so that a C-ABI caller need not have called b200_cache_imp first.
template <class FT>
__global__ void __launch_bounds__(256) k_u3_filter(FT* Yf, int ncols, int nlev) {
  pdl_launch();
  pdl_wait(Yf);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * ncols) Yf[(size_t)(i >> 1) * nlev + ((i & 1) ? nlev - 1 : 0)] = FT(0);
}

constexpr int AXPY_MAX = 8;
template <class FT>
struct AxpyArgs {
  const FT* T[AXPY_MAX];
  FT c[AXPY_MAX];
  int n;
  unsigned dmask;  // bit k set: term k enters as c_k·(T_k − base) (a stage solution N_j measured from u, see impl_step)
};

// `nlev` > 0 marks a face field with nlev levels per column whose first and last level are forced to
// zero (the u₃ impenetrability filter of cache_imp!, folded into the increment by the native stepper).
template <class FT, int VEC>
__global__ void __launch_bounds__(256) k_axpy_n(FT* out, const FT* base, AxpyArgs<FT> A, size_t nvec, int nlev) {
  struct alignas(sizeof(FT) * VEC) Vt { FT x[VEC]; };
  pdl_launch();
  pdl_wait(out, base);
#pragma unroll
  for (int k = 0; k < AXPY_MAX; ++k) pdl_launder(A.T[k]);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    Vt r = reinterpret_cast<const Vt*>(base)[i];
    for (int k = 0; k < A.n; ++k) {
      Vt t = reinterpret_cast<const Vt*>(A.T[k])[i];
      // explicit fma: the increment kernels (k_axpy_n, axv in k_axpy_dss / k_pack_axpy_p2p) must round identically
      if (A.dmask >> k & 1) {
        const Vt b0 = reinterpret_cast<const Vt*>(base)[i];
#pragma unroll
        for (int q = 0; q < VEC; ++q) r.x[q] = fma_(A.c[k], t.x[q] - b0.x[q], r.x[q]);
      } else {
#pragma unroll
        for (int q = 0; q < VEC; ++q) r.x[q] = fma_(A.c[k], t.x[q], r.x[q]);
      }
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

// ---------------------------------------------------------------------------------------------
// k_axpy_dss — stage increment FUSED with the state DSS (single-rank contexts):
//     U = dss!(u + Σ_j c_j T_j)          (CTS fused_increment! followed by dss!, constrain_state.jl:59-64)
// The separate passes move (2 + n)·S for the increment and then 2·(12/16)·S for the DSS; here every unique
// perimeter node assembles the increment of each of its members on the fly (same FMA order as k_axpy_n, so the
// result is bitwise the one of k_axpy_n → k_dss2), sums, and writes the members once; the four interior nodes of
// an element are a plain increment.  The u₃ boundary filter of cache_imp! is folded in (face levels 0 and nv are
// written as zero).  N = number of tendency terms (compile time: all member loads are issued back to back).
template <class FT>
struct AxDssArgs {
  FT* out_c; FT* out_f;
  const FT* base_c; const FT* base_f;
  const FT* Tc[AXPY_MAX]; const FT* Tf[AXPY_MAX];
  FT c[AXPY_MAX];
  unsigned dmask;  // bit k set: term k enters as c_k·(T_k − base)
  int ncf, nv;
  // multi-rank: members with element index >= nh are ghosts whose ASSEMBLED slabs the owner has written into the peer-memory
  // ghost block (k_pack_axpy_p2p): [nh_ghost][ncf·16·nv] centre slabs, then [nh_ghost][16·(nv+1)] face slabs
  const FT* ghost; long long gpar; const int* seq; int nh, nh_ghost;
};
template <class FT, int N>
__device__ __forceinline__ FT axv(const FT* __restrict__ b, const FT* const* T, const FT* c, int o, unsigned dmask = 0) {
  FT t[N];
  const FT b0 = b[o];
  FT r = b0;
#pragma unroll
  for (int k = 0; k < N; ++k) t[k] = T[k][o];
#pragma unroll
  for (int k = 0; k < N; ++k) r = fma_(c[k], (dmask >> k & 1) ? t[k] - b0 : t[k], r);
  return r;
}
template <class FT, int N, int CNT, bool HALO>
__device__ __forceinline__ void axdss_body(const AxDssArgs<FT>& A, const DssNode<FT>& R, int cnt, int v) {
  const int nv = A.nv, nf = nv + 1, ec = A.ncf * 16 * nv, ef = 16 * nf;
  int oc[CNT], of[CNT];
  bool gh[CNT];
  const FT *gc = nullptr, *gf = nullptr;
  if (HALO) {
    gc = reinterpret_cast<const FT*>(reinterpret_cast<const char*>(A.ghost) + (size_t)(*A.seq & 1) * (size_t)A.gpar);
    gf = gc + (size_t)ec * A.nh_ghost;
  }
#pragma unroll
  for (int q = 0; q < CNT; ++q) {
    const int el = R.mem[q] >> 4, nd = R.mem[q] & 15;
    gh[q] = HALO && el >= A.nh && (CNT == 2 || q < cnt);
    const int le = gh[q] ? el - A.nh : el;
    oc[q] = le * ec + nd * nv + v; of[q] = le * ef + nd * nf + v;
  }
  // value of member q at centre-plane offset ko: assembled on the fly (local) or read from the ghost block (already assembled)
  auto cval = [&](int q, int ko) -> FT {
    if (!(CNT == 2 || q < cnt)) return FT(0);
    if (gh[q]) return gc[oc[q] + ko];
    return axv<FT, N>(A.base_c, A.Tc, A.c, oc[q] + ko, A.dmask);
  };
  // All loads and sums first, all stores last: out/base/T may alias as far as the compiler knows, so a store between two items
  // would pin the loads of the later item behind it (four dependent DRAM round trips per thread instead of one).
  // ρ, the Covariant12 pair (uₕ₁, uₕ₂) in the local physical basis, ρe_tot, u₃; further tracers are handled one by one below.
  const bool cvl = v < nv, fvl = v < nf;
  FT s_rho = FT(0), su = FT(0), sv = FT(0), s_re = FT(0), s_u3 = FT(0);
  if (cvl) {
    FT x[CNT], y[CNT], z[CNT], e[CNT];
#pragma unroll
    for (int q = 0; q < CNT; ++q) { x[q] = cval(q, 0); y[q] = cval(q, 16 * nv); z[q] = cval(q, 32 * nv); e[q] = cval(q, 48 * nv); }
#pragma unroll
    for (int q = 0; q < CNT; ++q) if (CNT == 2 || q < cnt) s_rho += R.w[q] * x[q];
#pragma unroll
    for (int q = 0; q < CNT; ++q)
      if (CNT == 2 || q < cnt) {
        FT uu = R.ai[q][0] * y[q] + R.ai[q][1] * z[q];
        FT vv = R.ai[q][2] * y[q] + R.ai[q][3] * z[q];
        su += R.w[q] * uu; sv += R.w[q] * vv;
      }
#pragma unroll
    for (int q = 0; q < CNT; ++q) if (CNT == 2 || q < cnt) s_re += R.w[q] * e[q];
  }
  if (fvl && v > 0 && v < nv) {
    FT x[CNT];
#pragma unroll
    for (int q = 0; q < CNT; ++q)
      x[q] = !(CNT == 2 || q < cnt) ? FT(0) : (gh[q] ? gf[of[q]] : axv<FT, N>(A.base_f, A.Tf, A.c, of[q], A.dmask));
#pragma unroll
    for (int q = 0; q < CNT; ++q) if (CNT == 2 || q < cnt) s_u3 += R.w[q] * x[q];
  }
  if (cvl) {
#pragma unroll
    for (int q = 0; q < CNT; ++q)
      if ((CNT == 2 || q < cnt) && !gh[q]) {
        A.out_c[oc[q]] = s_rho;
        A.out_c[oc[q] + 16 * nv] = R.a[q][0] * su + R.a[q][1] * sv;
        A.out_c[oc[q] + 32 * nv] = R.a[q][2] * su + R.a[q][3] * sv;
        A.out_c[oc[q] + 48 * nv] = s_re;
      }
  }
  if (fvl) {
#pragma unroll
    for (int q = 0; q < CNT; ++q) if ((CNT == 2 || q < cnt) && !gh[q]) A.out_f[of[q]] = s_u3;
  }
  if (cvl) {
    for (int k = 4; k < A.ncf; ++k) {  // passive tracers
      const int ko = k * 16 * nv;
      FT x[CNT];
#pragma unroll
      for (int q = 0; q < CNT; ++q) x[q] = cval(q, ko);
      FT s = FT(0);
#pragma unroll
      for (int q = 0; q < CNT; ++q) if (CNT == 2 || q < cnt) s += R.w[q] * x[q];
#pragma unroll
      for (int q = 0; q < CNT; ++q) if ((CNT == 2 || q < cnt) && !gh[q]) A.out_c[oc[q] + ko] = s;
    }
  }
}
// Multi-rank: assemble the state of the SEND elements and write it straight into the neighbours' ghost blocks (peer memory),
// same arithmetic as k_axpy_n including the u₃ boundary filter — the receiving k_axpy_dss reads assembled values.
template <class FT, int N>
__device__ __forceinline__ void p2p_pack_axpy_body(const AxDssArgs<FT>& A, const P2PPlan& Q, int slot, int* nodes) {
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
  const int e = Q.send_elems[slot], q = Q.slot_nbr[slot], g = Q.slot_dst[slot];
  const int cnt = p2p_nodes((unsigned)Q.slot_mask[slot], nodes);
  const int nv = A.nv, nf = nv + 1, ec = A.ncf * 16 * nv, ef = 16 * nf;
  const int value = *Q.S.seq + 1;
  FT* base = reinterpret_cast<FT*>(Q.dst[(value & 1) * Q.S.nn + q]);
  FT* dc = base + (size_t)g * ec;
  FT* df = base + (size_t)ec * Q.nbr_nh_ghost[q] + (size_t)g * ef;
  for (int i = tid; i < A.ncf * cnt * nv; i += nt) {
    const int lev = i % nv, t = i / nv, o = ((t / cnt) * 16 + nodes[t % cnt]) * nv + lev;
    dc[o] = axv<FT, N>(A.base_c, A.Tc, A.c, e * ec + o, A.dmask);
  }
  for (int i = tid; i < cnt * nf; i += nt) {
    const int lev = i % nf, o = nodes[i / nf] * nf + lev;
    df[o] = (lev == 0 || lev == nv) ? FT(0) : axv<FT, N>(A.base_f, A.Tf, A.c, e * ef + o, A.dmask);
  }
  p2p_block_done(Q.S, Q.nblocks, value);
}
// Records [node0, nnodes) in nbn node blocks; nint interior blocks (0 = none) interleaved with them.
template <class FT, int N, bool HALO, bool PACK = false>
__global__ void __launch_bounds__(256) k_axpy_dss(AxDssArgs<FT> A, const DssNode<FT>* __restrict__ rec, int node0, int nnodes, int nbn, int nint,
                                                   P2PWait W, P2PPlan Q = P2PPlan()) {
  __shared__ DssNode<FT> sr[4];
  __shared__ int pk_nodes[16];
  pdl_launch();
  const int v = threadIdx.x;
  const int npack = PACK ? Q.nblocks : 0;
  if (PACK && (int)blockIdx.x < npack) {  // the first blocks assemble and send this rank's boundary columns
    pdl_wait();
    p2p_pack_axpy_body<FT, N>(A, Q, blockIdx.x, pk_nodes);
    return;
  }
  const long long tot = (long long)nbn + nint, b = (long long)blockIdx.x - npack;
  const int ib = (int)(b * nint / tot);               // interior blocks before this one
  if ((int)((b + 1) * nint / tot) != ib) {            // this block is interior block ib
    pdl_wait();
    const int e = ib, nv = A.nv, nf = nv + 1;
    const int nd = 5 + (threadIdx.y & 1) + 4 * (threadIdx.y >> 1);  // nodes (j, i) ∈ {1,2}²
    const int o = e * A.ncf * 16 * nv + nd * nv + v, of_ = e * 16 * nf + nd * nf + v;
    FT r[4], rf = FT(0);
    if (v < nv) {
#pragma unroll
      for (int k = 0; k < 4; ++k) r[k] = axv<FT, N>(A.base_c, A.Tc, A.c, o + k * 16 * nv, A.dmask);
    }
    if (v > 0 && v < nv) rf = axv<FT, N>(A.base_f, A.Tf, A.c, of_, A.dmask);
    if (v < nv) {
#pragma unroll
      for (int k = 0; k < 4; ++k) A.out_c[o + k * 16 * nv] = r[k];
    }
    if (v < nf) A.out_f[of_] = rf;
    if (v < nv)
      for (int k = 4; k < A.ncf; ++k) A.out_c[o + k * 16 * nv] = axv<FT, N>(A.base_c, A.Tc, A.c, o + k * 16 * nv, A.dmask);
    return;
  }
  const int node = node0 + ((int)b - ib) * 4 + threadIdx.y;
  constexpr int RW = sizeof(DssNode<FT>) / 4;
  if (node < nnodes && v < RW) reinterpret_cast<uint32_t*>(&sr[threadIdx.y])[v] = reinterpret_cast<const uint32_t*>(&rec[node])[v];
  if (RW > 64 && node < nnodes && v + 64 < RW)
    reinterpret_cast<uint32_t*>(&sr[threadIdx.y])[v + 64] = reinterpret_cast<const uint32_t*>(&rec[node])[v + 64];
  __syncthreads();
  pdl_wait();
  if (HALO && W.seq) p2p_block_wait(W);
  if (node >= nnodes) return;
  const DssNode<FT>& R = sr[threadIdx.y];
  const int cnt = R.cnt;
  if (cnt == 2) axdss_body<FT, N, 2, HALO>(A, R, cnt, v);
  else axdss_body<FT, N, 4, HALO>(A, R, cnt, v);
}

// out = (a - b) * s   (T_imp[i] = (U - temp)/dtγ)
template <class FT, int VEC>
__global__ void __launch_bounds__(256) k_diff_scale(FT* out, const FT* a, const FT* b, FT s, size_t nvec) {
  struct alignas(sizeof(FT) * VEC) Vt { FT x[VEC]; };
  pdl_launch();
  pdl_wait(out, a, b);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    Vt x = reinterpret_cast<const Vt*>(a)[i], y = reinterpret_cast<const Vt*>(b)[i], r;
#pragma unroll
    for (int q = 0; q < VEC; ++q) r.x[q] = (x.x[q] - y.x[q]) / s;
    reinterpret_cast<Vt*>(out)[i] = r;
  }
}

}  // namespace b200
