// capi.cu — C-ABI of libb200dycore.so (declared in include/b200_dycore.h).
// Context creation (geometry/topology ingestion), hook entry points and the native ARS343 stepper.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <limits.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/b200_dycore.h"
#include "kernels_dss.cuh"
#include "kernels_implicit.cuh"
#include "kernels_row.cuh"
#include "kernels_pair.cuh"
#include "kernels_lvl.cuh"
#include "kernels_imp5.cuh"
#include "kernels_limiter.cuh"
#include "kernels_vdiff.cuh"
#include "kernels_imp8.cuh"
#include "kernels_imp8d.cuh"

using namespace b200;

static thread_local std::string g_err;
static thread_local b200_ctx* g_cur = nullptr;  // context of the entry point being executed on this thread
static void ctx_set_err(b200_ctx* c, const std::string& m);
static int fail(const std::string& m) {
  g_err = m;
  if (g_cur) ctx_set_err(g_cur, m);
  return -1;
}
// A halo wait of an earlier call timed out (a peer rank died or fell out of step): the state is invalid; say so instead of computing on.
struct b200_ctx;
static int halo_failed(const b200_ctx* c, const char* who);
struct CtxScope {
  b200_ctx* prev;
  explicit CtxScope(b200_ctx* c) : prev(g_cur) { g_cur = c; }
  ~CtxScope() { g_cur = prev; }
};
#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) return fail(std::string(#x) + ": " + cudaGetErrorString(e_)); \
  } while (0)

// ---- NCCL through dlopen (the torch-bundled libnccl.so.2 is already resident in the host process)
typedef struct ncclComm* ncclComm_t;
struct Id128 { char b[128]; };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, /* ncclUniqueId by value */ Id128, int) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int nccl_load() {
  if (g_nccl.lib) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) return fail("dlopen(libnccl.so.2) failed: multi-GPU DSS halo needs NCCL");
#define SYM(f, name)                                                   \
  *(void**)(&g_nccl.f) = dlsym(g_nccl.lib, name);                      \
  if (!g_nccl.f) return fail(std::string("dlsym failed: ") + name);
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  return 0;
}
#define NK(x)                                                                                     \
  do {                                                                                            \
    int r_ = (x);                                                                                 \
    if (r_ != 0) return fail(std::string(#x) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "nccl error")); \
  } while (0)

struct b200_ctx {
  std::string err;  // message of the last failing call on this context (b200_last_error)
  b200_dims dims;
  b200_params prm;
  int ft = 4;
  void* d_hgeo = nullptr;
  void* d_vlev = nullptr;
  int* d_off = nullptr;
  int* d_mem = nullptr;
  int nnodes = 0;
  std::vector<int32_t> h_off, h_mem;
  void* d_dssrec = nullptr;
  int32_t* d_node_off = nullptr;
  int32_t *d_lim_nbr_off = nullptr, *d_lim_nbr = nullptr;  // vertex neighbours of every local element (limiter bounds)
  void* d_lim_bnd = nullptr;    // [n_tracers][nh][2][64] element bounds of q
  void* d_lim_E = nullptr;      // multi-rank: the bounds as a centre-shaped field [nh][2·n_tracers][16][nv] for the halo
  int32_t* d_lim_ghost_node = nullptr;  // per ghost element: a node whose column this rank receives
  void* Tlc[4] = {nullptr, nullptr, nullptr, nullptr};  // T_lim of the four stages (stepper, limiter on)  // records [d_node_off[e], d_node_off[e+1]) are owned by local element e
  void* d_jac = nullptr;
  void *d_jsnap_c = nullptr, *d_jsnap_f = nullptr;  // snapshot of the state Wfact was called with (dry path: ldiv! recomputes the coefficients)
  double jsnap_dtg = 0; bool jac_planes_valid = false;
  void* d_jacd = nullptr;  // vertical-diffusion Jacobian planes (k_vdiff_jac)
  void* d_kdec = nullptr;  // [LV] DecayWithHeightDiffusion K(z_c)
  // native stepper storage (allocated lazily)
  void *Uc[2] = {nullptr, nullptr}, *Uf[2] = {nullptr, nullptr};
  void *Tec[4] = {}, *Tef[4] = {}, *Tic[4] = {}, *Tif[4] = {};
  void* H = nullptr;
  void* Hw = nullptr;  // moist (0M): ρ(h_eff + Φ) [nh][16][nv] for the water enthalpy flux of the hyperdiffusion apply
  void *Rc = nullptr, *Rf = nullptr, *dc = nullptr, *df = nullptr;
  // halo
  int rank = 0, nranks = 1;
  ncclComm_t comm = nullptr;
  std::vector<int32_t> nbr, send_off, recv_off;
  int* d_send_elems = nullptr;
  int* d_slot_mask = nullptr;   // per send slot: bit n set ⇔ node n of the element is shared with the receiving rank
  int n_send = 0;
  void *sendbuf = nullptr, *ghostbuf = nullptr;  // sized for the largest DSS call
  size_t halo_cap = 0;
  // peer-memory halo
  void* p2p_buf = nullptr;      // [2 parities][p2p_cap bytes] ghost slabs written by the neighbours, then int flags[nranks]
  size_t p2p_cap = 0;
  bool p2p_ready = false;
  int* h_p2p_err = nullptr;     // host-mapped error word written by a halo wait that timed out (kernels_dss.cuh: g_p2p_err)
  int* d_p2p_seq = nullptr;     // device-side exchange number, followed by the pack-kernel completion counter (kernels_dss.cuh: P2PSig)
  int n_int_nodes = 0;          // records [0, n_int_nodes) have no ghost member (sorted first)
  std::vector<void*> p2p_peer;  // mapped neighbour buffers
  void** d_p2p_dst = nullptr;   // device array [n_neighbors] of destination base pointers (rewritten per call)
  int** d_p2p_flags = nullptr;  // device array [n_neighbors]: address of my flag in neighbour q
  int *d_slot_nbr = nullptr, *d_slot_dst = nullptr, *d_nbr_nhg = nullptr, *d_nbr_rank = nullptr;
  int64_t launches = 0;
  double dz_sfc = 0;  // Δz of the lowest cell (VerticalDiffusion: z_a = Δz/2)
  // CUDA graph of one fused step (single-rank contexts): captured on the second call with the same (Yc, Yf, stream)
  int use_graph = 1;  // B200_GRAPH=0 disables
  int pdl = 63;       // B200_PDL=<bit mask>: programmatic dependent launch per kernel group (1 exp_a, 2 exp_c, 4 dss2, 8 axpy, 16 imp, 32 diff); 0 = off
  int zform = 1;       // B200_ZFORM=0: the fused stepper forms T_imp[j] = (N_j − U_j)/dtγ (side stream) instead of using the stage solutions
  void *Nsc[4] = {nullptr, nullptr, nullptr, nullptr}, *Nsf[4] = {nullptr, nullptr, nullptr, nullptr};  // stage solutions N_j (zform)
  int imp_kernel = 8;  // B200_IMP_KERNEL=5: the shared-memory slab version k5_imp_stage (A/B runs; LDIV and moist contexts always use it)
  int stiff_final = 1; // B200_STIFF_FINAL=0: literal final increment u + dt Σ b_j (T_exp[j] + T_imp[j]) in the fused path too
  int fuse_axdss = 1; // B200_FUSE_AXDSS=0: stage increment and state DSS as two passes (k_axpy_n, k_dss2) instead of k_axpy_dss
  struct StepGraph { cudaGraphExec_t exec; void *Yc, *Yf; int64_t launches; };
  std::vector<StepGraph> graphs;       // small cache (double-buffered callers alternate between two states)
  cudaStream_t gstream = nullptr;      // capture/replay stream used when the caller passes the legacy default stream
  cudaEvent_t ev_gin = nullptr, ev_gout = nullptr;
  int eager_steps = 0;
  cudaStream_t side = nullptr;           // side stream: T_imp = (U − temp)/dtγ runs concurrently with the T_exp kernels
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int generic_nv = 0;  // B200_GENERIC_NV=1: do not use the kernels specialised for nv = 63 (test coverage of the run-time-nv builds)
  int ncf() const { return 4 + dims.n_tracers; }
  size_t nc() const { return (size_t)dims.nh * ncf() * 16 * dims.nv; }
  size_t nf() const { return (size_t)dims.nh * 16 * (dims.nv + 1); }
};

// Error text: per context — every entry point that takes a context opens a CtxScope, so fail() records the message in
// that context; ctx == NULL returns the calling thread's last message (b200_create failures, host-only helpers).
extern "C" const char* b200_last_error(const b200_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }
static void ctx_set_err(b200_ctx* c, const std::string& m) { c->err = m; }

extern "C" int b200_nccl_unique_id(void* out128) {
  if (nccl_load()) return -1;
  NK(g_nccl.GetUniqueId(out128));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// topology → CSR of unique perimeter nodes (must match climaatmos.jl_b200/grid.py:dss_node_csr)
static void face_node(int f, int q, int& i, int& j) {
  switch (f) {
    case 0: i = q; j = 0; break;
    case 1: i = 3; j = q; break;
    case 2: i = 3 - q; j = 3; break;
    default: i = 0; j = 3 - q; break;
  }
}
static void vert_node(int v, int& i, int& j) {
  static const int vi[4] = {0, 3, 3, 0}, vj[4] = {0, 0, 3, 3};
  i = vi[v]; j = vj[v];
}

static void build_csr(const b200_topology* T, std::vector<int32_t>& off, std::vector<int32_t>& mem) {
  off.clear(); mem.clear();
  off.push_back(0);
  for (int v = 0; v < T->n_verts; ++v) {
    for (int q = T->local_vertex_offset[v]; q < T->local_vertex_offset[v + 1]; ++q) {
      int e = T->local_vertices[2 * q], vert = T->local_vertices[2 * q + 1], i, j;
      vert_node(vert, i, j);
      mem.push_back(e * 16 + j * 4 + i);
    }
    off.push_back((int32_t)mem.size());
  }
  for (int f = 0; f < T->n_faces; ++f) {
    const int32_t* F = T->interior_faces + 5 * f;
    for (int q = 1; q < 3; ++q) {
      int q2 = F[4] ? 3 - q : q, i, j;
      face_node(F[1], q, i, j);
      mem.push_back(F[0] * 16 + j * 4 + i);
      face_node(F[3], q2, i, j);
      mem.push_back(F[2] * 16 + j * 4 + i);
      off.push_back((int32_t)mem.size());
    }
  }
  if (T->elem_gid) {  // rank-count independent summation order: ascending global element id
    for (size_t n = 0; n + 1 < off.size(); ++n)
      std::stable_sort(mem.begin() + off[n], mem.begin() + off[n + 1],
                       [&](int32_t a, int32_t b) { return T->elem_gid[a >> 4] < T->elem_gid[b >> 4]; });
  }
}

extern "C" int b200_build_dss_csr(const b200_topology* T, int32_t* off_out, int32_t cap_nodes, int32_t* mem_out, int32_t cap_mem,
                                  int32_t* nnodes, int32_t* nmem) {
  std::vector<int32_t> off, mem;
  build_csr(T, off, mem);
  *nnodes = (int32_t)off.size() - 1; *nmem = (int32_t)mem.size();
  if ((int32_t)off.size() > cap_nodes + 1 || (int32_t)mem.size() > cap_mem) return fail("b200_build_dss_csr: capacity too small");
  memcpy(off_out, off.data(), off.size() * sizeof(int32_t));
  memcpy(mem_out, mem.data(), mem.size() * sizeof(int32_t));
  return 0;
}

extern "C" int b200_debug_dss_csr(b200_ctx* c, const int32_t** off, const int32_t** mem, int32_t* nnodes, int32_t* nmem) {
  CtxScope scope_(c);
  if (!c) return fail("b200_debug_dss_csr: null context");
  *off = c->h_off.data(); *mem = c->h_mem.data();
  *nnodes = c->nnodes; *nmem = (int32_t)c->h_mem.size();
  return 0;
}

template <class FT>
static Par<FT> make_par(const b200_ctx* c) {
  const b200_params& p = c->prm;
  Par<FT> P;
  P.R_d = (FT)p.R_d; P.cp_d = (FT)p.cp_d; P.cv_d = (FT)p.cv_d; P.T_0 = (FT)p.T_0; P.p0 = (FT)p.p_ref_theta;
  P.kappa = (FT)(p.R_d / p.cp_d); P.Ts_ref = (FT)p.T_surf_ref; P.Tmin_ref = (FT)p.T_min_ref;
  P.T_min_sgs = (FT)p.T_min_sgs; P.dt = (FT)p.dt;
  P.icv = (FT)(1.0 / p.cv_d); P.ip0 = (FT)(1.0 / p.p_ref_theta); P.dTs7 = (FT)((p.T_surf_ref - p.T_min_ref) / 7.0);
  P.RT0 = (FT)(p.R_d * p.T_0);
  P.hs = p.held_suarez;
  if (p.held_suarez) {
    P.hs_ka = (FT)(1.0 / (40 * p.hs_day)); P.hs_ks = (FT)(1.0 / (4 * p.hs_day)); P.hs_kf = (FT)(1.0 / p.hs_day);
    P.hs_sigb = (FT)p.hs_sigma_b; P.hs_isig = (FT)(1.0 / (1.0 - p.hs_sigma_b)); P.hs_dTy = (FT)p.hs_dT_y; P.hs_Teq = (FT)p.hs_T_equator;
    P.hs_dthz = (FT)p.hs_dtheta_z; P.hs_Tmin = (FT)p.hs_T_min; P.hs_iMSLP = (FT)(1.0 / p.MSLP); P.hs_ikap = (FT)(p.cp_d / p.R_d);
  } else {
    P.hs_ka = P.hs_ks = P.hs_kf = P.hs_sigb = P.hs_isig = P.hs_dTy = P.hs_Teq = P.hs_dthz = P.hs_Tmin = P.hs_iMSLP = P.hs_ikap = (FT)0;
  }
  P.nu4v = (FT)p.nu4_vorticity; P.nu4s = (FT)p.nu4_scalar; P.ddf = (FT)p.divergence_damping_factor;
  P.nh = c->dims.nh; P.nv = c->dims.nv; P.ncf = c->ncf(); P.tupw = p.tracer_upwinding;
  P.hyperdiff = p.hyperdiff; P.rayleigh = p.rayleigh_sponge; P.viscous = p.viscous_sponge; P.upwinding = p.energy_upwinding;
  P.moist = p.microphysics_0M;
  memset(&P.M, 0, sizeof(P.M));
  if (p.microphysics_0M) {
    P.M.R_v = (FT)p.R_v; P.M.cv_v = (FT)(p.cp_v - p.R_v); P.M.cp_v = (FT)p.cp_v; P.M.cp_l = (FT)p.cp_l; P.M.cp_i = (FT)p.cp_i;
    P.M.LH_v0 = (FT)p.LH_v0; P.M.LH_s0 = (FT)p.LH_s0; P.M.e_v0 = (FT)(p.LH_v0 - p.R_v * p.T_0); P.M.e_i0 = (FT)(p.LH_s0 - p.LH_v0);
    P.M.T_tr = (FT)p.T_triple; P.M.ln_ptr = (FT)log(p.press_triple); P.M.T_frz = (FT)p.T_freeze; P.M.T_icn = (FT)p.T_icenuc;
    P.M.pow_icn = (FT)p.pow_icenuc;
    P.M.A_liq = (FT)((p.cp_v - p.cp_l) / p.R_v); P.M.A_ice = (FT)((p.cp_v - p.cp_i) / p.R_v);
    P.M.B_liq = (FT)((p.LH_v0 - (p.cp_v - p.cp_l) * p.T_0) / p.R_v); P.M.B_ice = (FT)((p.LH_s0 - (p.cp_v - p.cp_i) * p.T_0) / p.R_v);
    P.M.iT_tr = (FT)(1.0 / p.T_triple); P.M.epsv = (FT)(p.R_v / p.R_d);
    P.M.q_neg = (FT)(0.25 * (sizeof(FT) == 4 ? 1.1920929e-07 : 2.220446049250313e-16) * p.cv_d / p.LH_s0);
  }
  return P;
}

template <class FT>
static int create_geo(b200_ctx* c, const b200_geometry* G, const b200_params* p) {
  const int nht = c->dims.nh + c->dims.nh_ghost, nv = c->dims.nv, nf = nv + 1;
  // ---- per-level constants
  VLev<FT> V;
  memset(&V, 0, sizeof(V));
  const double R = G->radius;
  auto sfun = [&](double z) { return c->dims.deep ? (R + z) / R : 1.0; };
  auto zeta = [&](double z, double zd) { double s = sin(M_PI * ((z - zd) / (G->z_max - zd) / 2)); return s * s; };
  for (int v = 0; v < nv; ++v) {
    double s = sfun(G->z_c[v]);
    V.sc2i[v] = (FT)(1.0 / (s * s));
    V.dzc[v] = (FT)G->dz_c[v];
    V.mc[v] = (FT)(s * s * G->dz_c[v]);
    V.rmc[v] = (FT)(1.0 / (s * s * G->dz_c[v]));
    V.phic[v] = (FT)(p->grav * G->z_c[v]);
    V.bruh[v] = (FT)(p->rayleigh_sponge && G->z_c[v] > p->zd_rayleigh ? p->alpha_rayleigh_uh * zeta(G->z_c[v], p->zd_rayleigh) : 0.0);
    V.bvc[v] = (FT)(p->viscous_sponge && G->z_c[v] > p->zd_viscous ? p->kappa_2_sponge * zeta(G->z_c[v], p->zd_viscous) : 0.0);
  }
  for (int f = 0; f < nf; ++f) {
    double s = sfun(G->z_f[f]);
    V.sf2i[f] = (FT)(1.0 / (s * s));
    V.sf[f] = (FT)s;
    V.dzf[f] = (FT)G->dz_f[f];
    V.g33f[f] = (FT)(1.0 / (G->dz_f[f] * G->dz_f[f]));
    V.dphif[f] = (f > 0 && f < nv) ? (FT)((FT)(p->grav * G->z_c[f]) - (FT)(p->grav * G->z_c[f - 1])) : (FT)0;
    V.brw[f] = (FT)(p->rayleigh_sponge && G->z_f[f] > p->zd_rayleigh ? p->alpha_rayleigh_w * zeta(G->z_f[f], p->zd_rayleigh) : 0.0);
    V.bvf[f] = (FT)(p->viscous_sponge && G->z_f[f] > p->zd_viscous ? p->kappa_2_sponge * zeta(G->z_f[f], p->zd_viscous) : 0.0);
  }
  for (int i = 0; i < 4; ++i)
    for (int k = 0; k < 4; ++k) {
      V.D[i * 4 + k] = (FT)G->gll_D[i * 4 + k];
      V.Dw[i * 4 + k] = (FT)(-G->gll_D[k * 4 + i] * G->gll_w[k] / G->gll_w[i]);
    }
  CK(cudaMalloc(&c->d_vlev, sizeof(V)));
  CK(cudaMemcpy(c->d_vlev, &V, sizeof(V), cudaMemcpyHostToDevice));
  if (p->vert_diff) {  // eddy_diffusivity_coefficient_H (precomputed_quantities.jl:652-654): D₀ exp(−(z − z_sfc)/H)
    FT kd[LV];
    for (int v = 0; v < LV; ++v) kd[v] = v < nv ? (FT)(p->D_0_diffusion * exp(-(G->z_c[v] - G->z_f[0]) / p->H_diffusion)) : (FT)0;
    CK(cudaMalloc(&c->d_kdec, sizeof(kd)));
    CK(cudaMemcpy(c->d_kdec, kd, sizeof(kd), cudaMemcpyHostToDevice));
    c->dz_sfc = G->dz_c[0];
  }
  {  // derivative matrices in the constant bank (register-resident kernels)
    float mf[32]; double md[32];
    for (int k = 0; k < 16; ++k) { md[k] = (double)V.D[k]; md[16 + k] = (double)V.Dw[k]; mf[k] = (float)V.D[k]; mf[16 + k] = (float)V.Dw[k]; }
    if (sizeof(FT) == 4) CK(cudaMemcpyToSymbol(c_Df, mf, sizeof(mf)));
    else CK(cudaMemcpyToSymbol(c_Dd, md, sizeof(md)));
    // paired layout for the FFMA2 kernels: [(w*4 + k)*2 + p] = (M_w[2p][k], M_w[2p+1][k])
    float2 pf[16]; double2 pd[16];
    for (int w = 0; w < 2; ++w)
      for (int k = 0; k < 4; ++k)
        for (int pp = 0; pp < 2; ++pp) {
          double lo = md[w * 16 + (2 * pp) * 4 + k], hi = md[w * 16 + (2 * pp + 1) * 4 + k];
          pd[(w * 4 + k) * 2 + pp] = make_double2(lo, hi);
          pf[(w * 4 + k) * 2 + pp] = make_float2((float)lo, (float)hi);
        }
    if (sizeof(FT) == 4) CK(cudaMemcpyToSymbol(c_Pf, pf, sizeof(pf)));
    else CK(cudaMemcpyToSymbol(c_Pd, pd, sizeof(pd)));
  }
  // ---- horizontal geometry
  std::vector<double> WJ((size_t)nht * 16), tot((size_t)nht * 16);
  for (int h = 0; h < nht; ++h)
    for (int n = 0; n < 16; ++n) {
      int i = n & 3, j = n >> 2;
      WJ[h * 16 + n] = G->gll_w[i] * G->gll_w[j] * G->J2[h * 16 + n];
      tot[h * 16 + n] = WJ[h * 16 + n];
    }
  for (int nd = 0; nd < c->nnodes; ++nd) {
    double s = 0;
    for (int q = c->h_off[nd]; q < c->h_off[nd + 1]; ++q) s += WJ[c->h_mem[q]];
    for (int q = c->h_off[nd]; q < c->h_off[nd + 1]; ++q) tot[c->h_mem[q]] = s;
  }
  std::vector<FT> hg((size_t)nht * HG_N * 16);
  for (int h = 0; h < nht; ++h)
    for (int n = 0; n < 16; ++n) {
      const double* A = G->dxdxi + ((size_t)h * 16 + n) * 4;  // A[a][b] = A[a*2+b]
      double a00 = A[0], a01 = A[1], a10 = A[2], a11 = A[3];
      double gc11 = a00 * a00 + a10 * a10, gc12 = a00 * a01 + a10 * a11, gc22 = a01 * a01 + a11 * a11;
      double det = gc11 * gc22 - gc12 * gc12, dA = a00 * a11 - a01 * a10;
      double lat = G->lat[h * 16 + n] * M_PI / 180.0;
      double fv = 2 * p->Omega * cos(lat), fw = 2 * p->Omega * sin(lat);
      double ai00 = a11 / dA, ai01 = -a01 / dA, ai10 = -a10 / dA, ai11 = a00 / dA;
      FT* o = hg.data() + (size_t)h * HG_N * 16 + n;
      o[HG_J2 * 16] = (FT)G->J2[h * 16 + n];
      o[HG_RJ2 * 16] = (FT)(1.0 / G->J2[h * 16 + n]);
      o[HG_GI11 * 16] = (FT)(gc22 / det); o[HG_GI12 * 16] = (FT)(-gc12 / det); o[HG_GI22 * 16] = (FT)(gc11 / det);
      o[HG_GC11 * 16] = (FT)gc11; o[HG_GC12 * 16] = (FT)gc12; o[HG_GC22 * 16] = (FT)gc22;
      o[HG_COR1 * 16] = (FT)(c->dims.deep ? ai01 * fv : 0.0);
      o[HG_COR2 * 16] = (FT)(c->dims.deep ? ai11 * fv : 0.0);
      o[HG_COR3 * 16] = (FT)fw;
      o[HG_SIN2 * 16] = (FT)(sin(lat) * sin(lat)); o[HG_COS2 * 16] = (FT)(cos(lat) * cos(lat));
      o[HG_DSSW * 16] = (FT)(WJ[h * 16 + n] / tot[h * 16 + n]);
      o[HG_WJ * 16] = (FT)WJ[h * 16 + n];
      o[HG_A00 * 16] = (FT)a00; o[HG_A01 * 16] = (FT)a01; o[HG_A10 * 16] = (FT)a10; o[HG_A11 * 16] = (FT)a11;
      o[HG_AI00 * 16] = (FT)ai00; o[HG_AI01 * 16] = (FT)ai01; o[HG_AI10 * 16] = (FT)ai10; o[HG_AI11 * 16] = (FT)ai11;
    }
  CK(cudaMalloc(&c->d_hgeo, hg.size() * sizeof(FT)));
  CK(cudaMemcpy(c->d_hgeo, hg.data(), hg.size() * sizeof(FT), cudaMemcpyHostToDevice));
  // per-node DSS records (same arithmetic as k_dss: weight·(A⁻¹)ᵀ in, Aᵀ out, members in summation order)
  std::vector<DssNode<FT>> rec((size_t)std::max(1, c->nnodes));
  memset(rec.data(), 0, rec.size() * sizeof(DssNode<FT>));
  for (int nd = 0; nd < c->nnodes; ++nd) {
    DssNode<FT>& R = rec[nd];
    R.cnt = c->h_off[nd + 1] - c->h_off[nd];
    for (int q = 0; q < R.cnt; ++q) {
      const int m = c->h_mem[c->h_off[nd] + q];
      const FT* o = hg.data() + (size_t)(m >> 4) * HG_N * 16 + (m & 15);
      R.mem[q] = m;
      R.w[q] = o[HG_DSSW * 16];
      R.ai[q][0] = o[HG_AI00 * 16]; R.ai[q][1] = o[HG_AI10 * 16]; R.ai[q][2] = o[HG_AI01 * 16]; R.ai[q][3] = o[HG_AI11 * 16];
      R.a[q][0] = o[HG_A00 * 16]; R.a[q][1] = o[HG_A10 * 16]; R.a[q][2] = o[HG_A01 * 16]; R.a[q][3] = o[HG_A11 * 16];
    }
  }
  // Device order of the records: by owner element (the lowest local member), then by node address, so that the unique nodes of one
  // element are processed together (k_axpy_dss runs one CTA per owner element; k_dss2 gets the same L2 locality).  The host CSR
  // (h_off/h_mem, the bit-exact index-map contract) keeps ClimaCore's vertex-then-face enumeration.
  {
    const int nh = c->dims.nh;
    std::vector<int> key(c->nnodes), perm(c->nnodes);
    std::vector<char> ghosty(c->nnodes, 0);
    c->n_int_nodes = 0;
    for (int nd = 0; nd < c->nnodes; ++nd) {
      int k = INT_MAX;
      for (int q = 0; q < rec[nd].cnt; ++q) {
        if ((rec[nd].mem[q] >> 4) < nh) k = std::min(k, rec[nd].mem[q]);
        else ghosty[nd] = 1;
      }
      key[nd] = k; perm[nd] = nd;
      c->n_int_nodes += !ghosty[nd];
    }
    // nodes without a ghost member first (they can be summed while the halo is in flight), each group by owner element
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return ghosty[a] != ghosty[b] ? ghosty[a] < ghosty[b] : key[a] < key[b]; });
    std::vector<DssNode<FT>> sorted(rec.size());
    std::vector<int32_t> noff((size_t)nh + 1, 0);
    for (int k = 0; k < c->nnodes; ++k) {
      sorted[k] = rec[perm[k]];
      const int owner = key[perm[k]] == INT_MAX ? nh - 1 : (key[perm[k]] >> 4);
      noff[owner + 1]++;
    }
    for (int e = 0; e < nh; ++e) noff[e + 1] += noff[e];
    if (c->nnodes > 0) rec.swap(sorted);
    CK(cudaMalloc(&c->d_node_off, noff.size() * sizeof(int32_t)));
    CK(cudaMemcpy(c->d_node_off, noff.data(), noff.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  CK(cudaMalloc(&c->d_dssrec, rec.size() * sizeof(DssNode<FT>)));
  CK(cudaMemcpy(c->d_dssrec, rec.data(), rec.size() * sizeof(DssNode<FT>), cudaMemcpyHostToDevice));
  return 0;
}

// shared-memory footprints (bytes)
template <class FT>
static VDiff<FT> make_vdiff(const b200_ctx* c) {
  const b200_params& p = c->prm;
  VDiff<FT> D;
  D.mode = p.vert_diff; D.momentum = !p.disable_momentum_vertical_diffusion; D.n_iters = p.approximate_linear_solve_iters;
  D.ce_za = (FT)(p.C_E * c->dz_sfc / 2); D.eps = sizeof(FT) == 4 ? (FT)1.1920928955078125e-7 : (FT)2.220446049250313e-16;
  D.cpcv = (FT)(p.cp_d / p.cv_d); D.kdec = (const FT*)c->d_kdec;
  return D;
}
static bool vdiff_implicit(const b200_ctx* c) { return c->prm.vert_diff != 0 && c->prm.implicit_diffusion != 0; }
static bool vdiff_explicit(const b200_ctx* c) { return c->prm.vert_diff != 0 && c->prm.implicit_diffusion == 0; }

template <class FT> static size_t smem_base() { return sizeof(VLev<FT>) + HG_ELEM * 16 * sizeof(FT); }
template <class FT> static size_t smem_slabs(int n) { return smem_base<FT>() + (size_t)n * SLAB * sizeof(FT); }

template <class FT> static size_t smem_row(int n) { return (HG_ELEM * 16 + (size_t)n * SLAB) * sizeof(FT); }
template <class FT> static size_t smem_rowq(int n) { return (HG_ELEM * 16 + (size_t)n * XSLAB) * sizeof(FT); }  // pair-layout slabs (k5_exp_a/c)

template <class FT>
static int set_attrs() {
  CK(cudaFuncSetAttribute(k5_exp_a<FT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rowq<FT>(9)));
  CK(cudaFuncSetAttribute(k5_exp_a<FT, 63>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rowq<FT>(9)));
  CK(cudaFuncSetAttribute(k5_tracer_a<FT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_row<FT>(3)));
  CK(cudaFuncSetAttribute(k5_imp_stage<FT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_imp5<FT>()));
  CK(cudaFuncSetAttribute(k5_imp_stage<FT, 63>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_imp5<FT>()));
  CK(cudaFuncSetAttribute(k5_imp_stage<FT, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_imp5<FT>()));
  CK(cudaFuncSetAttribute(k5_imp_stage<FT, 63, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_imp5<FT>()));
  CK(cudaFuncSetAttribute(k5_exp_a<FT, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rowq<FT>(9)));
  CK(cudaFuncSetAttribute(k5_exp_a<FT, 63, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rowq<FT>(9)));
  CK(cudaFuncSetAttribute(k5_imp_stage<FT, 0, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_imp5<FT>(true)));
  CK(cudaFuncSetAttribute(k5_imp_stage<FT, 63, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_imp5<FT>(true)));
  CK(cudaFuncSetAttribute(k5_imp_stage<FT, 0, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_imp5<FT>(true)));
  CK(cudaFuncSetAttribute(k_cache_imp<FT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_slabs<FT>(1)));
  CK(cudaFuncSetAttribute(k_lim_vborrow<FT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * SLAB * sizeof(FT))));
  CK(cudaFuncSetAttribute(k_vdiff_jac<FT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_slabs<FT>(14)));
  CK(cudaFuncSetAttribute(k_ldiv_diff<FT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((22 * SLAB + LV) * sizeof(FT))));
  CK(cudaFuncSetAttribute(k_imp_stage_diff<FT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(QD_PROFILES * 4 * LVP * sizeof(FT))));
  return 0;
}

static int halo_failed(const b200_ctx* c, const char* who) {
  if (!c->h_p2p_err) return 0;
  const int e = *reinterpret_cast<volatile int*>(c->h_p2p_err);
  if (!e) return 0;
  return fail(std::string(who) + ": the DSS halo timed out earlier (neighbour rank " + std::to_string((e & 0xff) - 1) + " never signalled exchange " +
              std::to_string(e >> 8) + "): a peer rank died or fell out of step; the state of this context is invalid");
}
extern "C" int b200_destroy(b200_ctx* c);
static int create_body(b200_ctx* c, const b200_dims* d, const b200_geometry* G, const b200_topology* T, const b200_params* p,
                       const void* nccl_id, int rank, int nranks) {
  c->dims = *d; c->prm = *p; c->ft = d->ft_bytes; c->rank = rank; c->nranks = nranks;
  if (const char* e = getenv("B200_GRAPH")) c->use_graph = atoi(e);
  if (const char* e = getenv("B200_PDL")) c->pdl = atoi(e);
  if (const char* e = getenv("B200_STIFF_FINAL")) c->stiff_final = atoi(e);
  if (const char* e = getenv("B200_IMP_KERNEL")) c->imp_kernel = atoi(e);
  if (const char* e = getenv("B200_ZFORM")) c->zform = atoi(e);
  if (const char* e = getenv("B200_FUSE_AXDSS")) c->fuse_axdss = atoi(e);
  if (const char* e = getenv("B200_GENERIC_NV")) c->generic_nv = atoi(e);
  build_csr(T, c->h_off, c->h_mem);
  {  // Topologies.local_neighboring_elements: elements sharing a vertex, from the vertex tables (limiter bounds)
    const int nh = d->nh;
    std::vector<std::vector<int32_t>> nb(nh);
    for (int v = 0; v < T->n_verts; ++v)
      for (int a = T->local_vertex_offset[v]; a < T->local_vertex_offset[v + 1]; ++a)
        for (int b = T->local_vertex_offset[v]; b < T->local_vertex_offset[v + 1]; ++b) {
          const int ea = T->local_vertices[2 * a], eb = T->local_vertices[2 * b];
          if (ea != eb && ea < nh) nb[ea].push_back(eb);  // eb >= nh: ghost neighbour (bounds arrive through the halo)
        }
    std::vector<int32_t> off(nh + 1, 0), lst;
    for (int e = 0; e < nh; ++e) {
      std::sort(nb[e].begin(), nb[e].end());
      nb[e].erase(std::unique(nb[e].begin(), nb[e].end()), nb[e].end());
      lst.insert(lst.end(), nb[e].begin(), nb[e].end());
      off[e + 1] = (int32_t)lst.size();
    }
    if (cudaMalloc(&c->d_lim_nbr_off, off.size() * sizeof(int32_t)) != cudaSuccess ||
        cudaMalloc(&c->d_lim_nbr, std::max<size_t>(1, lst.size()) * sizeof(int32_t)) != cudaSuccess) return fail("b200_create: cudaMalloc (limiter tables)");
    CK(cudaMemcpy(c->d_lim_nbr_off, off.data(), off.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_lim_nbr, lst.data(), lst.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  c->nnodes = (int)c->h_off.size() - 1;
  // keep only nodes with at least one local member
  {
    std::vector<int32_t> off2{0}, mem2;
    for (int n = 0; n < c->nnodes; ++n) {
      bool local = false;
      for (int q = c->h_off[n]; q < c->h_off[n + 1]; ++q) local |= (c->h_mem[q] >> 4) < d->nh;
      if (!local) continue;
      for (int q = c->h_off[n]; q < c->h_off[n + 1]; ++q) mem2.push_back(c->h_mem[q]);
      off2.push_back((int32_t)mem2.size());
    }
    c->h_off.swap(off2); c->h_mem.swap(mem2);
    c->nnodes = (int)c->h_off.size() - 1;
  }
  for (int n = 0; n < c->nnodes; ++n)
    if (c->h_off[n + 1] - c->h_off[n] > 4) return fail("b200_create: node shared by more than 4 elements");
  CK(cudaMalloc(&c->d_off, c->h_off.size() * sizeof(int)));
  CK(cudaMalloc(&c->d_mem, std::max<size_t>(1, c->h_mem.size()) * sizeof(int)));
  CK(cudaMemcpy(c->d_off, c->h_off.data(), c->h_off.size() * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_mem, c->h_mem.data(), c->h_mem.size() * sizeof(int), cudaMemcpyHostToDevice));
  int r = (c->ft == 4) ? create_geo<float>(c, G, p) : create_geo<double>(c, G, p);
  if (r) return r;
  r = (c->ft == 4) ? set_attrs<float>() : set_attrs<double>();
  if (r) return r;
  // halo plan
  if (nranks > 1 && T->n_neighbors > 0) {
    if (!nccl_id) return fail("b200_create: nccl_unique_id required for nranks > 1");
    if (nccl_load()) return -1;
    c->nbr.assign(T->neighbor_ranks, T->neighbor_ranks + T->n_neighbors);
    c->send_off.assign(T->send_offset, T->send_offset + T->n_neighbors + 1);
    c->recv_off.assign(T->recv_offset, T->recv_offset + T->n_neighbors + 1);
    c->n_send = c->send_off.back();
    CK(cudaMalloc(&c->d_send_elems, std::max(1, c->n_send) * sizeof(int)));
    CK(cudaMemcpy(c->d_send_elems, T->send_elems, c->n_send * sizeof(int), cudaMemcpyHostToDevice));
    {  // node masks of the send slots, from the DSS CSR: a local member is sent to rank q iff its node has a ghost member owned by q
      const int nn = T->n_neighbors, nh = c->dims.nh;
      std::vector<int> mask(std::max(1, c->n_send), 0);
      std::vector<std::vector<int>> slot_of(nn, std::vector<int>(nh, -1));
      for (int q = 0; q < nn; ++q)
        for (int k = c->send_off[q]; k < c->send_off[q + 1]; ++k) slot_of[q][T->send_elems[k]] = k;
      for (size_t nd = 0; nd + 1 < c->h_off.size(); ++nd) {
        unsigned qs = 0;  // neighbours owning a ghost member of this node
        for (int m = c->h_off[nd]; m < c->h_off[nd + 1]; ++m) {
          const int el = c->h_mem[m] >> 4;
          if (el >= nh)
            for (int q = 0; q < nn; ++q)
              if (el - nh >= c->recv_off[q] && el - nh < c->recv_off[q + 1]) qs |= 1u << q;
        }
        if (!qs) continue;
        for (int m = c->h_off[nd]; m < c->h_off[nd + 1]; ++m) {
          const int el = c->h_mem[m] >> 4, n = c->h_mem[m] & 15;
          if (el >= nh) continue;
          for (int q = 0; q < nn; ++q)
            if ((qs >> q & 1) && slot_of[q][el] >= 0) mask[slot_of[q][el]] |= 1 << n;
        }
      }
      CK(cudaMalloc(&c->d_slot_mask, mask.size() * sizeof(int)));
      CK(cudaMemcpy(c->d_slot_mask, mask.data(), mask.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    Id128 id;
    memcpy(&id, nccl_id, 128);
    NK(g_nccl.CommInitRank(&c->comm, nranks, id, rank));
  }
  return 0;
}


extern "C" int b200_create(b200_ctx** out, const b200_dims* d, const b200_geometry* G, const b200_topology* T,
                           const b200_params* p, const void* nccl_id, int rank, int nranks) {
  if (!out || !d || !G || !T || !p) return fail("b200_create: null argument");
  if (d->nq != 4) return fail("b200_create: only Nq = 4 (nh_poly = 3) is supported");
  if (d->nh <= 0 || d->nh_ghost < 0) return fail("b200_create: need nh > 0 and nh_ghost >= 0");
  if (T->n_neighbors < 0 || T->n_neighbors > 32) return fail("b200_create: 0 <= n_neighbors <= 32 (neighbour bit mask)");
  if (d->nv + 1 > LV || d->nv < 2) return fail("b200_create: need 2 <= nv <= 63");
  if (d->ft_bytes != 4 && d->ft_bytes != 8) return fail("b200_create: ft_bytes must be 4 or 8");
  if (d->n_tracers < 0 || d->n_tracers > 4) return fail("b200_create: 0 <= n_tracers <= 4");
  for (int u : {p->energy_upwinding, p->tracer_upwinding})
    if (u < 0 || u > 3) return fail("b200_create: upwinding must be 0 (none), 1 (first_order), 2 (third_order) or 3 (vanleer_limiter)");
  if (p->vert_diff < 0 || p->vert_diff > 2) return fail("b200_create: vert_diff must be 0 (none), 1 (VerticalDiffusion) or 2 (DecayWithHeightDiffusion)");
  if (p->implicit_diffusion && !p->vert_diff)
    return fail("b200_create: implicit_diffusion needs a vert_diff model (the reference's update_diffusion_jacobian! has no diffusivity otherwise)");
  if (p->vert_diff && p->approximate_linear_solve_iters < 0) return fail("b200_create: approximate_linear_solve_iters < 0");
  if (p->microphysics_0M) {
    if (p->microphysics_0M != 1) return fail("b200_create: microphysics_0M must be 0 (DryModel) or 1 (EquilibriumMicrophysics0M)");
    if (d->n_tracers < 1) return fail("b200_create: microphysics_0M needs n_tracers >= 1 (rho*q_tot is component 4 of Y.c)");
    if (p->vert_diff || p->held_suarez)
      return fail("b200_create: the moist (0M) state is built without vertical diffusion and without the Held-Suarez forcing");
    if (!(p->R_v > 0 && p->cp_v > p->R_v && p->cp_l > 0 && p->cp_i > 0 && p->T_triple > 0 && p->press_triple > 0 && p->T_freeze > p->T_icenuc))
      return fail("b200_create: microphysics_0M needs the Thermodynamics parameters (R_v, cp_v, cp_l, cp_i, LH_v0, LH_s0, T_triple, press_triple, T_freeze > T_icenuc)");
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("b200_create: no CUDA device (this library has no CPU fallback)");
  b200_ctx* c = new b200_ctx();
  const int rc = create_body(c, d, G, T, p, nccl_id, rank, nranks);
  if (rc != 0) {  // one cleanup path: everything allocated so far is released by b200_destroy
    b200_destroy(c);
    return rc;
  }
  *out = c;
  return 0;
}

extern "C" int b200_destroy(b200_ctx* c) {
  if (!c) return 0;
  auto fr = [](void* p) { if (p) cudaFree(p); };
  fr(c->d_hgeo); fr(c->d_vlev); fr(c->d_dssrec); fr(c->d_node_off); fr(c->d_lim_nbr_off); fr(c->d_lim_nbr); fr(c->d_lim_bnd); fr(c->d_lim_E); fr(c->d_lim_ghost_node);
  for (int i = 0; i < 4; ++i) fr(c->Tlc[i]);
  for (void* p : c->p2p_peer) if (p) cudaIpcCloseMemHandle(p);
  fr(c->p2p_buf); fr(c->d_p2p_seq); fr(c->d_slot_nbr); fr(c->d_slot_dst); fr(c->d_nbr_nhg); fr(c->d_nbr_rank); fr(c->d_p2p_dst); fr(c->d_p2p_flags); fr(c->d_off); fr(c->d_mem); fr(c->d_jac); fr(c->d_jsnap_c); fr(c->d_jsnap_f); fr(c->d_jacd); fr(c->d_kdec); fr(c->H); fr(c->Hw);
  for (int i = 0; i < 2; ++i) { fr(c->Uc[i]); fr(c->Uf[i]); }
  for (int i = 0; i < 4; ++i) { fr(c->Nsc[i]); fr(c->Nsf[i]); }
  for (int i = 0; i < 4; ++i) { fr(c->Tec[i]); fr(c->Tef[i]); fr(c->Tic[i]); fr(c->Tif[i]); }
  fr(c->Rc); fr(c->Rf); fr(c->dc); fr(c->df); fr(c->d_send_elems); fr(c->d_slot_mask); fr(c->sendbuf); fr(c->ghostbuf);
  for (auto& g : c->graphs) cudaGraphExecDestroy(g.exec);
  if (c->gstream) { cudaStreamDestroy(c->gstream); cudaEventDestroy(c->ev_gin); cudaEventDestroy(c->ev_gout); }
  if (c->side) { cudaStreamDestroy(c->side); cudaEventDestroy(c->ev_fork); cudaEventDestroy(c->ev_join); }
  if (c->comm) g_nccl.CommDestroy(c->comm);
  if (c->h_p2p_err) cudaFreeHost(c->h_p2p_err);
  delete c;
  return 0;
}

extern "C" int64_t b200_launch_count(b200_ctx* c) { return c ? c->launches : 0; }

// Kernel launch through cudaLaunchKernelEx, optionally as a programmatic dependent launch (PDL): the grid may start while the tail
// of the previous kernel of the stream is still running; every kernel launched this way executes griddepcontrol.wait before it touches
// global data (common.cuh: pdl_wait), so the stream order of memory effects is unchanged.
template <typename... P, typename... A>
static cudaError_t launchx(int pdl, void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, A&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<P>(args)...);
}

#define LAUNCH_CHECK(c)                                                               \
  do {                                                                                \
    (c)->launches++;                                                                  \
    cudaError_t e_ = cudaGetLastError();                                              \
    if (e_ != cudaSuccess) return fail(std::string("kernel launch: ") + cudaGetErrorString(e_)); \
  } while (0)

// ---------------------------------------------------------------------------------------------
template <class FT>
static int impl_cache_imp(b200_ctx* c, void* Yc, void* Yf, const b200_cacheptrs* o, cudaStream_t s) {
  b200_cacheptrs z = {};
  if (!o) o = &z;
  if (!o->u_c && !o->u3_f && !o->K_c && !o->T_c && !o->p_c && !o->h_tot_c) {
    // no p.precomputed field requested: all that is left of cache_imp! is the u₃ boundary filter (the tendency kernels recompute the
    // thermodynamics from Y themselves, DESIGN.md "Deviations")
    const int ncols = c->dims.nh * 16;
    launchx(c->pdl & 8, k_u3_filter<FT>, dim3((2 * ncols + 255) / 256), dim3(256), 0, s, (FT*)Yf, ncols, c->dims.nv + 1);
    LAUNCH_CHECK(c);
    return 0;
  }
  k_cache_imp<FT><<<c->dims.nh, NT, smem_slabs<FT>(1), s>>>(make_par<FT>(c), (const FT*)c->d_hgeo, (const VLev<FT>*)c->d_vlev,
                                                          (const FT*)Yc, (FT*)Yf, (FT*)o->u_c, (FT*)o->u3_f, (FT*)o->K_c,
                                                          (FT*)o->T_c, (FT*)o->p_c, (FT*)o->h_tot_c);
  LAUNCH_CHECK(c);
  return 0;
}
extern "C" int b200_cache_imp(b200_ctx* c, void* Yc, void* Yf, const b200_cacheptrs* o, void* stream) {
  CtxScope scope_(c);
  if (!c) return fail("b200_cache_imp: null context");
  return c->ft == 4 ? impl_cache_imp<float>(c, Yc, Yf, o, (cudaStream_t)stream) : impl_cache_imp<double>(c, Yc, Yf, o, (cudaStream_t)stream);
}

// Yₜ.c += vertical_diffusion_boundary_layer_tendency!(Y)  (kernels_vdiff.cuh)
template <class FT>
static int launch_vdiff_tend(b200_ctx* c, void* Ytc, const void* Yc, const void* Yf, cudaStream_t s) {
  // quarter element per CTA, no state slabs (HBM-bound; the element-slab first generation took 267 µs against 77 µs)
  k_vdiff_tend2<FT><<<c->dims.nh * 4, NT, VD2_ARR * 4 * VD2_ST * sizeof(FT), s>>>(make_par<FT>(c), make_vdiff<FT>(c), (const FT*)c->d_hgeo,
                                                                               (const VLev<FT>*)c->d_vlev, (const FT*)Yc,
                                                                               (const FT*)Yf, (FT*)Ytc);
  LAUNCH_CHECK(c);
  return 0;
}

template <class FT>
static int impl_t_imp(b200_ctx* c, void* Ytc, void* Ytf, const void* Yc, const void* Yf, cudaStream_t s) {
#define K8_HOOK(KERNEL, NV_, MOIST_)                                                                                                      \
  launchx(0, KERNEL<FT, NV_, MOIST_>, c->dims.nh, 256, 0, s, make_par<FT>(c), (const FT*)c->d_hgeo, (const VLev<FT>*)c->d_vlev, (const FT*)Yc, \
          (const FT*)Yf, (FT*)Ytc, (FT*)Ytf)
  const bool nv63h = c->dims.nv == 63 && !c->generic_nv;
  if (c->imp_kernel == 8) {  // warp per column pair (kernels_imp8.cuh)
    if (c->prm.microphysics_0M) { if (nv63h) K8_HOOK(k8_t_imp, 63, true); else K8_HOOK(k8_t_imp, 0, true); }
    else { if (nv63h) K8_HOOK(k8_t_imp, 63, false); else K8_HOOK(k8_t_imp, 0, false); }
  } else if (c->prm.microphysics_0M)
    k_t_imp2<FT, true><<<c->dims.nh * 4, NT, Q_WORDS * sizeof(FT), s>>>(make_par<FT>(c), (const FT*)c->d_hgeo, (const VLev<FT>*)c->d_vlev,
                                                                      (const FT*)Yc, (const FT*)Yf, (FT*)Ytc, (FT*)Ytf);
  else
    k_t_imp2<FT><<<c->dims.nh * 4, NT, Q_WORDS_DRY * sizeof(FT), s>>>(make_par<FT>(c), (const FT*)c->d_hgeo, (const VLev<FT>*)c->d_vlev,
                                                                (const FT*)Yc, (const FT*)Yf, (FT*)Ytc, (FT*)Ytf);
  LAUNCH_CHECK(c);
  if (vdiff_implicit(c)) return launch_vdiff_tend<FT>(c, Ytc, Yc, Yf, s);  // implicit_tendency.jl:69-78
  return 0;
}
extern "C" int b200_t_imp(b200_ctx* c, void* Ytc, void* Ytf, const void* Yc, const void* Yf, double, void* stream) {
  CtxScope scope_(c);
  if (!c) return fail("b200_t_imp: null context");
  return c->ft == 4 ? impl_t_imp<float>(c, Ytc, Ytf, Yc, Yf, (cudaStream_t)stream) : impl_t_imp<double>(c, Ytc, Ytf, Yc, Yf, (cudaStream_t)stream);
}

// Coefficient planes of the last Wfact (k_wfact2): what k_ldiv_diff consumes with implicit vertical diffusion, and what
// b200_debug_jacobian exports.
template <class FT>
static int launch_wfact_planes(b200_ctx* c, const void* Yc, const void* Yf, double dtg, cudaStream_t s) {
  if (!c->d_jac) CK(cudaMalloc(&c->d_jac, (size_t)c->dims.nh * JC_N * 16 * (c->dims.nv + 1) * sizeof(FT)));
  k_wfact2<FT><<<c->dims.nh * 4, NT, Q_WORDS * sizeof(FT), s>>>(make_par<FT>(c), (const FT*)c->d_hgeo, (const VLev<FT>*)c->d_vlev,
                                                              (const FT*)Yc, (const FT*)Yf, (FT)dtg, (FT*)c->d_jac);
  LAUNCH_CHECK(c);
  c->jac_planes_valid = true;
  return 0;
}
// Wfact (update_jacobian!).  Dry path: the Jacobian is a function of (Y, dtγ) only, so Wfact keeps a SNAPSHOT of Y (S bytes) and
// ldiv! recomputes the band coefficients from it in registers (k5_imp_stage<…, LDIV>), instead of streaming 15 coefficient planes
// (3·S) out here and back in there (profiles/r2_hook_path.md).  The snapshot (not the pointer) is kept because a stepper may update Y
// between Wfact and a later ldiv! (Jacobian reuse across Newton iterations).
template <class FT>
static int impl_wfact(b200_ctx* c, const void* Yc, const void* Yf, double dtg, cudaStream_t s) {
  c->jac_planes_valid = false;
  if (vdiff_implicit(c)) {  // update_diffusion_jacobian! (manual_sparse_jacobian.jl:1031-1261): planes for k_ldiv_diff
    if (launch_wfact_planes<FT>(c, Yc, Yf, dtg, s)) return -1;
    if (!c->d_jacd) CK(cudaMalloc(&c->d_jacd, (size_t)c->dims.nh * JD_N * 16 * (c->dims.nv + 1) * sizeof(FT)));
    k_vdiff_jac<FT><<<c->dims.nh, NT, smem_slabs<FT>(14), s>>>(make_par<FT>(c), make_vdiff<FT>(c), (const FT*)c->d_hgeo,
                                                             (const VLev<FT>*)c->d_vlev, (const FT*)Yc, (const FT*)Yf, (FT)dtg,
                                                             (FT*)c->d_jacd);
    LAUNCH_CHECK(c);
    return 0;
  }
  if (!c->d_jsnap_c) { CK(cudaMalloc(&c->d_jsnap_c, c->nc() * sizeof(FT))); CK(cudaMalloc(&c->d_jsnap_f, c->nf() * sizeof(FT))); }
  CK(cudaMemcpyAsync(c->d_jsnap_c, Yc, c->nc() * sizeof(FT), cudaMemcpyDeviceToDevice, s));
  CK(cudaMemcpyAsync(c->d_jsnap_f, Yf, c->nf() * sizeof(FT), cudaMemcpyDeviceToDevice, s));
  c->jsnap_dtg = dtg;
  return 0;
}
extern "C" int b200_wfact(b200_ctx* c, const void* Yc, const void* Yf, double dtg, double, void* stream) {
  CtxScope scope_(c);
  if (!c) return fail("b200_wfact: null context");
  return c->ft == 4 ? impl_wfact<float>(c, Yc, Yf, dtg, (cudaStream_t)stream) : impl_wfact<double>(c, Yc, Yf, dtg, (cudaStream_t)stream);
}

template <class FT>
static int impl_ldiv(b200_ctx* c, void* dYc, void* dYf, const void* Rc, const void* Rf, cudaStream_t s) {
  if (vdiff_implicit(c)) {  // ApproximateBlockArrowheadIterativeSolve (manual_sparse_jacobian.jl:538-578)
    if (!c->d_jac || !c->d_jacd) return fail("b200_ldiv: b200_wfact has not been called");
    // 16-lane Thomas sweeps (a parallel-cyclic-reduction version measured slower on B200: 15.2 vs 10.5 ms/step, profiles/r2_opt_in_validation.md)
    k_ldiv_diff<FT><<<c->dims.nh, NT, (22 * SLAB + LV) * sizeof(FT), s>>>(make_par<FT>(c), make_vdiff<FT>(c), (const VLev<FT>*)c->d_vlev,
                                                                 (const FT*)c->d_jac, (const FT*)c->d_jacd, (const FT*)Rc,
                                                                 (const FT*)Rf, (FT*)dYc, (FT*)dYf);
    LAUNCH_CHECK(c);
    return 0;
  }
  if (!c->d_jsnap_c) return fail("b200_ldiv: b200_wfact has not been called");
  // exact BlockArrowheadSolve (manual_sparse_jacobian.jl:579-584): Schur complement onto u₃, PCR, back-substitution — the coefficient
  // code of the fused implicit stage on the Wfact snapshot
  const bool nv63_ = c->dims.nv == 63 && !c->generic_nv;
#define K8_LDIV(NV_, MOIST_)                                                                                                                     \
  launchx(c->pdl & 16, k8_imp_stage<FT, NV_, true, MOIST_>, c->dims.nh, 256, 0, s, make_par<FT>(c), (const FT*)c->d_hgeo, (const VLev<FT>*)c->d_vlev, \
          (const FT*)c->d_jsnap_c, (const FT*)c->d_jsnap_f, (FT*)dYc, (FT*)dYf, (FT)c->jsnap_dtg, (const FT*)Rc, (const FT*)Rf)
  if (c->imp_kernel == 8) {  // warp per column pair (kernels_imp8.cuh)
    if (c->prm.microphysics_0M) { if (nv63_) K8_LDIV(63, true); else K8_LDIV(0, true); }
    else { if (nv63_) K8_LDIV(63, false); else K8_LDIV(0, false); }
  } else if (c->prm.microphysics_0M)
    launchx(c->pdl & 16, k5_imp_stage<FT, 0, true, true>, c->dims.nh, 256, smem_imp5<FT>(true), s, make_par<FT>(c), (const FT*)c->d_hgeo,
            (const VLev<FT>*)c->d_vlev, (const FT*)c->d_jsnap_c, (const FT*)c->d_jsnap_f, (FT*)dYc, (FT*)dYf, (FT)c->jsnap_dtg, (const FT*)Rc,
            (const FT*)Rf);
  else if (c->dims.nv == 63 && !c->generic_nv)
    launchx(c->pdl & 16, k5_imp_stage<FT, 63, true>, c->dims.nh, 256, smem_imp5<FT>(), s, make_par<FT>(c), (const FT*)c->d_hgeo,
            (const VLev<FT>*)c->d_vlev, (const FT*)c->d_jsnap_c, (const FT*)c->d_jsnap_f, (FT*)dYc, (FT*)dYf, (FT)c->jsnap_dtg, (const FT*)Rc,
            (const FT*)Rf);
  else
    launchx(c->pdl & 16, k5_imp_stage<FT, 0, true>, c->dims.nh, 256, smem_imp5<FT>(), s, make_par<FT>(c), (const FT*)c->d_hgeo,
            (const VLev<FT>*)c->d_vlev, (const FT*)c->d_jsnap_c, (const FT*)c->d_jsnap_f, (FT*)dYc, (FT*)dYf, (FT)c->jsnap_dtg, (const FT*)Rc,
            (const FT*)Rf);
  LAUNCH_CHECK(c);
  return 0;
}
// Debug/test aid: copy the Jacobian coefficient planes stored by the last b200_wfact — [nh][15][16][nv+1] values of FT in the order of
// the JC_* enum (kernels_implicit.cuh): Schur tridiagonal l, d, u of the u₃ rows; (u₃,ρ), (u₃,ρe_tot), (u₃,uₕ₁), (u₃,uₕ₂) bidiagonals
// (lo, hi = centres f−1, f); (ρ,u₃), (ρe_tot,u₃) bidiagonals (lo, hi = faces k, k+1) — into a caller-owned DEVICE buffer.
extern "C" int b200_debug_jacobian(b200_ctx* c, void* dst, int64_t capacity_bytes, void* stream) {
  CtxScope scope_(c);
  if (!c) return fail("b200_debug_jacobian: null context");
  if (c->prm.microphysics_0M) return fail("b200_debug_jacobian: the coefficient planes are a dry-path debug aid (moist contexts solve from the Wfact snapshot)");
  if (!c->jac_planes_valid) {  // dry path: Wfact kept a snapshot; materialise the planes from it now
    if (!c->d_jsnap_c) return fail("b200_debug_jacobian: b200_wfact has not been called");
    const int rc = c->ft == 4 ? launch_wfact_planes<float>(c, c->d_jsnap_c, c->d_jsnap_f, c->jsnap_dtg, (cudaStream_t)stream)
                              : launch_wfact_planes<double>(c, c->d_jsnap_c, c->d_jsnap_f, c->jsnap_dtg, (cudaStream_t)stream);
    if (rc) return rc;
  }
  const size_t bytes = (size_t)c->dims.nh * JC_N * 16 * (c->dims.nv + 1) * (size_t)c->ft;
  if ((size_t)capacity_bytes < bytes) return fail("b200_debug_jacobian: destination too small");
  CK(cudaMemcpyAsync(dst, c->d_jac, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}
extern "C" int b200_ldiv(b200_ctx* c, void* dYc, void* dYf, const void* Rc, const void* Rf, void* stream) {
  CtxScope scope_(c);
  if (!c) return fail("b200_ldiv: null context");
  return c->ft == 4 ? impl_ldiv<float>(c, dYc, dYf, Rc, Rf, (cudaStream_t)stream) : impl_ldiv<double>(c, dYc, dYf, Rc, Rf, (cudaStream_t)stream);
}

template <class FT>
static int impl_t_post(b200_ctx* c, void* Ytc, void* Ytf, const void* Yc, const void* Yf, cudaStream_t s) {
  if (c->prm.energy_upwinding == 0) {  // vtt(:none) − vtt(:none) ≡ 0 (the reference does not even wire the hook, integrator.jl:212-214)
    CK(cudaMemsetAsync(Ytc, 0, c->nc() * sizeof(FT), s));
    CK(cudaMemsetAsync(Ytf, 0, c->nf() * sizeof(FT), s));
    return 0;
  }
  const bool nv63h = c->dims.nv == 63 && !c->generic_nv;
  if (c->imp_kernel == 8) {
    if (c->prm.microphysics_0M) { if (nv63h) K8_HOOK(k8_t_post_imp, 63, true); else K8_HOOK(k8_t_post_imp, 0, true); }
    else { if (nv63h) K8_HOOK(k8_t_post_imp, 63, false); else K8_HOOK(k8_t_post_imp, 0, false); }
  } else if (c->prm.microphysics_0M)
    k_t_post_imp2<FT, true><<<c->dims.nh * 4, NT, Q_WORDS * sizeof(FT), s>>>(make_par<FT>(c), (const FT*)c->d_hgeo, (const VLev<FT>*)c->d_vlev,
                                                                           (const FT*)Yc, (const FT*)Yf, (FT*)Ytc, (FT*)Ytf);
  else
    k_t_post_imp2<FT><<<c->dims.nh * 4, NT, Q_WORDS_DRY * sizeof(FT), s>>>(make_par<FT>(c), (const FT*)c->d_hgeo, (const VLev<FT>*)c->d_vlev,
                                                                     (const FT*)Yc, (const FT*)Yf, (FT*)Ytc, (FT*)Ytf);
  LAUNCH_CHECK(c);
  return 0;
}
extern "C" int b200_t_post_imp(b200_ctx* c, void* Ytc, void* Ytf, const void* Yc, const void* Yf, double, void* stream) {
  CtxScope scope_(c);
  if (!c) return fail("b200_t_post_imp: null context");
  return c->ft == 4 ? impl_t_post<float>(c, Ytc, Ytf, Yc, Yf, (cudaStream_t)stream) : impl_t_post<double>(c, Ytc, Ytf, Yc, Yf, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// Peer-memory halo set-up
static size_t p2p_state_slab(const b200_ctx* c) { return (size_t)(c->ncf() * 16 * c->dims.nv + 16 * (c->dims.nv + 1)); }
extern "C" int b200_halo_export(b200_ctx* c, void* handle64_out) {
  CtxScope scope_(c);
  if (!c) return fail("b200_halo_export: null context");
  if (c->nbr.empty()) return fail("b200_halo_export: context has no neighbours");
  if (!c->p2p_buf) {
    c->p2p_cap = p2p_state_slab(c) * (size_t)std::max(1, (int)c->dims.nh_ghost) * c->ft;
    c->p2p_cap = (c->p2p_cap + 255) / 256 * 256;
    CK(cudaMalloc(&c->p2p_buf, 2 * c->p2p_cap + sizeof(int) * (size_t)c->nranks));
    CK(cudaMemset(c->p2p_buf, 0, 2 * c->p2p_cap + sizeof(int) * (size_t)c->nranks));
  }
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, c->p2p_buf));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64_out, &h, 64);
  return 0;
}
extern "C" int b200_halo_import(b200_ctx* c, const void* handles, const int32_t* their_recv_offset, const int32_t* their_nh_ghost) {
  CtxScope scope_(c);
  if (!c) return fail("b200_halo_import: null context");
  if (!c->p2p_buf) return fail("b200_halo_import: call b200_halo_export first");
  const int nn = (int)c->nbr.size();
  std::vector<int*> flags(nn);
  c->p2p_peer.resize(nn);
  for (int q = 0; q < nn; ++q) {
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + 64 * (size_t)q, 64);
    CK(cudaIpcOpenMemHandle(&c->p2p_peer[q], h, cudaIpcMemLazyEnablePeerAccess));
  }
  // per send slot: which neighbour and which ghost slot over there
  std::vector<int> slot_nbr(c->n_send), slot_dst(c->n_send), nhg(nn), ranks(nn);
  for (int q = 0; q < nn; ++q) {
    nhg[q] = their_nh_ghost[q]; ranks[q] = c->nbr[q];
    for (int k = c->send_off[q]; k < c->send_off[q + 1]; ++k) { slot_nbr[k] = q; slot_dst[k] = their_recv_offset[q] + (k - c->send_off[q]); }
    // my flag lives at index `rank` of neighbour q's flag array, which sits after its two parity blocks
    size_t their_cap = (p2p_state_slab(c) * (size_t)std::max(1, nhg[q]) * c->ft + 255) / 256 * 256;
    flags[q] = reinterpret_cast<int*>((char*)c->p2p_peer[q] + 2 * their_cap) + c->rank;
  }
  CK(cudaMalloc(&c->d_slot_nbr, std::max(1, c->n_send) * sizeof(int)));
  CK(cudaMalloc(&c->d_slot_dst, std::max(1, c->n_send) * sizeof(int)));
  CK(cudaMalloc(&c->d_nbr_nhg, nn * sizeof(int)));
  CK(cudaMalloc(&c->d_nbr_rank, nn * sizeof(int)));
  CK(cudaMalloc(&c->d_p2p_dst, 2 * nn * sizeof(void*)));
  CK(cudaMalloc(&c->d_p2p_flags, nn * sizeof(int*)));
  CK(cudaMemcpy(c->d_slot_nbr, slot_nbr.data(), c->n_send * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_slot_dst, slot_dst.data(), c->n_send * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_nbr_nhg, nhg.data(), nn * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_nbr_rank, ranks.data(), nn * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_p2p_flags, flags.data(), nn * sizeof(int*), cudaMemcpyHostToDevice));
  // destination base pointers for parity 0 and parity 1
  std::vector<void*> dst(2 * nn);
  for (int q = 0; q < nn; ++q) {
    size_t their_cap = (p2p_state_slab(c) * (size_t)std::max(1, nhg[q]) * c->ft + 255) / 256 * 256;
    dst[q] = c->p2p_peer[q];
    dst[nn + q] = (char*)c->p2p_peer[q] + their_cap;
  }
  CK(cudaMemcpy(c->d_p2p_dst, dst.data(), 2 * nn * sizeof(void*), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&c->d_p2p_seq, 2 * sizeof(int)));
  CK(cudaMemset(c->d_p2p_seq, 0, 2 * sizeof(int)));
  {  // error word of the halo waits: host-mapped, so a timeout is visible to the host without a device synchronisation
    if (!c->h_p2p_err) CK(cudaHostAlloc((void**)&c->h_p2p_err, sizeof(int), cudaHostAllocMapped));
    *c->h_p2p_err = 0;
    int* dptr = nullptr;
    CK(cudaHostGetDevicePointer((void**)&dptr, c->h_p2p_err, 0));
    CK(cudaMemcpyToSymbol(g_p2p_err, &dptr, sizeof(dptr)));
    if (const char* e = getenv("B200_P2P_SPIN_LIMIT")) { long long lim = atoll(e); CK(cudaMemcpyToSymbol(g_p2p_spin, &lim, sizeof(lim))); }
  }
  c->p2p_ready = true;
  return 0;
}

static P2PSig p2p_sig(const b200_ctx* c);
static P2PPlan p2p_plan(const b200_ctx* c) {
  P2PPlan Q;
  Q.send_elems = c->d_send_elems; Q.slot_nbr = c->d_slot_nbr; Q.slot_dst = c->d_slot_dst; Q.slot_mask = c->d_slot_mask;
  Q.dst = c->d_p2p_dst; Q.nbr_nh_ghost = c->d_nbr_nhg; Q.S = p2p_sig(c); Q.nblocks = c->n_send;
  return Q;
}
static P2PSig p2p_sig(const b200_ctx* c) { return P2PSig{c->d_p2p_flags, c->d_p2p_seq, c->d_p2p_seq + 1, (int)c->nbr.size()}; }
static P2PWait p2p_wait_args(const b200_ctx* c, bool on) {
  if (!on) return P2PWait{nullptr, nullptr, nullptr, 0};
  return P2PWait{reinterpret_cast<const int*>((char*)c->p2p_buf + 2 * c->p2p_cap), c->d_nbr_rank, c->d_p2p_seq, (int)c->nbr.size()};
}
// ---------------------------------------------------------------------------------------------
// DSS: local gather–scatter, with a whole-slab halo exchange for elements owned by other ranks.
struct DssField { void* ptr; int nf; int is_face; int kind; };

template <class FT>
static int impl_dss(b200_ctx* c, const DssField* F, int nfields, cudaStream_t s) {
  const int nv = c->dims.nv, nh = c->dims.nh;
  DssArgs A;
  A.n = 0;
  // halo: exchange the raw slabs of the send elements; ghosts land in ghostbuf per field
  size_t tot_slab = 0;
  for (int k = 0; k < nfields; ++k) tot_slab += (size_t)F[k].nf * 16 * (nv + F[k].is_face);
  const bool halo = c->comm != nullptr;
  const bool p2p = halo && c->p2p_ready && nfields <= 4 && tot_slab <= p2p_state_slab(c) && !getenv("B200_HALO_NCCL");
  void* ghost_base = nullptr;
  P2PArgs PAp; PAp.nfields = 0;
  bool have_pack = false;  // this exchange has slabs to send (packed inside the ghost-free DSS launch, or by k_pack_p2p)
  if (p2p) {
    // peer-memory halo: pack straight into the neighbours' ghost buffers, raise flags, wait for theirs
    const int nn = (int)c->nbr.size();
    P2PArgs PA;
    PA.nfields = nfields;
    long long goff = 0;
    for (int f = 0; f < nfields; ++f) {
      int slab = F[f].nf * 16 * (nv + F[f].is_face);
      PA.f[f] = {F[f].ptr, slab, goff, F[f].nf, nv + F[f].is_face};
      goff += slab;
    }
    PAp = PA; have_pack = c->n_send > 0;
    (void)nn;
    ghost_base = c->p2p_buf;  // parity block 0; the kernels add (*seq & 1)·p2p_cap
  } else if (halo) {
    size_t need_s = tot_slab * c->n_send * sizeof(FT), need_g = tot_slab * c->dims.nh_ghost * sizeof(FT);
    if (need_s + need_g > c->halo_cap) {
      if (c->sendbuf) cudaFree(c->sendbuf);
      if (c->ghostbuf) cudaFree(c->ghostbuf);
      CK(cudaMalloc(&c->sendbuf, std::max<size_t>(need_s, 16)));
      CK(cudaMalloc(&c->ghostbuf, std::max<size_t>(need_g, 16)));
      c->halo_cap = need_s + need_g;
    }
    size_t so = 0, go = 0;
    for (int k = 0; k < nfields; ++k) {
      int slab = F[k].nf * 16 * (nv + F[k].is_face);
      if (c->n_send > 0) {
        k_pack<FT><<<c->n_send, 256, 0, s>>>((const FT*)F[k].ptr, (FT*)c->sendbuf + so, c->d_send_elems, slab);
        LAUNCH_CHECK(c);
      }
      so += (size_t)slab * c->n_send; go += (size_t)slab * c->dims.nh_ghost;
    }
    NK(g_nccl.GroupStart());
    so = 0; go = 0;
    const int dt_nccl = sizeof(FT) == 4 ? 7 /*ncclFloat32*/ : 8 /*ncclFloat64*/;
    for (int k = 0; k < nfields; ++k) {
      int slab = F[k].nf * 16 * (nv + F[k].is_face);
      for (size_t q = 0; q < c->nbr.size(); ++q) {
        int ns = c->send_off[q + 1] - c->send_off[q], nr = c->recv_off[q + 1] - c->recv_off[q];
        if (ns > 0) NK(g_nccl.Send((FT*)c->sendbuf + so + (size_t)c->send_off[q] * slab, (size_t)ns * slab, dt_nccl, c->nbr[q], c->comm, s));
        if (nr > 0) NK(g_nccl.Recv((FT*)c->ghostbuf + go + (size_t)c->recv_off[q] * slab, (size_t)nr * slab, dt_nccl, c->nbr[q], c->comm, s));
      }
      so += (size_t)slab * c->n_send; go += (size_t)slab * c->dims.nh_ghost;
    }
    NK(g_nccl.GroupEnd());
    ghost_base = c->ghostbuf;
  }
  size_t go = 0;
  for (int k = 0; k < nfields; ++k) {
    const int nlev = nv + F[k].is_face;
    const int estride = F[k].nf * 16 * nlev;
    FT* base = (FT*)F[k].ptr;
    FT* gbase = halo ? (FT*)ghost_base + go : nullptr;
    go += (size_t)estride * c->dims.nh_ghost;
    int comp = 0;
    auto add = [&](bool pair) -> int {
      if (A.n >= DSS_MAX_ITEMS) return fail("b200_dss: too many components in one call (max 8 items)");
      DssItem& I = A.it[A.n++];
      I.p0 = base + (size_t)comp * 16 * nlev;
      I.p1 = pair ? base + (size_t)(comp + 1) * 16 * nlev : nullptr;
      I.g0 = gbase ? gbase + (size_t)comp * 16 * nlev : nullptr;
      I.g1 = (gbase && pair) ? gbase + (size_t)(comp + 1) * 16 * nlev : nullptr;
      I.nlev = nlev; I.estride = estride; I.gstride = estride;
      comp += pair ? 2 : 1;
      return 0;
    };
    if (F[k].kind == 2) { if (add(false)) return -1; if (add(true)) return -1; }
    else if (F[k].kind == 1) { if (add(true)) return -1; }
    while (comp < F[k].nf) if (add(false)) return -1;
  }
  dim3 blk(64, 4), grd((c->nnodes + 3) / 4, 1);
  const DssNode<FT>* rec = (const DssNode<FT>*)c->d_dssrec;
  int pairs = 0;
  for (int k = 0; k < A.n; ++k) pairs |= (A.it[k].p1 != nullptr) << k;
  A.seq = p2p ? c->d_p2p_seq : nullptr; A.gpar = (long long)c->p2p_cap;
  const bool small = (size_t)(nh + c->dims.nh_ghost) * 64 * (size_t)(nv + 1) < (size_t)INT32_MAX;  // 32-bit offsets
  // With the peer-memory halo the nodes without ghost members are summed while the neighbours' slabs are still in flight; the
  // flag wait and the (few) ghost-touching nodes follow.  k_dss2<…, HALO> takes the record range [node0, node1).
  int node0 = 0, node1 = c->nnodes;
  bool with_halo = halo, with_pack = false;
  auto pack_alone = [&]() -> int {  // pack + signal as their own launch (generic paths)
    if (have_pack) launchx(c->pdl & 4, k_pack_p2p<FT>, dim3(c->n_send), dim3(256), 0, s, PAp, p2p_plan(c));
    else k_p2p_signal<<<1, 32, 0, s>>>(c->d_p2p_flags, (int)c->nbr.size(), c->d_p2p_seq);
    LAUNCH_CHECK(c);
    return 0;
  };
  auto wait_p2p = [&]() -> int {
    k_p2p_wait<<<1, 32, 0, s>>>(reinterpret_cast<const int*>((char*)c->p2p_buf + 2 * c->p2p_cap), c->d_nbr_rank, (int)c->nbr.size(), c->d_p2p_seq);
    LAUNCH_CHECK(c);
    return 0;
  };
#define DSS2(NI, PM)                                                                                                       \
  (with_halo ? (void)launchx(c->pdl & 4, k_dss2<FT, NI, PM, true, false>, grd, blk, 0, s, A, rec, node0, node1, nh, p2p_wait_args(c, p2p), P2PArgs(), P2PPlan()) \
   : with_pack ? (void)launchx(c->pdl & 4, k_dss2<FT, NI, PM, false, true>, grd, blk, 0, s, A, rec, node0, node1, nh, p2p_wait_args(c, false), PAp, p2p_plan(c)) \
               : (void)launchx(c->pdl & 4, k_dss2<FT, NI, PM, false, false>, grd, blk, 0, s, A, rec, node0, node1, nh, p2p_wait_args(c, false), P2PArgs(), P2PPlan()))
#define DSS2_ANY()                                                                                                         \
  do {                                                                                                                     \
    if (A.n == 4 && pairs == 0x2) DSS2(4, 0x2);        /* state: ρ, (uₕ₁,uₕ₂), ρe_tot, u₃ */                                 \
    else if (A.n == 3 && pairs == 0x1) DSS2(3, 0x1);   /* ∇² fields: (∇²u₁,∇²u₂), ∇²u₃, ∇²s_d */                             \
    else if (A.n == 5 && pairs == 0x2) DSS2(5, 0x2);   /* state with one passive tracer */                                 \
    else if (A.n == 4 && pairs == 0x1) DSS2(4, 0x1);   /* ∇² fields with one passive tracer */                             \
    else if (A.n == 6 && pairs == 0x2) DSS2(6, 0x2);   /* two tracers */                                                   \
    else if (A.n == 5 && pairs == 0x1) DSS2(5, 0x1);                                                                       \
    else if (A.n == 1 && pairs == 0x0) DSS2(1, 0x0);                                                                       \
    else if (A.n == 1 && pairs == 0x1) DSS2(1, 0x1);                                                                       \
    else DSS2(2, 0x0);                                                                                                     \
    LAUNCH_CHECK(c);                                                                                                       \
  } while (0)
  const bool special = (A.n == 4 && pairs == 0x2) || (A.n == 3 && pairs == 0x1) || (A.n == 5 && pairs == 0x2) || (A.n == 4 && pairs == 0x1) ||
                       (A.n == 6 && pairs == 0x2) || (A.n == 5 && pairs == 0x1) || (A.n == 1 && pairs == 0x0) || (A.n == 1 && pairs == 0x1) ||
                       (A.n == 2 && pairs == 0x0);
  if (!small || !special) {
    if (p2p && (pack_alone() || wait_p2p())) return -1;
    grd.y = A.n;
    k_dss<FT><<<grd, blk, 0, s>>>(A, c->d_off, c->d_mem, (const FT*)c->d_hgeo, c->nnodes, nh);
    LAUNCH_CHECK(c);
  } else if (p2p) {
    node0 = 0; node1 = c->n_int_nodes; with_halo = false;
    if (have_pack) {  // the first n_send blocks of this launch pack and signal, the others sum the ghost-free nodes
      with_pack = true;
      grd.x = c->n_send + (node1 - node0 + 3) / 4; DSS2_ANY();
      with_pack = false;
    } else {
      if (pack_alone()) return -1;
      if (node1 > node0) { grd.x = (node1 - node0 + 3) / 4; DSS2_ANY(); }
    }
    node0 = c->n_int_nodes; node1 = c->nnodes; with_halo = true;  // these blocks poll the neighbours' flags themselves
    if (node1 > node0) { grd.x = (node1 - node0 + 3) / 4; DSS2_ANY(); }
    else if (wait_p2p()) return -1;  // keep the exchange protocol in step even without ghost-touching nodes
  } else {
    DSS2_ANY();
  }
#undef DSS2_ANY
#undef DSS2
  return 0;
}
extern "C" int b200_dss(b200_ctx* c, void* const* fields, const int32_t* nf, const int32_t* is_face, const int32_t* kind,
                        int32_t nfields, void* stream) {
  CtxScope scope_(c);
  if (!c) return fail("b200_dss: null context");
  if (halo_failed(c, "b200_dss")) return -1;
  if (nfields > 8) return fail("b200_dss: at most 8 fields per call");
  DssField F[8];
  for (int k = 0; k < nfields; ++k) F[k] = {fields[k], nf[k], is_face[k], kind[k]};
  return c->ft == 4 ? impl_dss<float>(c, F, nfields, (cudaStream_t)stream) : impl_dss<double>(c, F, nfields, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
template <class FT>
static int launch_axpy(b200_ctx* c, FT* out, const FT* base, int n, const FT* const* T, const double* coef, size_t N, cudaStream_t s,
                       int nlev = 0, unsigned dmask = 0) {
  AxpyArgs<FT> A;
  A.n = 0; A.dmask = 0;
  for (int k = 0; k < n; ++k) {
    if (coef[k] == 0.0) continue;
    if (A.n >= AXPY_MAX) return fail("b200_axpy_n: too many terms");
    if (dmask >> k & 1) A.dmask |= 1u << A.n;
    A.T[A.n] = T[k]; A.c[A.n] = (FT)coef[k]; A.n++;
  }
  bool al = (N % 4 == 0) && (((uintptr_t)out | (uintptr_t)base) % (4 * sizeof(FT)) == 0);
  for (int k = 0; k < A.n; ++k) al = al && ((uintptr_t)A.T[k] % (4 * sizeof(FT)) == 0);
  int blocks = 148 * 8;
  if (al) launchx(c->pdl & 8, k_axpy_n<FT, 4>, blocks, 256, 0, s, out, base, A, N / 4, nlev);
  else launchx(c->pdl & 8, k_axpy_n<FT, 1>, blocks, 256, 0, s, out, base, A, N, nlev);
  LAUNCH_CHECK(c);
  return 0;
}
template <class FT>
static int impl_axpy(b200_ctx* c, void* Uc, void* Uf, const void* uc, const void* uf, int n, const void* const* Tc,
                     const void* const* Tf, const double* coef, cudaStream_t s, bool filter_u3 = false, unsigned dmask = 0) {
  if (launch_axpy<FT>(c, (FT*)Uc, (const FT*)uc, n, (const FT* const*)Tc, coef, c->nc(), s, 0, dmask)) return -1;
  return launch_axpy<FT>(c, (FT*)Uf, (const FT*)uf, n, (const FT* const*)Tf, coef, c->nf(), s, filter_u3 ? c->dims.nv + 1 : 0, dmask);
}
extern "C" int b200_axpy_n(b200_ctx* c, void* Uc, void* Uf, const void* uc, const void* uf, int32_t n, const void* const* Tc,
                           const void* const* Tf, const double* coef, void* stream) {
  CtxScope scope_(c);
  if (!c) return fail("b200_axpy_n: null context");
  return c->ft == 4 ? impl_axpy<float>(c, Uc, Uf, uc, uf, n, Tc, Tf, coef, (cudaStream_t)stream)
                    : impl_axpy<double>(c, Uc, Uf, uc, uf, n, Tc, Tf, coef, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
template <class FT>
static int impl_t_exp_phase(b200_ctx* c, int phase, void* Ytc, void* Ytf, const void* Yc, const void* Yf, cudaStream_t s,
                            void* Ylc = nullptr) {
  const bool hd = c->prm.hyperdiff != 0;
  if (hd && !c->H) CK(cudaMalloc(&c->H, c->nc() * sizeof(FT)));
  const bool nv63 = c->dims.nv == 63 && !c->generic_nv;
  const bool moist = c->prm.microphysics_0M != 0;
  const int n_passive = c->dims.n_tracers - (moist ? 1 : 0);
  if (moist && hd && !c->Hw) CK(cudaMalloc(&c->Hw, (size_t)c->dims.nh * 16 * c->dims.nv * sizeof(FT)));
  if (phase == 0) {
    if (moist && nv63)  // moist thermodynamic state + ∇²q_tot_eff → H[4], ρ(h_eff + Φ) → Hw
      launchx(c->pdl & 1, k5_exp_a<FT, 63, true>, c->dims.nh, CT, smem_rowq<FT>(9), s, make_par<FT>(c), (const FT*)c->d_hgeo, (const VLev<FT>*)c->d_vlev,
              (const FT*)Yc, (const FT*)Yf, (FT*)Ytc, (FT*)Ytf, hd ? (FT*)c->H : nullptr, (FT*)c->Hw, (FT*)Ylc);
    else if (moist)
      launchx(c->pdl & 1, k5_exp_a<FT, 0, true>, c->dims.nh, CT, smem_rowq<FT>(9), s, make_par<FT>(c), (const FT*)c->d_hgeo, (const VLev<FT>*)c->d_vlev,
              (const FT*)Yc, (const FT*)Yf, (FT*)Ytc, (FT*)Ytf, hd ? (FT*)c->H : nullptr, (FT*)c->Hw, (FT*)Ylc);
    else if (nv63)
      launchx(c->pdl & 1, k5_exp_a<FT, 63>, c->dims.nh, CT, smem_rowq<FT>(9), s, make_par<FT>(c), (const FT*)c->d_hgeo, (const VLev<FT>*)c->d_vlev, (const FT*)Yc,
                                                             (const FT*)Yf, (FT*)Ytc, (FT*)Ytf, hd ? (FT*)c->H : nullptr, (FT*)nullptr, (FT*)nullptr);
    else
      launchx(c->pdl & 1, k5_exp_a<FT, 0>, c->dims.nh, CT, smem_rowq<FT>(9), s, make_par<FT>(c), (const FT*)c->d_hgeo, (const VLev<FT>*)c->d_vlev, (const FT*)Yc,
                                                            (const FT*)Yf, (FT*)Ytc, (FT*)Ytf, hd ? (FT*)c->H : nullptr, (FT*)nullptr, (FT*)nullptr);
    LAUNCH_CHECK(c);
    if (n_passive > 0) {
      k5_tracer_a<FT><<<dim3(c->dims.nh, n_passive), CT, smem_row<FT>(3), s>>>(
          make_par<FT>(c), (const FT*)c->d_hgeo, (const VLev<FT>*)c->d_vlev, (const FT*)Yc, (const FT*)Yf, (FT*)Ytc, (FT*)Ylc,
          hd ? (FT*)c->H : nullptr);
      LAUNCH_CHECK(c);
    }
  } else if (phase == 1 && hd) {
    DssField F = {c->H, c->ncf(), 0, 1};  // (∇²u₁, ∇²u₂) pair, ∇²u₃, ∇²s_d, ∇²χ…
    if (impl_dss<FT>(c, &F, 1, s)) return -1;
  } else if (phase == 2 && hd) {
    const dim3 g7((c->dims.nh + LVL_EPB - 1) / LVL_EPB, 3 + n_passive);  // parts 3..: hyperdiffusion of the passive tracers
    const size_t sm7 = LVL_EPB * 16 * sizeof(FT);
    if (nv63)
      launchx(c->pdl & 2, k7_exp_c<FT, 63>, g7, LVL_EPB * 64, sm7, s, make_par<FT>(c), (const FT*)c->d_hgeo, (const VLev<FT>*)c->d_vlev,
              (const FT*)Yc, (const FT*)c->H, (FT*)Ytc, (FT*)Ytf, (FT*)(Ylc ? Ylc : Ytc), (const FT*)c->Hw);
    else
      launchx(c->pdl & 2, k7_exp_c<FT, 0>, g7, LVL_EPB * 64, sm7, s, make_par<FT>(c), (const FT*)c->d_hgeo, (const VLev<FT>*)c->d_vlev,
              (const FT*)Yc, (const FT*)c->H, (FT*)Ytc, (FT*)Ytf, (FT*)(Ylc ? Ylc : Ytc), (const FT*)c->Hw);
    LAUNCH_CHECK(c);
  }
  return 0;
}
extern "C" int b200_t_exp_phase(b200_ctx* c, int32_t phase, void* Ytc, void* Ytf, const void* Yc, const void* Yf, void* stream) {
  CtxScope scope_(c);
  if (!c) return fail("b200_t_exp_phase: null context");
  if (halo_failed(c, "b200_t_exp_phase")) return -1;
  return c->ft == 4 ? impl_t_exp_phase<float>(c, phase, Ytc, Ytf, Yc, Yf, (cudaStream_t)stream)
                    : impl_t_exp_phase<double>(c, phase, Ytc, Ytf, Yc, Yf, (cudaStream_t)stream);
}

template <class FT>
static int impl_t_exp(b200_ctx* c, void* Ytc, void* Ytf, void* Ylc, void* Ylf, const void* Yc, const void* Yf, cudaStream_t s) {
  if (Ylc) CK(cudaMemsetAsync(Ylc, 0, c->nc() * sizeof(FT), s));
  if (Ylf) CK(cudaMemsetAsync(Ylf, 0, c->nf() * sizeof(FT), s));
  for (int ph = 0; ph < 3; ++ph)
    if (impl_t_exp_phase<FT>(c, ph, Ytc, Ytf, Yc, Yf, s, Ylc)) return -1;
  if (vdiff_explicit(c)) return launch_vdiff_tend<FT>(c, Ytc, Yc, Yf, s);  // additional_tendency! (remaining_tendency.jl:185-195)
  return 0;
}
extern "C" int b200_t_exp_lim(b200_ctx* c, void* Ytc, void* Ytf, void* Ylc, void* Ylf, const void* Yc, const void* Yf, double,
                              void* stream) {
  CtxScope scope_(c);
  if (!c) return fail("b200_t_exp_lim: null context");
  if (halo_failed(c, "b200_t_exp_lim")) return -1;
  return c->ft == 4 ? impl_t_exp<float>(c, Ytc, Ytf, Ylc, Ylf, Yc, Yf, (cudaStream_t)stream)
                    : impl_t_exp<double>(c, Ytc, Ytf, Ylc, Ylf, Yc, Yf, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// ARS343 tableau (Ascher–Ruuth–Spiteri 1997 §2.7; ClimaTimeSteppers IMEXAlgorithm(ARS343))
struct Tableau { double ae[4][4], ai[4][4], be[4], bi[4]; };
static Tableau ars343() {
  Tableau t;
  memset(&t, 0, sizeof(t));
  const double g = 0.4358665215084590, a42 = 0.5529291480359398, a43 = a42;
  const double b1 = -3 * g * g / 2 + 4 * g - 0.25, b2 = 3 * g * g / 2 - 5 * g + 1.25;
  const double a31 = (1 - 9 * g / 2 + 3 * g * g / 2) * a42 + (11.0 / 4 - 21 * g / 2 + 15 * g * g / 4) * a43 - 3.5 + 13 * g - 9 * g * g / 2;
  const double a32 = (-1 + 9 * g / 2 - 3 * g * g / 2) * a42 + (-11.0 / 4 + 21 * g / 2 - 15 * g * g / 4) * a43 + 4 - 25 * g / 2 + 9 * g * g / 2;
  t.ae[1][0] = g; t.ae[2][0] = a31; t.ae[2][1] = a32; t.ae[3][0] = 1 - a42 - a43; t.ae[3][1] = a42; t.ae[3][2] = a43;
  t.ai[1][1] = g; t.ai[2][1] = (1 - g) / 2; t.ai[2][2] = g; t.ai[3][1] = b1; t.ai[3][2] = b2; t.ai[3][3] = g;
  t.be[1] = b1; t.be[2] = b2; t.be[3] = g;
  t.bi[1] = b1; t.bi[2] = b2; t.bi[3] = g;
  return t;
}

template <class FT>
static int launch_diff_scale(b200_ctx* c, FT* out, const FT* a, const FT* b, FT sc, size_t N, cudaStream_t s) {
  bool al = (N % 4 == 0) && (((uintptr_t)out | (uintptr_t)a | (uintptr_t)b) % (4 * sizeof(FT)) == 0);
  if (al) launchx(c->pdl & 32, k_diff_scale<FT, 4>, 148 * 8, 256, 0, s, out, a, b, sc, N / 4);
  else launchx(c->pdl & 32, k_diff_scale<FT, 1>, 148 * 8, 256, 0, s, out, a, b, sc, N);
  LAUNCH_CHECK(c);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// lim!(Y, p, t, ref_Y) (limited_tendencies.jl:64-122): SEM quasi-monotone limiter on every tracer of Y.c, bounds from ref_Y.
// The reference's no-op when no limiter is configured.
template <class FT>
static int impl_lim(b200_ctx* c, void* Yc, const void* refc, cudaStream_t s) {
  const int ntr = c->dims.n_tracers, nh = c->dims.nh, nv = c->dims.nv, ng = c->dims.nh_ghost;
  if (ntr == 0) return 0;
  // second branch of lim! (limited_tendencies.jl:95-121): vertical mass borrowing, column-local, after the SEM limiter
  auto vborrow = [&]() -> int {
    if (!c->prm.vertical_water_borrowing_limiter) return 0;
    k_lim_vborrow<FT><<<dim3(nh, ntr), NT, 2 * SLAB * sizeof(FT), s>>>((const VLev<FT>*)c->d_vlev, (FT*)Yc, c->ncf(), nv, (FT)0);
    LAUNCH_CHECK(c);
    return 0;
  };
  if (!c->prm.sem_quasimonotone_limiter) return vborrow();
  const bool multi = c->nranks > 1 && !c->nbr.empty();
  if (multi && (!c->p2p_ready || getenv("B200_HALO_NCCL")))
    return fail("b200_lim: on multi-rank contexts the limiter needs the peer-memory halo (neighbour bounds travel through it)");
  if (!c->d_lim_bnd) CK(cudaMalloc(&c->d_lim_bnd, (size_t)ntr * nh * 2 * LV * sizeof(FT)));
  if (multi && !c->d_lim_E) {
    CK(cudaMalloc(&c->d_lim_E, (size_t)nh * 2 * ntr * 16 * nv * sizeof(FT)));
    // a received node column of every ghost element: any ghost member of a DSS record
    std::vector<int32_t> gn(std::max(1, ng), 0);
    for (int32_t m : c->h_mem)
      if ((m >> 4) >= nh) gn[(m >> 4) - nh] = m & 15;
    CK(cudaMalloc(&c->d_lim_ghost_node, gn.size() * sizeof(int32_t)));
    CK(cudaMemcpy(c->d_lim_ghost_node, gn.data(), gn.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  k_lim_bounds<FT><<<dim3(nh, ntr), 64, 0, s>>>((const FT*)refc, c->ncf(), nv, nh, (FT*)c->d_lim_bnd, multi ? (FT*)c->d_lim_E : nullptr, ntr);
  LAUNCH_CHECK(c);
  if (multi) {  // one more exchange of the peer-memory halo: the bounds field of the send elements
    P2PArgs PA;
    PA.nfields = 1;
    PA.f[0] = {c->d_lim_E, 2 * ntr * 16 * nv, 0, 2 * ntr, nv};
    if (c->n_send > 0) launchx(0, k_pack_p2p<FT>, dim3(c->n_send), dim3(256), 0, s, PA, p2p_plan(c));
    else k_p2p_signal<<<1, 32, 0, s>>>(c->d_p2p_flags, (int)c->nbr.size(), c->d_p2p_seq);
    LAUNCH_CHECK(c);
    k_p2p_wait<<<1, 32, 0, s>>>(reinterpret_cast<const int*>((char*)c->p2p_buf + 2 * c->p2p_cap), c->d_nbr_rank, (int)c->nbr.size(), c->d_p2p_seq);
    LAUNCH_CHECK(c);
  }
  k_lim_apply<FT><<<dim3(nh, ntr), 64, 0, s>>>((FT*)Yc, c->ncf(), nv, nh, (const FT*)c->d_lim_bnd, c->d_lim_nbr_off, c->d_lim_nbr, (const FT*)c->d_hgeo,
                                            multi ? (const FT*)c->p2p_buf : nullptr, (long long)c->p2p_cap, c->d_p2p_seq, c->d_lim_ghost_node, ntr);
  LAUNCH_CHECK(c);
  return vborrow();
}
extern "C" int b200_lim(b200_ctx* c, void* Yc, void* Yf, const void* ref_Yc, const void* ref_Yf, double, void* stream) {
  CtxScope scope_(c);
  if (!c) return fail("b200_lim: null context");
  if (halo_failed(c, "b200_lim")) return -1;
  (void)Yf; (void)ref_Yf;
  return c->ft == 4 ? impl_lim<float>(c, Yc, ref_Yc, (cudaStream_t)stream) : impl_lim<double>(c, Yc, ref_Yc, (cudaStream_t)stream);
}

// U = dss!(u + Σ c_j T_j) in one pass (k_axpy_dss); returns 1 when this context/call cannot use it (caller falls back).
// Multi-rank contexts with the peer-memory halo: the send elements are assembled straight into the neighbours' ghost blocks
// (k_pack_axpy_p2p), the nodes without ghost members and the interior nodes are processed while those slabs are in flight, then the
// flag wait and the ghost-touching nodes.
template <class FT>
static int impl_axpy_dss(b200_ctx* c, void* Uc, void* Uf, const void* uc, const void* uf, int n, const void* const* Tc,
                         const void* const* Tf, const double* coef, cudaStream_t s, unsigned dmask = 0) {
  const int nh = c->dims.nh, nv = c->dims.nv;
  const bool small = (size_t)(nh + c->dims.nh_ghost) * c->ncf() * 16 * (size_t)(nv + 1) < (size_t)INT32_MAX;
  const bool multi = c->comm != nullptr || !c->nbr.empty() || c->dims.nh_ghost > 0;
  const bool p2p = multi && c->p2p_ready && !getenv("B200_HALO_NCCL");
  if (!c->fuse_axdss || (multi && !p2p) || !small) return 1;
  AxDssArgs<FT> A;
  int m = 0;
  A.dmask = 0;
  for (int k = 0; k < n; ++k) {
    if (coef[k] == 0.0) continue;
    if (m >= AXPY_MAX) return 1;
    if (dmask >> k & 1) A.dmask |= 1u << m;
    A.Tc[m] = (const FT*)Tc[k]; A.Tf[m] = (const FT*)Tf[k]; A.c[m] = (FT)coef[k]; ++m;
  }
  if (m < 1) return 1;
  A.out_c = (FT*)Uc; A.out_f = (FT*)Uf; A.base_c = (const FT*)uc; A.base_f = (const FT*)uf; A.ncf = c->ncf(); A.nv = nv;
  A.ghost = (const FT*)c->p2p_buf; A.gpar = (long long)c->p2p_cap; A.seq = c->d_p2p_seq; A.nh = nh; A.nh_ghost = c->dims.nh_ghost;
  const DssNode<FT>* rec = (const DssNode<FT>*)c->d_dssrec;
  dim3 blk(64, 4);
  const int nn = (int)c->nbr.size();
  const int n_first = p2p ? c->n_int_nodes : c->nnodes;  // records of the first launch
  const int nbn1 = (n_first + 3) / 4, nbn2 = (c->nnodes - n_first + 3) / 4;
#define AXD_CASES(STMT) \
  switch (m) { case 1: { constexpr int N_ = 1; STMT; } break; case 2: { constexpr int N_ = 2; STMT; } break; \
               case 3: { constexpr int N_ = 3; STMT; } break; case 4: { constexpr int N_ = 4; STMT; } break; \
               case 5: { constexpr int N_ = 5; STMT; } break; case 6: { constexpr int N_ = 6; STMT; } break; \
               case 7: { constexpr int N_ = 7; STMT; } break; case 8: { constexpr int N_ = 8; STMT; } break; default: return 1; }
  if (p2p && c->n_send > 0) {  // the first n_send blocks assemble + send the boundary columns and signal
    AXD_CASES((launchx(c->pdl & 8, k_axpy_dss<FT, N_, false, true>, dim3(c->n_send + nbn1 + nh), blk, 0, s, A, rec, 0, n_first, nbn1, nh,
                       p2p_wait_args(c, false), p2p_plan(c))));
  } else {
    if (p2p) { k_p2p_signal<<<1, 32, 0, s>>>(c->d_p2p_flags, nn, c->d_p2p_seq); LAUNCH_CHECK(c); }
    AXD_CASES((launchx(c->pdl & 8, k_axpy_dss<FT, N_, false, false>, dim3(nbn1 + nh), blk, 0, s, A, rec, 0, n_first, nbn1, nh,
                       p2p_wait_args(c, false), P2PPlan())));
  }
  LAUNCH_CHECK(c);
  if (p2p) {
    if (nbn2 > 0) {  // these blocks poll the neighbours' flags themselves
      AXD_CASES((launchx(c->pdl & 8, k_axpy_dss<FT, N_, true, false>, dim3(nbn2), blk, 0, s, A, rec, n_first, c->nnodes, nbn2, 0, p2p_wait_args(c, true), P2PPlan())));
    } else {
      k_p2p_wait<<<1, 32, 0, s>>>(reinterpret_cast<const int*>((char*)c->p2p_buf + 2 * c->p2p_cap), c->d_nbr_rank, nn, c->d_p2p_seq);
    }
    LAUNCH_CHECK(c);
  }
#undef AXD_CASES
  return 0;
}

// ---------------------------------------------------------------------------------------------
// One fused implicit stage (one Newton iteration of the CTS implicit solve, integrator.jl:63-120): N = U − J⁻¹·R(U) with
// cache_imp!, Wfact, T_imp!, ldiv! and T_post_imp! in ONE kernel; U may carry unfiltered u₃ boundary values.
template <class FT>
static int impl_imp_stage(b200_ctx* c, void* Nc, void* Nf, const void* Uc, const void* Uf, double dtg, cudaStream_t s) {
  const size_t bc = c->nc() * sizeof(FT), bf = c->nf() * sizeof(FT);
  if (vdiff_implicit(c)) {  // implicit vertical diffusion: the fused stage with the diffusion blocks and the approximate arrowhead iteration
    // warp per column pair, in-warp tridiagonal solves (kernels_imp8d.cuh); B200_VDIFF_STAGE=1 selects the one-point-per-thread
    // k_imp_stage_diff (the first generation, kept as the Float32-emulatable reference of the CPU tests).  The intermediate packed
    // shared-memory version (block-wide PCR, 3.24 ms/step) was removed once its A/B record was in profiles/.
    static const int stage_sel = getenv("B200_VDIFF_STAGE") ? atoi(getenv("B200_VDIFF_STAGE")) : 8;
    if (stage_sel == 1)
      k_imp_stage_diff<FT><<<c->dims.nh * 4, NT, (size_t)QD_PROFILES * 4 * LVP * sizeof(FT), s>>>(
          make_par<FT>(c), make_vdiff<FT>(c), (const FT*)c->d_hgeo, (const VLev<FT>*)c->d_vlev, (const FT*)Uc, (const FT*)Uf, (FT*)Nc, (FT*)Nf, (FT)dtg);
    else if (c->dims.nv == 63 && !c->generic_nv)
      launchx(c->pdl & 16, k8_imp_stage_diff<FT, 63>, c->dims.nh, 256, 0, s, make_par<FT>(c), make_vdiff<FT>(c), (const FT*)c->d_hgeo,
              (const VLev<FT>*)c->d_vlev, (const FT*)Uc, (const FT*)Uf, (FT*)Nc, (FT*)Nf, (FT)dtg);
    else
      launchx(c->pdl & 16, k8_imp_stage_diff<FT, 0>, c->dims.nh, 256, 0, s, make_par<FT>(c), make_vdiff<FT>(c), (const FT*)c->d_hgeo,
              (const VLev<FT>*)c->d_vlev, (const FT*)Uc, (const FT*)Uf, (FT*)Nc, (FT*)Nf, (FT)dtg);
    LAUNCH_CHECK(c);
    return 0;
  }
  // moist stage: two saturation adjustments per point dominate, the shared-memory layout measures 242 µs against 256 µs for the
  // warp-per-column-pair one (profiles/r2_moist.md): k8 only on request (B200_IMP_KERNEL=88)
  if (c->prm.microphysics_0M && c->imp_kernel == 88 && c->dims.nv == 63 && !c->generic_nv)
    launchx(c->pdl & 16, k8_imp_stage<FT, 63, false, true>, c->dims.nh, 256, 0, s, make_par<FT>(c), (const FT*)c->d_hgeo, (const VLev<FT>*)c->d_vlev,
            (const FT*)Uc, (const FT*)Uf, (FT*)Nc, (FT*)Nf, (FT)dtg, (const FT*)nullptr, (const FT*)nullptr);
  else if (c->prm.microphysics_0M && c->imp_kernel == 88)
    launchx(c->pdl & 16, k8_imp_stage<FT, 0, false, true>, c->dims.nh, 256, 0, s, make_par<FT>(c), (const FT*)c->d_hgeo, (const VLev<FT>*)c->d_vlev,
            (const FT*)Uc, (const FT*)Uf, (FT*)Nc, (FT*)Nf, (FT)dtg, (const FT*)nullptr, (const FT*)nullptr);
  else if (c->prm.microphysics_0M && c->dims.nv == 63 && !c->generic_nv)
    launchx(c->pdl & 16, k5_imp_stage<FT, 63, false, true>, c->dims.nh, 256, smem_imp5<FT>(true), s, make_par<FT>(c), (const FT*)c->d_hgeo,
            (const VLev<FT>*)c->d_vlev, (const FT*)Uc, (const FT*)Uf, (FT*)Nc, (FT*)Nf, (FT)dtg, (const FT*)nullptr, (const FT*)nullptr);
  else if (c->prm.microphysics_0M)
    launchx(c->pdl & 16, k5_imp_stage<FT, 0, false, true>, c->dims.nh, 256, smem_imp5<FT>(true), s, make_par<FT>(c), (const FT*)c->d_hgeo,
            (const VLev<FT>*)c->d_vlev, (const FT*)Uc, (const FT*)Uf, (FT*)Nc, (FT*)Nf, (FT)dtg, (const FT*)nullptr, (const FT*)nullptr);
  else if (c->imp_kernel == 8 && c->dims.nv == 63 && !c->generic_nv)  // warp per column pair: no shared memory, shuffle PCR (kernels_imp8.cuh)
    launchx(c->pdl & 16, k8_imp_stage<FT, 63>, c->dims.nh, 256, 0, s, make_par<FT>(c), (const FT*)c->d_hgeo, (const VLev<FT>*)c->d_vlev,
            (const FT*)Uc, (const FT*)Uf, (FT*)Nc, (FT*)Nf, (FT)dtg, (const FT*)nullptr, (const FT*)nullptr);
  else if (c->imp_kernel == 8)
    launchx(c->pdl & 16, k8_imp_stage<FT, 0>, c->dims.nh, 256, 0, s, make_par<FT>(c), (const FT*)c->d_hgeo, (const VLev<FT>*)c->d_vlev,
            (const FT*)Uc, (const FT*)Uf, (FT*)Nc, (FT*)Nf, (FT)dtg, (const FT*)nullptr, (const FT*)nullptr);
  else if (c->dims.nv == 63 && !c->generic_nv)
    launchx(c->pdl & 16, k5_imp_stage<FT, 63>, c->dims.nh, 256, smem_imp5<FT>(), s, make_par<FT>(c), (const FT*)c->d_hgeo, (const VLev<FT>*)c->d_vlev,
            (const FT*)Uc, (const FT*)Uf, (FT*)Nc, (FT*)Nf, (FT)dtg, (const FT*)nullptr, (const FT*)nullptr);
  else
    launchx(c->pdl & 16, k5_imp_stage<FT, 0>, c->dims.nh, 256, smem_imp5<FT>(), s, make_par<FT>(c), (const FT*)c->d_hgeo, (const VLev<FT>*)c->d_vlev,
            (const FT*)Uc, (const FT*)Uf, (FT*)Nc, (FT*)Nf, (FT)dtg, (const FT*)nullptr, (const FT*)nullptr);
  LAUNCH_CHECK(c);
  return 0;
}
extern "C" int b200_implicit_stage(b200_ctx* c, void* Nc, void* Nf, const void* Uc, const void* Uf, double dtgamma, void* stream) {
  CtxScope scope_(c);
  if (!c) return fail("b200_implicit_stage: null context");
  return c->ft == 4 ? impl_imp_stage<float>(c, Nc, Nf, Uc, Uf, dtgamma, (cudaStream_t)stream)
                    : impl_imp_stage<double>(c, Nc, Nf, Uc, Uf, dtgamma, (cudaStream_t)stream);
}

template <class FT>
static int impl_step(b200_ctx* c, void* Yc, void* Yf, int fused, cudaStream_t s) {
  const size_t bc = c->nc() * sizeof(FT), bf = c->nf() * sizeof(FT);
  if (!c->Uc[0]) {
    for (int i = 0; i < 2; ++i) { CK(cudaMalloc(&c->Uc[i], bc)); CK(cudaMalloc(&c->Uf[i], bf)); }
    for (int i = 0; i < 4; ++i) {
      CK(cudaMalloc(&c->Tec[i], bc)); CK(cudaMalloc(&c->Tef[i], bf));
      CK(cudaMalloc(&c->Tic[i], bc)); CK(cudaMalloc(&c->Tif[i], bf));
    }
    CK(cudaMalloc(&c->Rc, bc)); CK(cudaMalloc(&c->Rf, bf)); CK(cudaMalloc(&c->dc, bc)); CK(cudaMalloc(&c->df, bf));
  }
  const Tableau tb = ars343();
  const double dt = c->prm.dt;
  // lim! between the limited and the unlimited increments (CTS update_stage!): only when a limiter is configured AND tracers exist
  const bool limiter = (c->prm.sem_quasimonotone_limiter || c->prm.vertical_water_borrowing_limiter) && c->dims.n_tracers > 0;
  if (limiter)
    for (int i = 0; i < 4; ++i)
      if (!c->Tlc[i]) CK(cudaMalloc(&c->Tlc[i], bc));
  // U = u + dt Σ coef_j T_lim[j] (centres; faces are copied), then lim!(U, ref = u)
  auto limited_increment = [&](void* Uc_, void* Uf_, const double* coef, int nst) -> int {
    const void* Tl[4]; double cl[4]; int nl = 0;
    for (int j = 0; j < nst; ++j)
      if (coef[j] != 0) { Tl[nl] = c->Tlc[j]; cl[nl++] = dt * coef[j]; }
    if (launch_axpy<FT>(c, (FT*)Uc_, (const FT*)Yc, nl, (const FT* const*)Tl, cl, c->nc(), s)) return -1;
    if (launch_axpy<FT>(c, (FT*)Uf_, (const FT*)Yf, 0, (const FT* const*)Tl, cl, c->nf(), s)) return -1;
    return nl > 0 ? impl_lim<FT>(c, Uc_, Yc, s) : 0;
  };
  bool stiff = fused && c->stiff_final && !limiter;
  for (int j = 0; j < 4; ++j) stiff = stiff && tb.bi[j] == tb.ai[3][j];
  // Stage-solution form of the increments (fused path, stiffly accurate tableau).  T_imp[j] ≡ (N_j − U_j)/(dt·a_imp[j][j]) only
  // ever enters later increments, and U_j is itself an increment, so by recursion every stage state is
  //     U_i = u + Σ_j α_ij (N_j − u) + dt Σ_j β_ij T_exp[j]
  // with host-computed α, β: the increment kernels read the stage solutions N_j directly (term c·(N_j − u), the difference of two
  // close numbers is exact) and T_imp is never formed.  The two 3·S passes that formed it could not overlap anything (every kernel
  // of the step fills the register file) and cost 86 µs per step.  Same reads per increment as before (N_j replaces T_imp[j]).
  const bool zform = stiff && c->zform;
  double al[4][4] = {{0}}, bt[4][4] = {{0}};
  if (zform) {
    for (int i = 1; i < 4; ++i) {
      for (int j = 0; j < i; ++j) bt[i][j] = tb.ae[i][j];
      for (int j = 1; j < i; ++j) {
        if (tb.ai[i][j] == 0) continue;
        const double w = tb.ai[i][j] / tb.ai[j][j];
        al[i][j] += w;
        for (int k = 0; k < j; ++k) { al[i][k] -= w * al[j][k]; bt[i][k] -= w * bt[j][k]; }
      }
    }
    for (int i = 1; i < 4; ++i)
      if (!c->Nsc[i]) { CK(cudaMalloc(&c->Nsc[i], bc)); CK(cudaMalloc(&c->Nsf[i], bf)); }
  }
  auto dss_state = [&](void* ac, void* af) -> int {
    DssField F[2] = {{ac, c->ncf(), 0, 2}, {af, 1, 1, 0}};
    return impl_dss<FT>(c, F, 2, s);
  };
  if (fused) {  // stage 1 evaluates T_exp(u) directly: apply cache_imp!'s u₃ boundary filter to the incoming state (the literal path calls
                // impl_cache_imp below; later stages get it from the increment kernels)
    const int ncols = c->dims.nh * 16;
    launchx(c->pdl & 8, k_u3_filter<FT>, dim3((2 * ncols + 255) / 256), dim3(256), 0, s, (FT*)Yf, ncols, c->dims.nv + 1);
    LAUNCH_CHECK(c);
  }
  for (int i = 0; i < 4; ++i) {
    void *Uc = Yc, *Uf = Yf;  // stage 1: U = u
    if (i > 0) {
      Uc = c->Uc[0]; Uf = c->Uf[0];
      const void* Tc[8]; const void* Tf[8]; double cf[8]; int n = 0;
      unsigned dmask = 0;
      if (zform) {
        for (int j = 1; j < i; ++j)
          if (al[i][j] != 0) { dmask |= 1u << n; Tc[n] = c->Nsc[j]; Tf[n] = c->Nsf[j]; cf[n++] = al[i][j]; }
        for (int j = 0; j < i; ++j)
          if (bt[i][j] != 0) { Tc[n] = c->Tec[j]; Tf[n] = c->Tef[j]; cf[n++] = dt * bt[i][j]; }
      } else {
        for (int j = 0; j < i; ++j) {
          if (tb.ae[i][j] != 0) { Tc[n] = c->Tec[j]; Tf[n] = c->Tef[j]; cf[n++] = dt * tb.ae[i][j]; }
          if (tb.ai[i][j] != 0) { Tc[n] = c->Tic[j]; Tf[n] = c->Tif[j]; cf[n++] = dt * tb.ai[i][j]; }
        }
      }
      // fused path: the u₃ boundary filter of cache_imp! is folded into the increment (the DSS keeps zeros)
      // fused path: increment and DSS in one kernel where possible (single rank), else two passes
      if (limiter) {
        if (limited_increment(Uc, Uf, tb.ae[i], i)) return -1;
        if (impl_axpy<FT>(c, Uc, Uf, Uc, Uf, n, Tc, Tf, cf, s, fused != 0)) return -1;
        if (dss_state(Uc, Uf)) return -1;
      } else {
        int rc = fused ? impl_axpy_dss<FT>(c, Uc, Uf, Yc, Yf, n, Tc, Tf, cf, s, dmask) : 1;
        if (rc < 0) return -1;
        if (rc == 1) {
          if (impl_axpy<FT>(c, Uc, Uf, Yc, Yf, n, Tc, Tf, cf, s, fused != 0, dmask)) return -1;
          if (dss_state(Uc, Uf)) return -1;
        }
      }
      const double dtg = dt * tb.ai[i][i];
      void *Nc = zform ? c->Nsc[i] : c->Uc[1], *Nf = zform ? c->Nsf[i] : c->Uf[1];  // Newton-updated state
      if (fused) {
        if (impl_imp_stage<FT>(c, Nc, Nf, Uc, Uf, dtg, s)) return -1;
      } else {
        if (impl_cache_imp<FT>(c, Uc, Uf, nullptr, s)) return -1;
        // temp = U stays in (Uc, Uf); the Newton iterate starts as U, so Wfact / T_imp! read U itself and the update is written out of
        // place (N = U − ΔU): the same values as copying U into N first, without the two state copies
        if (impl_wfact<FT>(c, Uc, Uf, dtg, s)) return -1;
        if (impl_t_imp<FT>(c, c->Rc, c->Rf, Uc, Uf, s)) return -1;
        {  // R = temp + dtγ·T_imp(U) − U
          const void* Tc[2] = {c->Rc, Uc}; const void* Tf[2] = {c->Rf, Uf}; double cf[2] = {dtg, -1.0};
          if (impl_axpy<FT>(c, c->Rc, c->Rf, Uc, Uf, 2, Tc, Tf, cf, s)) return -1;
        }
        if (impl_ldiv<FT>(c, c->dc, c->df, c->Rc, c->Rf, s)) return -1;
        {
          const void* Tc[1] = {c->dc}; const void* Tf[1] = {c->df}; double cf[1] = {-1.0};
          if (impl_axpy<FT>(c, Nc, Nf, Uc, Uf, 1, Tc, Tf, cf, s)) return -1;
        }
        if (impl_cache_imp<FT>(c, Nc, Nf, nullptr, s)) return -1;
        if (c->prm.energy_upwinding != 0) {
          if (impl_t_post<FT>(c, c->Rc, c->Rf, Nc, Nf, s)) return -1;
          const void* Tc[1] = {c->Rc}; const void* Tf[1] = {c->Rf}; double cf[1] = {dtg};
          if (impl_axpy<FT>(c, Nc, Nf, Nc, Nf, 1, Tc, Tf, cf, s)) return -1;
        }
      }
      if (dss_state(Nc, Nf)) return -1;
      if (!fused && impl_cache_imp<FT>(c, Nc, Nf, nullptr, s)) return -1;
      // T_imp[i] = (U − temp)/dtγ.  It only feeds later increments, so the fused path runs this HBM-bound pass on
      // a side stream, concurrently with the latency-bound T_exp kernels of the same stage.
      // Stiffly accurate tableau (the last row of A_imp equals b, ARS343): the stage-4 solution already contains every implicit
      // term of the step, u_new = N₄ + dt Σ_j (b_j − a_exp[4][j]) T_exp[j]  (T_imp[4] ≡ (N₄ − U₄)/dtγ by definition), so the fused
      // path neither forms T_imp[4] nor reads u and the three T_imp vectors in the final increment (5 vectors instead of 7).
      if (zform || (stiff && i == 3)) { Uc = Nc; Uf = Nf; goto t_exp_of_stage; }
      cudaStream_t sd = s;
      if (fused) {
        if (!c->side) {
          CK(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
          CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
          CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
        }
        CK(cudaEventRecord(c->ev_fork, s));
        CK(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
        sd = c->side;
      }
      if (launch_diff_scale<FT>(c, (FT*)c->Tic[i], (const FT*)Nc, (const FT*)Uc, (FT)dtg, c->nc(), sd)) return -1;
      if (launch_diff_scale<FT>(c, (FT*)c->Tif[i], (const FT*)Nf, (const FT*)Uf, (FT)dtg, c->nf(), sd)) return -1;
      if (sd != s) { CK(cudaEventRecord(c->ev_join, sd)); }
      Uc = Nc; Uf = Nf;
    } else if (!fused) {
      if (impl_cache_imp<FT>(c, Uc, Uf, nullptr, s)) return -1;  // no-op on a filtered state; kept for the hook trace
    }
  t_exp_of_stage:
    if (impl_t_exp<FT>(c, c->Tec[i], c->Tef[i], limiter ? c->Tlc[i] : nullptr, nullptr, Uc, Uf, s)) return -1;
    if (i > 0 && fused && !zform && !(stiff && i == 3)) CK(cudaStreamWaitEvent(s, c->ev_join, 0));  // join the side stream
  }
  {
    const void* Tc[8]; const void* Tf[8]; double cf[8]; int n = 0;
    const void *bc_ = Yc, *bf_ = Yf;
    if (stiff) {
      bc_ = zform ? c->Nsc[3] : c->Uc[1]; bf_ = zform ? c->Nsf[3] : c->Uf[1];  // N₄
      for (int j = 0; j < 4; ++j)
        if (tb.be[j] - tb.ae[3][j] != 0) { Tc[n] = c->Tec[j]; Tf[n] = c->Tef[j]; cf[n++] = dt * (tb.be[j] - tb.ae[3][j]); }
    } else {
      for (int j = 0; j < 4; ++j) {
        if (tb.be[j] != 0) { Tc[n] = c->Tec[j]; Tf[n] = c->Tef[j]; cf[n++] = dt * tb.be[j]; }
        if (tb.bi[j] != 0) { Tc[n] = c->Tic[j]; Tf[n] = c->Tif[j]; cf[n++] = dt * tb.bi[j]; }
      }
    }
    if (limiter) {
      if (limited_increment(c->Uc[0], c->Uf[0], tb.be, 4)) return -1;
      if (impl_axpy<FT>(c, Yc, Yf, c->Uc[0], c->Uf[0], n, Tc, Tf, cf, s, fused != 0)) return -1;
      if (dss_state(Yc, Yf)) return -1;
    } else {
      int rc = fused ? impl_axpy_dss<FT>(c, Yc, Yf, bc_, bf_, n, Tc, Tf, cf, s) : 1;
      if (rc < 0) return -1;
      if (rc == 1) {
        if (impl_axpy<FT>(c, Yc, Yf, bc_, bf_, n, Tc, Tf, cf, s, fused != 0)) return -1;
        if (dss_state(Yc, Yf)) return -1;
      }
    }
  }
  return fused ? 0 : impl_cache_imp<FT>(c, Yc, Yf, nullptr, s);
}
extern "C" int b200_step_ars343(b200_ctx* c, void* Yc, void* Yf, double, int32_t fused, void* stream) {
  CtxScope scope_(c);
  if (!c) return fail("b200_step_ars343: null context");
  if (halo_failed(c, "b200_step_ars343")) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  auto run = [&](cudaStream_t q) { return c->ft == 4 ? impl_step<float>(c, Yc, Yf, fused, q) : impl_step<double>(c, Yc, Yf, fused, q); };
  // One CUDA graph per state buffer: ≈35 kernel launches, the side-stream fork/join and the memsets replay as one launch.
  // Multi-rank contexts replay too when the halo runs over peer memory (its exchange number lives in device memory); the NCCL
  // halo stays eager.
  const bool graphable = c->use_graph && fused && (c->nbr.empty() || (c->p2p_ready && !getenv("B200_HALO_NCCL")));
  if (!graphable) return run(s);
  if (c->eager_steps < 1) { c->eager_steps++; return run(s); }  // first step eager: performs the lazy allocations
  // the legacy default stream cannot be captured: order an internal stream after/before it with events instead
  cudaStream_t q = s;
  const bool legacy_stream = (s == nullptr || s == cudaStreamLegacy);
  if (legacy_stream) {
    if (!c->gstream) {
      CK(cudaStreamCreateWithFlags(&c->gstream, cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&c->ev_gin, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&c->ev_gout, cudaEventDisableTiming));
    }
    q = c->gstream;
    CK(cudaEventRecord(c->ev_gin, s));
    CK(cudaStreamWaitEvent(q, c->ev_gin, 0));
  }
  b200_ctx::StepGraph* G = nullptr;
  for (auto& g : c->graphs) if (g.Yc == Yc && g.Yf == Yf) G = &g;
  if (!G) {
    if (c->graphs.size() >= 4) { cudaGraphExecDestroy(c->graphs.front().exec); c->graphs.erase(c->graphs.begin()); }
    const int64_t before = c->launches;
    CK(cudaStreamBeginCapture(q, cudaStreamCaptureModeThreadLocal));
    const int rc = run(q);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(q, &g);
    if (rc != 0 || e != cudaSuccess) {
      if (g) cudaGraphDestroy(g);
      (void)cudaGetLastError();
      c->use_graph = 0;  // fall back to eager launches for this context
      c->launches = before;
      return rc != 0 ? rc : run(s);
    }
    cudaGraphExec_t ex = nullptr;
    CK(cudaGraphInstantiate(&ex, g, 0));
    cudaGraphDestroy(g);
    c->graphs.push_back({ex, Yc, Yf, c->launches - before});
    c->launches = before;
    G = &c->graphs.back();
  }
  CK(cudaGraphLaunch(G->exec, q));
  c->launches += G->launches;
  if (legacy_stream) {
    CK(cudaEventRecord(c->ev_gout, q));
    CK(cudaStreamWaitEvent(s, c->ev_gout, 0));
  }
  return 0;
}
