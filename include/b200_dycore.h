/* b200_dycore.h — C-ABI of libb200dycore.so: the B200-native dry/tracer dynamical-core step.
 *
 * Drop-in boundary: these entry points are what a `ccall` glue module binds in place of the
 * ClimaODEFunction hooks the reference wires at src/simulation/integrator.jl:215-225 (see
 * INTEGRATION.md for the Julia stub).  Conventions:
 *   - plain pointers and sizes only; every field pointer is a DEVICE pointer owned by the caller
 *     (`pointer(parent(Fields.field_values(Y.c)))`), never retained beyond the call unless
 *     registered at create time; `stream` is a cudaStream_t passed as void*;
 *   - every call is asynchronous on `stream`; no hidden device synchronisation;
 *   - return 0 on success, <0 on error (message via b200_last_error(ctx), per context); nothing throws;
 *   - FT is selected at create time (ft_bytes = 4 or 8); state pointers are FT*;
 *   - field layout is ClimaCore VIJFH: parent array (Nv, Ni, Nj, Nf, Nh), first index fastest,
 *     i.e. C order [h][f][j][i][v].  Y.c: Nf = 4 (ρ, uₕ₁, uₕ₂, ρe_tot), Nv levels;
 *     Y.f: Nf = 1 (u₃), Nv+1 levels.  Nq = 4 only.
 *   - there is NO CPU fallback: every compute entry point requires a CUDA device.
 *   - multi-rank contexts: a DSS halo wait that times out (a peer rank died or fell out of step; ~30 s of SM clocks,
 *     B200_P2P_SPIN_LIMIT overrides) does not trap and does not hang: the kernels run on, the context records the event in a
 *     host-mapped word, and the NEXT entry point that uses the halo (b200_dss, b200_t_exp_lim, b200_lim, b200_step_ars343) returns
 *     <0 with a message naming the neighbour rank; the state of that context is invalid from then on, the process is not.
 *
 * Limits (checked by b200_create, which fails with a message instead of computing something else):
 *   Nq = 4 (nh_poly 3); 2 <= nv <= 63 (a column of faces is one 64-lane row of the kernels); flat grid (NoWarp topography: the
 *   metric is used in factored form, 2-D per-node part x per-level scale); dry thermodynamics or the equilibrium-moist (0M) state
 *   with an active rho*q_tot, up to 4 tracer components in all (rho*q_tot counts as the first); the moist state excludes vertical
 *   diffusion and the Held-Suarez forcing; <= 32 neighbour ranks.
 *   b200_step_ars343 is ARS343 with ONE Newton iteration per implicit stage (max_newton_iters_ode: 1, the reference's setting
 *   for these configurations); any other stepper / Newton loop drives the individual hook entry points.  ldiv! is the direct
 *   BlockArrowheadSolve (or the ApproximateBlockArrowheadIterativeSolve with implicit vertical diffusion); the Krylov
 *   method of implicit/jacobian.jl:87-96 is not served.
 */
#ifndef B200_DYCORE_H
#define B200_DYCORE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_ctx b200_ctx;

/* Sizes. Replaces what the glue reads from `axes(Y.c)` (ClimaCore Spaces). */
typedef struct {
  int32_t nh;        /* local elements (this rank) */
  int32_t nh_ghost;  /* ghost elements appended after the local ones (0 for 1 GPU) */
  int32_t nv;        /* centre levels (z_elem); faces = nv + 1; nv + 1 <= 64 */
  int32_t nq;        /* GLL nodes per direction; must be 4 */
  int32_t ft_bytes;  /* 4 = Float32, 8 = Float64 */
  int32_t deep;      /* 1 = DeepSphericalGlobalGeometry, 0 = shallow (grids.jl:64-68) */
  int32_t n_tracers; /* grid-scale tracers ρχ appended to Y.c after ρe_tot (0..4): Y.c has Nf = 4 + n_tracers; with
                        params.microphysics_0M the first of them is ρq_tot (prognostic_variables.jl:54-61), the others passive */
} b200_dims;

/* Geometry, copied from the live ClimaCore objects (HOST pointers, double precision; the
 * library converts to FT).  Horizontal arrays are [nh + nh_ghost][4][4] in (j, i) order. */
typedef struct {
  const double* dxdxi;  /* [nh+g][j][i][2][2]: ∂x_a/∂ξ_b, local (east,north) basis, at radius */
  const double* J2;     /* [nh+g][j][i] horizontal Jacobian at radius */
  const double* lat;    /* [nh+g][j][i] degrees */
  const double* gll_w;  /* [4] quadrature weights */
  const double* gll_D;  /* [4][4] D[i][k] = l'_k(ξ_i) */
  const double* z_c;    /* [nv]   */
  const double* z_f;    /* [nv+1] */
  const double* dz_c;   /* [nv]   vertical ∂z/∂ξ³ at centres */
  const double* dz_f;   /* [nv+1] vertical ∂z/∂ξ³ at faces */
  double radius;
  double z_max;
} b200_geometry;

/* Connectivity in ClimaCore Topology2D form (HOST pointers, 0-based).  Element ids >= nh refer
 * to ghost elements.  Replaces Topologies.Topology2D tables used by Spaces.weighted_dss!. */
typedef struct {
  const int32_t* interior_faces;      /* [n_faces][5]: e1, f1, e2, f2, reversed */
  int32_t n_faces;
  const int32_t* local_vertices;      /* [n_lv][2]: elem, vert */
  const int32_t* local_vertex_offset; /* [n_verts + 1] */
  int32_t n_verts;
  /* multi-rank halo (all NULL/0 on one GPU) */
  int32_t n_neighbors;
  const int32_t* neighbor_ranks;   /* [n_neighbors] */
  const int32_t* send_offset;      /* [n_neighbors + 1] into send_elems */
  const int32_t* send_elems;       /* local element ids whose perimeter is sent */
  const int32_t* recv_offset;      /* [n_neighbors + 1] into ghost slots (0-based ghost index) */
  const int64_t* elem_gid;         /* [nh + nh_ghost] global element id (summation order) */
} b200_topology;

/* Parameters: subset of ClimaAtmosParameters + numerics read by the dycore
 * (src/parameters/create_parameters.jl:200-224, src/types.jl:499-503,782-866). */
typedef struct {
  double R_d, cp_d, cv_d, T_0, grav, Omega, p_ref_theta, T_surf_ref, T_min_ref, T_min_sgs;
  double dt;                        /* p.dt (van Leer Courant number) */
  double nu4_vorticity, nu4_scalar; /* ν₄ (hyperdiffusion.jl:21-28); 0,0 disables hyperdiff */
  double divergence_damping_factor;
  int32_t hyperdiff;
  int32_t rayleigh_sponge; double zd_rayleigh, alpha_rayleigh_uh, alpha_rayleigh_w;
  int32_t viscous_sponge;  double zd_viscous, kappa_2_sponge;
  int32_t energy_upwinding; /* 0 none, 1 first_order, 2 third_order (ᶠupwind3, abbreviations.jl:229-240), 3 vanleer_limiter */
  int32_t tracer_upwinding; /* same encoding (default_config.yml:321-323) */
  /* Held–Suarez forcing (src/parameterized_tendencies/radiation/held_suarez.jl:111-296); flat surface */
  int32_t held_suarez; double hs_day, hs_sigma_b, hs_dT_y, hs_T_equator, hs_dtheta_z, hs_T_min, MSLP;
  /* apply_sem_quasimonotone_limiter (config/default_configs/default_config.yml; type_getters.jl:129): lim! applies ClimaCore's
   * Limiters.QuasiMonotoneLimiter to every tracer; 0 = lim! is the reference's no-op */
  int32_t sem_quasimonotone_limiter;
  /* Vertical diffusion (src/prognostic_equations/vertical_diffusion_boundary_layer.jl:64-154; model_getters.jl:332-358):
   * vert_diff 0 = none (`~`), 1 = VerticalDiffusion (C_E), 2 = DecayWithHeightDiffusion (H_diffusion, D_0_diffusion);
   * implicit_diffusion → diff_mode (type_getters.jl:131): 0 the tendency joins T_exp (remaining_tendency.jl:185-195), 1 it joins
   * T_imp!, Wfact adds the diffusion blocks (manual_sparse_jacobian.jl:1031-1261) and ldiv! becomes the
   * ApproximateBlockArrowheadIterativeSolve with approximate_linear_solve_iters iterations (:538-578);
   * disable_momentum_vertical_diffusion: scalars only (Held–Suarez runs, type_getters.jl:46). */
  int32_t vert_diff, implicit_diffusion, approximate_linear_solve_iters, disable_momentum_vertical_diffusion;
  double C_E, H_diffusion, D_0_diffusion;
  /* tracer_nonnegativity_method: vertical_water_borrowing (default_config.yml:190-198; src/cache/cache.jl:216-219): lim! also applies
   * ClimaCore's Limiters.VerticalMassBorrowingLimiter((0,)) to χ = ρχ/ρ of every tracer (limited_tendencies.jl:95-121); 0 = off */
  int32_t vertical_water_borrowing_limiter;
  /* microphysics_model: 0 = DryModel, 1 = EquilibriumMicrophysics0M — component 4 of Y.c (the first of the n_tracers >= 1 tracer
   * components) is the thermodynamically ACTIVE rho*q_tot: moist thermodynamic state with saturation adjustment
   * (src/cache/precomputed_quantities.jl:735-815), central vertical transport + post-Newton correction of q_tot
   * (implicit_tendency.jl:210-214, 333-338), the (rho q_tot, u3) / (u3, rho q_tot) Jacobian blocks with kappa_m
   * (manual_sparse_jacobian.jl:653-690, 770-790, 827-831), grad^2 q_tot_eff and the water enthalpy / mass split of hyperdiffusion and
   * viscous sponge (hyperdiffusion.jl:148-165, 293-306, 475-484; viscous_sponge.jl:158-199).  The 0M precipitation sink itself is a
   * parameterised tendency outside the dycore.  Thermodynamics.jl parameters (docs/src/thermodynamics.md:60-150): */
  int32_t microphysics_0M;
  double R_v, cp_v, cp_l, cp_i, LH_v0, LH_s0, T_triple, press_triple, T_freeze, T_icenuc, pow_icenuc;
} b200_params;

/* Optional device pointers to p.precomputed fields written by b200_cache_imp (any may be NULL).
 * (src/cache/precomputed_quantities.jl:53-61) */
typedef struct {
  void* u_c;    /* ᶜu   C123: [nh][3][16][nv]   */
  void* u3_f;   /* ᶠu³  CT3:  [nh][1][16][nv+1] */
  void* K_c;    /* ᶜK   */
  void* T_c;    /* ᶜT   */
  void* p_c;    /* ᶜp   */
  void* h_tot_c;/* ᶜh_tot */
} b200_cacheptrs;

int b200_create(b200_ctx** out, const b200_dims*, const b200_geometry*, const b200_topology*,
                const b200_params*, const void* nccl_unique_id /* NULL for 1 GPU */, int rank,
                int nranks);
int b200_destroy(b200_ctx*);
/* Message of the last failing call on `ctx`; ctx == NULL: the calling thread's last message (b200_create failures). */
const char* b200_last_error(const b200_ctx* ctx);
/* 128-byte NCCL unique id for the DSS halo communicator (rank 0 calls, host broadcasts). */
int b200_nccl_unique_id(void* out128);

/* Peer-memory DSS halo over NVLink (optional; the NCCL send/recv path is used until it is set up).
 * b200_halo_export writes this rank's 64-byte cudaIpcMemHandle; after the host has gathered them,
 * b200_halo_import maps the neighbours' buffers: handles[q] / their_recv_offset[q] (first ghost slot my
 * slabs occupy in neighbour q) / their_nh_ghost[q], q in the order of topology.neighbor_ranks. */
int b200_halo_export(b200_ctx*, void* handle64_out);
int b200_halo_import(b200_ctx*, const void* handles, const int32_t* their_recv_offset,
                     const int32_t* their_nh_ghost);

/* cache_imp! — set_implicit_precomputed_quantities! (precomputed_quantities.jl:698-831):
 * applies the u₃ boundary filter to Yf IN PLACE and fills the optional precomputed fields. */
int b200_cache_imp(b200_ctx*, void* Yc, void* Yf, const b200_cacheptrs* out, void* stream);
/* T_exp_T_lim! — remaining_tendency! (remaining_tendency.jl:48-58). Ylc/Ylf may be NULL (dry). */
int b200_t_exp_lim(b200_ctx*, void* Ytc, void* Ytf, void* Ylc, void* Ylf, const void* Yc,
                   const void* Yf, double t, void* stream);
/* T_imp! — implicit_tendency! (implicit/implicit_tendency.jl:36-98). */
int b200_t_imp(b200_ctx*, void* Ytc, void* Ytf, const void* Yc, const void* Yf, double t,
               void* stream);
/* Wfact — update_jacobian! (implicit/jacobian.jl:74-75 → manual_sparse_jacobian.jl:1850). */
int b200_wfact(b200_ctx*, const void* Yc, const void* Yf, double dtgamma, double t, void* stream);
/* ldiv! — invert_jacobian! (implicit/jacobian.jl:78-82 → manual_sparse_jacobian.jl:1897). */
int b200_ldiv(b200_ctx*, void* dYc, void* dYf, const void* Rc, const void* Rf, void* stream);
/* T_post_imp! — correct_implicit_advection_tendency! (implicit_tendency.jl:322-339). */
int b200_t_post_imp(b200_ctx*, void* Ytc, void* Ytf, const void* Yc, const void* Yf, double t,
                    void* stream);
/* dss! — Spaces.weighted_dss! (constrain_state.jl:59-64; remaining_tendency.jl:18-21).
 * fields[k]: device pointer; nf[k]: components; is_face[k]: 0 centre / 1 face;
 * kind[k]: 0 scalars, 1 = (c12 pair followed by nf-2 scalars), 2 = (ρ, c12 pair, scalars…). */
int b200_dss(b200_ctx*, void* const* fields, const int32_t* nf, const int32_t* is_face,
             const int32_t* kind, int32_t nfields, void* stream);
/* Fused stage increment U = u + Σ_j c_j T_j over a state of (Yc, Yf) (ClimaTimeSteppers
 * fused_increment!). */
int b200_axpy_n(b200_ctx*, void* Uc, void* Uf, const void* uc, const void* uf, int32_t n,
                const void* const* Tc, const void* const* Tf, const double* coef, void* stream);
/* One full IMEX-ARK ARS343 step with one Newton iteration per implicit stage, using the hooks
 * above in the order of DESIGN.md "Step trace" (role of CTS.step!, solve.jl:62,125).  The
 * state (Yc, Yf) is advanced in place. `fused` selects the fused implicit-stage kernel. */
int b200_step_ars343(b200_ctx*, void* Yc, void* Yf, double t, int32_t fused, void* stream);
/* lim!(Y, p, t, ref_Y) (src/prognostic_equations/limited_tendencies.jl:64-122): SEM quasi-monotone limiter of every tracer ρχ of
 * Y.c (in place) with bounds from ref_Y (element min/max of χ widened over the vertex neighbours), then — with
 * params.vertical_water_borrowing_limiter — the column-wise vertical mass-borrowing limiter.  No-op unless one of the two is set
 * and n_tracers > 0.  Multi-rank contexts: needs the peer-memory halo (b200_halo_import), which
 * carries the bounds of the ghost elements. */
int b200_lim(b200_ctx*, void* Yc, void* Yf, const void* ref_Yc, const void* ref_Yf, double t, void* stream);
/* One fused implicit stage = one Newton iteration of ClimaTimeSteppers' implicit solve on the stage problem
 * (integrator.jl:63-120: initialize_imp!/cache_imp!, Wfact, T_imp!, ldiv!, U −= ΔU, cache_imp!, T_post_imp!):
 *   N = U − J(U)⁻¹ (dtγ·T_imp(U))  [+ dtγ·(vtt_upwind − vtt_central)(N) when energy upwinding is on]
 * in ONE kernel, out of place (U is read-only; its u₃ boundary faces are treated as zero). */
int b200_implicit_stage(b200_ctx*, void* Nc, void* Nf, const void* Uc, const void* Uf, double dtgamma,
                        void* stream);
/* Profiling aid: run ONE phase of b200_t_exp_lim — 0: pre-DSS kernel (Yₜ partial + ∇² fields),
 * 1: DSS of the ∇² fields, 2: hyperdiffusion apply kernel. */
int b200_t_exp_phase(b200_ctx*, int32_t phase, void* Ytc, void* Ytf, const void* Yc, const void* Yf,
                     void* stream);
/* Host-only: build the unique-perimeter-node CSR from the Topology2D tables (no device needed).
 * mem entries are elem*16 + j*4 + i.  Used by the bit-exact index-map tests; the same routine
 * feeds b200_create.  Returns -1 if the output capacities are too small. */
int b200_build_dss_csr(const b200_topology*, int32_t* off_out, int32_t cap_nodes, int32_t* mem_out,
                       int32_t cap_mem, int32_t* nnodes, int32_t* nmem);
/* Debug: the CSR actually held by a context (after dropping nodes without a local member). */
int b200_debug_dss_csr(b200_ctx*, const int32_t** off, const int32_t** mem, int32_t* nnodes,
                       int32_t* nmem);
/* Debug/test aid: copy the Jacobian coefficient planes of the last b200_wfact into a caller-owned DEVICE buffer of FT:
 * [nh][15][16][nv+1] = Schur tridiagonal (l, d, u) of the u₃ rows; the (u₃,ρ), (u₃,ρe_tot), (u₃,uₕ₁), (u₃,uₕ₂) bidiagonals
 * (lo, hi = centres f−1, f); the (ρ,u₃), (ρe_tot,u₃) bidiagonals (lo, hi = faces k, k+1) — the blocks of
 * manual_sparse_jacobian.jl:746-868, so that Wfact can be tested on its own and not only through ldiv!. */
int b200_debug_jacobian(b200_ctx*, void* dst_device, int64_t capacity_bytes, void* stream);
/* Number of kernels launched by this context since creation (bench evidence). */
int64_t b200_launch_count(b200_ctx*);

#ifdef __cplusplus
}
#endif
#endif
