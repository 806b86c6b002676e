# B200Dycore.jl — Julia glue between ClimaAtmos and libb200dycore.so (include/b200_dycore.h).
#
# NOT EXECUTED in the build image or on the GPU box (neither has `julia`; ClimaCore / ClimaTimeSteppers are not vendored —
# profiles/r2_gpu_box_probe.txt): this file is the binding a ClimaAtmos maintainer drops into `src/` together with
# julia/climaatmos_b200.patch.  It has no bodiless functions; what it cannot have here is a run.  The same calls, in the same
# order and with the same pointer conventions, are exercised by the Python twin `climaatmos.jl_b200/capi.py` (ctypes) in
# tests/test_gpu_parity.py; tests/test_grid_and_host_logic.py pins the struct layouts mirrored below and checks this file
# structurally (balanced blocks, every ccall symbol exported by the library, argument counts equal to the C prototypes).
# The multi-rank set-up deliberately uses ONLY public ClimaCore API: a serial Topology2D of the whole mesh supplies the global
# tables and geometry, and `halo_plan` partitions them exactly like climaatmos.jl_b200/partition.py (whose Python version is
# tested against the single-GPU result bitwise).  Accessors that are ClimaCore internals are marked [UPSTREAM-RECALL ClimaCore 0.15.1].
#
# Every hook has the name and signature of the reference method it replaces (file:line in the comments), mutates its
# first argument(s) and is asynchronous on the current CUDA stream.
module B200Dycore

import ClimaCore: Fields, Spaces, Topologies, Quadratures, Geometry, Meshes
import ClimaComms, CUDA, LinearAlgebra
import ..ClimaAtmos as CA
import ..ClimaAtmos.Parameters as CAP
import Thermodynamics.Parameters as TDP   # TD.TP in the reference (refstate_thermodynamics.jl:57)

const lib = get(ENV, "B200_DYCORE_LIB", joinpath(@__DIR__, "..", "deps", "libb200dycore.so"))

# ---- field-for-field mirrors of the C structs (include/b200_dycore.h) -----------------------------------------
struct Dims
    nh::Int32; nh_ghost::Int32; nv::Int32; nq::Int32; ft_bytes::Int32; deep::Int32; n_tracers::Int32
end
struct GeometryC
    dxdxi::Ptr{Float64}; J2::Ptr{Float64}; lat::Ptr{Float64}; gll_w::Ptr{Float64}; gll_D::Ptr{Float64}
    z_c::Ptr{Float64}; z_f::Ptr{Float64}; dz_c::Ptr{Float64}; dz_f::Ptr{Float64}
    radius::Float64; z_max::Float64
end
struct TopologyC
    interior_faces::Ptr{Int32}; n_faces::Int32
    local_vertices::Ptr{Int32}; local_vertex_offset::Ptr{Int32}; n_verts::Int32
    n_neighbors::Int32
    neighbor_ranks::Ptr{Int32}; send_offset::Ptr{Int32}; send_elems::Ptr{Int32}; recv_offset::Ptr{Int32}
    elem_gid::Ptr{Int64}
end
struct ParamsC
    R_d::Float64; cp_d::Float64; cv_d::Float64; T_0::Float64; grav::Float64; Omega::Float64; p_ref_theta::Float64
    T_surf_ref::Float64; T_min_ref::Float64; T_min_sgs::Float64
    dt::Float64
    nu4_vorticity::Float64; nu4_scalar::Float64
    divergence_damping_factor::Float64
    hyperdiff::Int32
    rayleigh_sponge::Int32; zd_rayleigh::Float64; alpha_rayleigh_uh::Float64; alpha_rayleigh_w::Float64
    viscous_sponge::Int32; zd_viscous::Float64; kappa_2_sponge::Float64
    energy_upwinding::Int32      # 0 none, 1 first_order, 2 third_order, 3 vanleer_limiter
    tracer_upwinding::Int32
    held_suarez::Int32
    hs_day::Float64; hs_sigma_b::Float64; hs_dT_y::Float64; hs_T_equator::Float64; hs_dtheta_z::Float64; hs_T_min::Float64
    MSLP::Float64
    sem_quasimonotone_limiter::Int32
    vert_diff::Int32             # 0 none, 1 VerticalDiffusion, 2 DecayWithHeightDiffusion
    implicit_diffusion::Int32
    approximate_linear_solve_iters::Int32
    disable_momentum_vertical_diffusion::Int32
    C_E::Float64; H_diffusion::Float64; D_0_diffusion::Float64
    vertical_water_borrowing_limiter::Int32
    microphysics_0M::Int32       # 0 DryModel, 1 EquilibriumMicrophysics0M (ρq_tot = first tracer component, thermodynamically active)
    R_v::Float64; cp_v::Float64; cp_l::Float64; cp_i::Float64; LH_v0::Float64; LH_s0::Float64
    T_triple::Float64; press_triple::Float64; T_freeze::Float64; T_icenuc::Float64; pow_icenuc::Float64
end
struct CachePtrs
    u_c::Ptr{Cvoid}; u3_f::Ptr{Cvoid}; K_c::Ptr{Cvoid}; T_c::Ptr{Cvoid}; p_c::Ptr{Cvoid}; h_tot_c::Ptr{Cvoid}
end

mutable struct Ctx
    ptr::Ptr{Cvoid}
    keep::Any   # host arrays handed to b200_create (it copies them; kept only until create returns)
end

"Backend switch: `CLIMAATMOS_DYCORE_BACKEND=b200` on a CUDA device, dry or passive-tracer configuration (what the library serves)."
enabled(Y, atmos) = get(ENV, "CLIMAATMOS_DYCORE_BACKEND", "") == "b200" && ClimaComms.device(Y.c) isa ClimaComms.CUDADevice
ctxptr(p) = (p.numerics.b200[]::Ctx).ptr      # the context created at the end of build_cache (julia/climaatmos_b200.patch)
check(rc, what, ctx = C_NULL) = rc == 0 || error("$what: " * unsafe_string(ccall((:b200_last_error, lib), Cstring, (Ptr{Cvoid},), ctx)))
# VIJFH parent array of a field on the device: (Nv, 4, 4, Nf, Nh), level fastest
dptr(f) = reinterpret(Ptr{Cvoid}, pointer(parent(Fields.field_values(f))))
stream() = reinterpret(Ptr{Cvoid}, CUDA.stream().handle)
# 0 none, 1 first_order, 2 third_order, 3 vanleer_limiter (default_config.yml:321-326)
upw(x) = x == Val(:none) ? Int32(0) : x == Val(:first_order) ? Int32(1) : x == Val(:third_order) ? Int32(2) : Int32(3)

"""
    create(Y, p; approximate_solve_iters = 1) -> Ctx

Called once at the end of `build_cache` (src/cache/cache.jl:165-305, see julia/climaatmos_b200.patch): copies geometry,
connectivity and parameters from the live ClimaCore objects into the library.
"""
function create(Y, p; approximate_solve_iters = 1)   # = B200Jacobian(...).approximate_solve_iters (YAML approximate_linear_solve_iters)
    FT = eltype(Y)
    space = axes(Y.c)
    hspace = Spaces.horizontal_space(space)
    topo = Spaces.topology(hspace)
    quad = Spaces.quadrature_style(hspace)
    Nq = Quadratures.degrees_of_freedom(quad)
    @assert Nq == 4 "libb200dycore supports nh_poly = 3 (Nq = 4) only"
    Nv = Spaces.nlevels(Spaces.center_space(space))
    comms = ClimaComms.context(Y.c)
    rank, nranks = ClimaComms.mypid(comms) - 1, ClimaComms.nprocs(comms)
    # Global tables and geometry from a SERIAL topology/space of the same mesh, quadrature and element order (public API only).
    gtopo = nranks == 1 ? topo :
            Topologies.Topology2D(ClimaComms.SingletonCommsContext(ClimaComms.CPUSingleThreaded()), Topologies.mesh(topo), topo.elemorder)
    ghspace = nranks == 1 ? hspace : Spaces.SpectralElementSpace2D(gtopo, quad)
    plan = halo_plan(gtopo, rank, nranks)                              # partition.py:partition_grid, 0-based ids
    nh, ng = plan.nh, plan.nh_ghost
    @assert nh == Topologies.nlocalelems(topo) "rank split differs from ClimaCore's: expected contiguous near-equal SFC ranges"
    dxdxi, J2, lat = extract_dxdxi_J(ghspace, Nq, plan.elems_ext)      # Float64 host vectors in (j, i) order per element
    _, w = Quadratures.quadrature_points(Float64, quad)
    D = Quadratures.differentiation_matrix(Float64, quad)              # D[i, k] = l'_k(ξ_i)
    zc, zf, dzc, dzf = vertical_jacobians(space)                       # levels and ∂z/∂ξ³ of one column at centres / faces
    gll_w = collect(Float64, w); gll_D = collect(Float64, permutedims(D))   # row-major D[i][k] for C
    faces, lv, lvo = plan.interior_faces, plan.local_vertices, plan.local_vertex_offset
    nbr, soff, selems, roff, gid = plan.neighbor_ranks, plan.send_offset, plan.send_elems, plan.recv_offset, plan.elem_gid
    params = p.params
    hd = p.atmos.numerics.hyperdiff                                     # types.jl:1850-1866 (AtmosNumerics)
    deep = Spaces.global_geometry(space) isa Geometry.DeepSphericalGlobalGeometry   # grids.jl:64-68, cache.jl:321-322
    ν₄ᵥ, ν₄ₛ = isnothing(hd) ? (0.0, 0.0) : (Float64(CA.ν₄(hd, Y).ν₄_vorticity), Float64(CA.ν₄(hd, Y).ν₄_scalar))   # hyperdiffusion.jl:21-28
    ntr = count(CA.is_tracer_var, propertynames(Y.c))
    dims = Dims(nh, ng, Nv, Nq, sizeof(FT), deep ? 1 : 0, ntr)
    geo = GeometryC(pointer(dxdxi), pointer(J2), pointer(lat), pointer(gll_w), pointer(gll_D), pointer(zc), pointer(zf),
                    pointer(dzc), pointer(dzf), Float64(CAP.planet_radius(params)), zf[end])
    topo_c = TopologyC(pointer(faces), length(faces) ÷ 5, pointer(lv), pointer(lvo), length(lvo) - 1, length(nbr),
                       pointer(nbr), pointer(soff), pointer(selems), pointer(roff), pointer(gid))
    rs, vs = p.atmos.rayleigh_sponge, p.atmos.viscous_sponge
    thp = CAP.thermodynamics_params(params)                             # Thermodynamics.Parameters accessors (TD.TP, refstate_thermodynamics.jl:57)
    p.atmos.microphysics_model isa Union{CA.DryModel, CA.EquilibriumMicrophysics0M} ||
        error("B200Dycore: only DryModel and EquilibriumMicrophysics0M are served")
    vd = p.atmos.vertical_diffusion
    prm = ParamsC(CAP.R_d(params), CAP.cp_d(params), CAP.cv_d(params), CAP.T_0(params), CAP.grav(params), CAP.Omega(params),
                  CAP.p_ref_theta(params), CAP.T_surf_ref(params), CAP.T_min_ref(params), CAP.T_min_sgs(params),
                  Float64(p.dt), ν₄ᵥ, ν₄ₛ, isnothing(hd) ? 1.0 : Float64(hd.divergence_damping_factor),
                  isnothing(hd) ? 0 : 1,
                  rs === nothing ? 0 : 1, rs === nothing ? 0.0 : Float64(rs.zd), rs === nothing ? 0.0 : Float64(rs.α_uₕ), rs === nothing ? 0.0 : Float64(rs.α_w),
                  vs === nothing ? 0 : 1, vs === nothing ? 0.0 : Float64(vs.zd), vs === nothing ? 0.0 : Float64(vs.κ₂),
                  upw(p.atmos.numerics.energy_q_tot_upwinding), upw(p.atmos.numerics.tracer_upwinding),
                  p.atmos.radiation_mode isa CA.HeldSuarezForcing ? 1 : 0,            # remaining_tendency.jl:154-155
                  CAP.day(params), CAP.σ_b(params), CAP.ΔT_y_dry(params), CAP.T_equator_dry(params), CAP.Δθ_z(params),
                  CAP.T_min_hs(params), CAP.MSLP(params),
                  p.numerics.sem_quasimonotone_limiter === nothing ? 0 : 1,
                  vd === nothing ? 0 : (vd isa CA.VerticalDiffusion ? 1 : 2),                 # types.jl:564-597
                  p.atmos.numerics.diff_mode isa CA.Implicit ? 1 : 0,                        # types.jl:1862-1863
                  Int32(approximate_solve_iters), CA.disable_momentum_vertical_diffusion(vd) ? 1 : 0,
                  vd isa CA.VerticalDiffusion ? Float64(vd.C_E) : 0.0,
                  vd isa CA.DecayWithHeightDiffusion ? Float64(vd.H) : 1.0,
                  vd isa CA.DecayWithHeightDiffusion ? Float64(vd.D₀) : 0.0,
                  p.numerics.vertical_water_borrowing_limiter === nothing ? 0 : 1,                # cache.jl:213-219
                  p.atmos.microphysics_model isa CA.EquilibriumMicrophysics0M ? 1 : 0,            # precomputed_quantities.jl:735
                  TDP.R_v(thp), TDP.cp_v(thp), TDP.cp_l(thp), TDP.cp_i(thp), TDP.LH_v0(thp), TDP.LH_s0(thp),
                  TDP.T_triple(thp), TDP.press_triple(thp), TDP.T_freeze(thp), TDP.T_icenuc(thp), TDP.pow_icenuc(thp))
    id = zeros(UInt8, 128)
    if nranks > 1
        rank == 0 && check(ccall((:b200_nccl_unique_id, lib), Cint, (Ptr{UInt8},), id), "b200_nccl_unique_id")
        id = ClimaComms.bcast(comms, id)
    end
    out = Ref{Ptr{Cvoid}}()
    GC.@preserve dxdxi J2 lat gll_w gll_D zc zf dzc dzf faces lv lvo nbr soff selems roff gid id begin
        check(ccall((:b200_create, lib), Cint,
                    (Ref{Ptr{Cvoid}}, Ref{Dims}, Ref{GeometryC}, Ref{TopologyC}, Ref{ParamsC}, Ptr{UInt8}, Cint, Cint),
                    out, dims, geo, topo_c, prm, nranks > 1 ? pointer(id) : Ptr{UInt8}(C_NULL), rank, nranks), "b200_create")
    end
    ctx = Ctx(out[], nothing)
    nranks > 1 && setup_peer_halo!(ctx, comms, plan)       # b200_halo_export / allgather of the 64-byte handles / b200_halo_import
    finalizer(c -> ccall((:b200_destroy, lib), Cint, (Ptr{Cvoid},), c.ptr), ctx)
    return ctx
end

# ---- the hooks (ClimaODEFunction, src/simulation/integrator.jl:215-225) ------------------------------------------
function remaining_tendency!(Yₜ, Yₜ_lim, Y, p, t)            # src/prognostic_equations/remaining_tendency.jl:48
    check(ccall((:b200_t_exp_lim, lib), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}),
                ctxptr(p), dptr(Yₜ.c), dptr(Yₜ.f), dptr(Yₜ_lim.c), dptr(Yₜ_lim.f), dptr(Y.c), dptr(Y.f), Float64(t), stream()),
          "b200_t_exp_lim", ctxptr(p))
    return Yₜ
end
implicit_tendency!(Yₜ, Y, p, t) =                             # implicit/implicit_tendency.jl:36
    check(ccall((:b200_t_imp, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}),
                ctxptr(p), dptr(Yₜ.c), dptr(Yₜ.f), dptr(Y.c), dptr(Y.f), Float64(t), stream()), "b200_t_imp", ctxptr(p))
correct_implicit_advection_tendency!(Yₜ, Y, p, t) =           # implicit/implicit_tendency.jl:322 (T_post_imp!)
    check(ccall((:b200_t_post_imp, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}),
                ctxptr(p), dptr(Yₜ.c), dptr(Yₜ.f), dptr(Y.c), dptr(Y.f), Float64(t), stream()), "b200_t_post_imp", ctxptr(p))
function set_implicit_precomputed_quantities!(Y, p, t)        # cache/precomputed_quantities.jl:698 (cache_imp!)
    pc = p.precomputed
    cp = CachePtrs(dptr(pc.ᶜu), dptr(pc.ᶠu³), dptr(pc.ᶜK), dptr(pc.ᶜT), dptr(pc.ᶜp), dptr(pc.ᶜh_tot))
    check(ccall((:b200_cache_imp, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{CachePtrs}, Ptr{Cvoid}),
                ctxptr(p), dptr(Y.c), dptr(Y.f), cp, stream()), "b200_cache_imp", ctxptr(p))
end
function dss!(Y, p, t)                                         # prognostic_equations/constrain_state.jl:59
    fields = Ptr{Cvoid}[dptr(Y.c), dptr(Y.f)]
    ncomp = Int32[size(parent(Y.c), 4), 1]; is_face = Int32[0, 1]; kind = Int32[2, 0]   # kind 2: (ρ, Covariant12 pair, scalars…)
    GC.@preserve fields ncomp is_face kind check(ccall((:b200_dss, lib), Cint,
        (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Int32, Ptr{Cvoid}),
        ctxptr(p), fields, ncomp, is_face, kind, Int32(2), stream()), "b200_dss", ctxptr(p))
end
limiters_func!(Y, p, t, ref_Y) =                              # prognostic_equations/limited_tendencies.jl:64 (lim!)
    check(ccall((:b200_lim, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}),
                ctxptr(p), dptr(Y.c), dptr(Y.f), dptr(ref_Y.c), dptr(ref_Y.f), Float64(t), stream()), "b200_lim", ctxptr(p))

# ---- Jacobian: the three-method JacobianAlgorithm contract of implicit/jacobian.jl:16-26 ---------------------------
struct B200Jacobian <: CA.JacobianAlgorithm
    approximate_solve_iters::Int     # as ManualSparseJacobian.approximate_solve_iters (manual_sparse_jacobian.jl:45-60); used when diff_mode is Implicit
end
B200Jacobian() = B200Jacobian(1)
CA.jacobian_cache(::B200Jacobian, Y, atmos) = (; ctx = Ref{Ptr{Cvoid}}(C_NULL))   # filled by the first Wfact (ldiv! gets no `p`)
function CA.update_jacobian!(::B200Jacobian, cache, Y, p, dtγ, t)    # Wfact, implicit/jacobian.jl:74
    cache.ctx[] = ctxptr(p)
    check(ccall((:b200_wfact, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Ptr{Cvoid}),
                ctxptr(p), dptr(Y.c), dptr(Y.f), Float64(dtγ), Float64(t), stream()), "b200_wfact", ctxptr(p))
end
CA.invert_jacobian!(::B200Jacobian, cache, ΔY, R) =                  # ldiv!, implicit/jacobian.jl:78
    check(ccall((:b200_ldiv, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                cache.ctx[], dptr(ΔY.c), dptr(ΔY.f), dptr(R.c), dptr(R.f), stream()), "b200_ldiv", cache.ctx[])

# ---- optional: the native stepper and the fused implicit stage ----------------------------------------------------
"One ARS343 step of the whole dycore in place (24 launches, CUDA-graph replay); replaces CTS.step! (src/simulation/solve.jl:62,125)."
step_ars343!(Y, p, t; fused = true) =
    check(ccall((:b200_step_ars343, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Int32, Ptr{Cvoid}),
                ctxptr(p), dptr(Y.c), dptr(Y.f), Float64(t), Int32(fused), stream()), "b200_step_ars343", ctxptr(p))
"N = U − J(U)⁻¹ R(U) (+ T_post_imp! correction): one Newton iteration of the implicit stage as one kernel."
implicit_stage!(N, U, p, dtγ) =
    check(ccall((:b200_implicit_stage, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}),
                ctxptr(p), dptr(N.c), dptr(N.f), dptr(U.c), dptr(U.f), Float64(dtγ), stream()), "b200_implicit_stage", ctxptr(p))

# ---- helpers used by `create` ---------------------------------------------------------------------------------------
"""
    extract_dxdxi_J(hspace, Nq, elems_ext) -> (dxdxi, J2, lat)

∂x/∂ξ (row-major 2×2 per node: [a*2 + b] = ∂x_a/∂ξ_b in the local east/north basis), the horizontal Jacobian and the latitude
(degrees) for the elements `elems_ext` (0-based global ids: local elements, then ghosts) of the serial horizontal space, in (j, i)
order per element — what `b200_geometry` asks for.  `Fields.local_geometry_field(space).∂x∂ξ` is an Axis2Tensor whose components
are an SMatrix stored column-major: data.:1 = [1,1], :2 = [2,1], :3 = [1,2], :4 = [2,2]  [UPSTREAM-RECALL].
"""
function extract_dxdxi_J(hspace, Nq, elems_ext)
    lgf = Fields.local_geometry_field(hspace)
    J = Array(parent(lgf.J))                                  # (Nq, Nq, 1, Nh) — IJFH
    M = lgf.∂x∂ξ.components.data
    a11, a21, a12, a22 = (Array(parent(getproperty(M, k))) for k in 1:4)
    φ = Array(parent(Fields.coordinate_field(hspace).lat))
    n = length(elems_ext)
    dxdxi = Vector{Float64}(undef, 4 * Nq * Nq * n)
    J2 = Vector{Float64}(undef, Nq * Nq * n)
    lat = similar(J2)
    for (s, g) in enumerate(elems_ext), j in 1:Nq, i in 1:Nq
        h = g + 1
        k = ((s - 1) * Nq + (j - 1)) * Nq + i
        J2[k] = J[i, j, 1, h]
        lat[k] = φ[i, j, 1, h]
        dxdxi[4k - 3] = a11[i, j, 1, h]
        dxdxi[4k - 2] = a12[i, j, 1, h]
        dxdxi[4k - 1] = a21[i, j, 1, h]
        dxdxi[4k] = a22[i, j, 1, h]
    end
    return dxdxi, J2, lat
end

"""
    vertical_jacobians(space) -> (z_c, z_f, dz_c, dz_f)

Levels and vertical ∂z/∂ξ³ of one column at centres and faces (flat grid: identical in every column).  On the extruded space the
3-D Jacobian is J = J2 · ((R+z)/R)² · ∂z/∂ξ³ (ClimaCore `product_geometry`), so the vertical factor is read from the vertical
finite-difference grid rather than divided out: Δz of the cells at centres, centre-to-centre spacing at interior faces and twice the
half-cell at the two boundary faces (the convention of ClimaCore's face LocalGeometry [UPSTREAM-RECALL], restated in grid.py).
"""
function vertical_jacobians(space)
    zc = Float64.(Array(parent(Fields.coordinate_field(Spaces.center_space(space)).z))[:, 1, 1, 1, 1])
    zf = Float64.(Array(parent(Fields.coordinate_field(Spaces.face_space(space)).z))[:, 1, 1, 1, 1])
    dzc = zf[2:end] .- zf[1:(end - 1)]
    dzf = similar(zf)
    dzf[2:(end - 1)] .= zc[2:end] .- zc[1:(end - 1)]
    dzf[1] = 2 * (zc[1] - zf[1])
    dzf[end] = 2 * (zf[end] - zc[end])
    return zc, zf, dzc, dzf
end

"""
    halo_plan(gtopo, rank, nranks)

Partition of the GLOBAL (serial) Topology2D tables for `rank`: the same algorithm as climaatmos.jl_b200/partition.py
(`partition_grid`), whose output the library is tested with — contiguous near-equal ranges of the space-filling-curve order
(the first `nelems % nranks` ranks get one extra element), ghosts = elements of other ranks sharing a vertex with a local one,
ordered by (owner rank, global id), send lists in ascending global id.  All ids returned are 0-based; tables are flattened
row-major as `b200_topology` expects.
"""
function halo_plan(gtopo, rank, nranks)
    nel = Topologies.nlocalelems(gtopo)
    base, extra = divrem(nel, nranks)
    starts = [r * base + min(r, extra) for r in 0:nranks]
    lo, hi = starts[rank + 1], starts[rank + 2]                        # local global ids lo:(hi-1)
    islocal(e) = lo <= e < hi
    owner(e) = searchsortedlast(starts, e) - 1
    lvs = [(Int(e) - 1, Int(v) - 1) for (e, v) in gtopo.local_vertices]
    off = Int.(gtopo.local_vertex_offset) .- 1                         # 0-based offsets, length nverts + 1
    nverts = length(off) - 1
    keep = Int[]
    ghosts = Set{Int}()
    for v in 1:nverts
        mem = [lvs[q][1] for q in (off[v] + 1):off[v + 1]]
        if any(islocal, mem)
            push!(keep, v)
            union!(ghosts, filter(!islocal, mem))
        end
    end
    ghost_list = sort(collect(ghosts); by = e -> (owner(e), e))
    owners = owner.(ghost_list)
    nbrs = sort(unique(owners))
    recv_offset = Int32[0]
    for r in nbrs
        push!(recv_offset, recv_offset[end] + count(==(r), owners))
    end
    elems_ext = vcat(collect(lo:(hi - 1)), ghost_list)
    g2l = Dict(g => k - 1 for (k, g) in enumerate(elems_ext))
    send = Dict(r => Set{Int}() for r in nbrs)
    for v in keep
        mem = [lvs[q][1] for q in (off[v] + 1):off[v + 1]]
        mine = filter(islocal, mem)
        for e in mem
            islocal(e) || union!(send[owner(e)], mine)
        end
    end
    send_offset = Int32[0]
    send_elems = Int32[]
    for r in nbrs
        append!(send_elems, Int32[g2l[e] for e in sort(collect(send[r]))])
        push!(send_offset, length(send_elems))
    end
    faces = Int32[]
    for (e1, f1, e2, f2, rev) in Topologies.interior_faces(gtopo)
        a, b = Int(e1) - 1, Int(e2) - 1
        if islocal(a) || islocal(b)
            append!(faces, Int32[g2l[a], f1 - 1, g2l[b], f2 - 1, rev ? 1 : 0])
        end
    end
    lv = Int32[]
    lvo = Int32[0]
    for v in keep
        for q in (off[v] + 1):off[v + 1]
            append!(lv, Int32[g2l[lvs[q][1]], lvs[q][2]])
        end
        push!(lvo, length(lv) ÷ 2)
    end
    return (; nh = hi - lo, nh_ghost = length(ghost_list), elems_ext, interior_faces = faces, local_vertices = lv,
            local_vertex_offset = lvo, neighbor_ranks = Int32.(nbrs), send_offset, send_elems, recv_offset,
            elem_gid = Int64.(elems_ext))
end

"""
    setup_peer_halo!(ctx, comms, plan)

NVLink peer-memory DSS halo: every rank exports the cudaIpc handle of its ghost buffer (b200_halo_export), the 64-byte handles and
each rank's (neighbour list, receive offsets, ghost count) are all-gathered through ClimaComms, and b200_halo_import maps the
neighbours' buffers.  Mirrors capi.py:setup_peer_halo.  If mapping fails on any rank the library keeps the NCCL send/recv halo.
"""
function setup_peer_halo!(ctx, comms, plan)
    rank = ClimaComms.mypid(comms) - 1
    handle = zeros(UInt8, 64)
    check(ccall((:b200_halo_export, lib), Cint, (Ptr{Cvoid}, Ptr{UInt8}), ctx.ptr, handle), "b200_halo_export", ctx.ptr)
    MAXN = 32                                                           # b200_create bounds n_neighbors by 32
    meta = fill(Int32(-1), 2 + 2 * MAXN)                                # nh_ghost, n_neighbors, neighbour ranks, their receive offsets
    nn = length(plan.neighbor_ranks)
    meta[1] = plan.nh_ghost
    meta[2] = nn
    meta[3:(2 + nn)] .= plan.neighbor_ranks
    meta[(3 + MAXN):(2 + MAXN + nn)] .= plan.recv_offset[1:nn]
    all_handles = ClimaComms.allgather(comms, handle)                   # 64·nranks bytes, rank-major
    all_meta = ClimaComms.allgather(comms, meta)
    M(q) = all_meta[(q * length(meta) + 1):((q + 1) * length(meta))]
    handles = UInt8[]
    their_off = Int32[]
    their_nhg = Int32[]
    for q in plan.neighbor_ranks
        m = M(q)
        k = findfirst(==(Int32(rank)), m[3:(2 + m[2])])                 # my position in neighbour q's neighbour list
        append!(handles, all_handles[(q * 64 + 1):((q + 1) * 64)])
        push!(their_off, m[2 + MAXN + k])
        push!(their_nhg, m[1])
    end
    rc = GC.@preserve handles their_off their_nhg ccall((:b200_halo_import, lib), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Ptr{Int32}, Ptr{Int32}),
                                                        ctx.ptr, handles, their_off, their_nhg)
    ok = ClimaComms.allreduce(comms, rc == 0 ? 1 : 0, min) == 1
    ok || @warn "b200 peer-memory halo unavailable; using the NCCL send/recv halo"
    return ok
end

end # module
