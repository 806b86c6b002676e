# B200Dycore.jl — Julia glue between ClimaAtmos and libb200dycore.so (include/b200_dycore.h).
#
# UNTESTED in the build image (no `julia`, ClimaCore / ClimaTimeSteppers not vendored): this file is the binding a
# ClimaAtmos maintainer drops into `src/`; the same calls, in the same order and with the same pointer conventions,
# are exercised by the Python twin `climaatmos.jl_b200/capi.py` (ctypes) in tests/test_gpu_parity.py, and
# tests/test_grid_and_host_logic.py::test_ctypes_mirrors_match_the_c_structs pins the struct layouts mirrored below.
# Places that depend on ClimaCore accessor names are marked [UPSTREAM-RECALL ClimaCore 0.15.1].
#
# Every hook has the name and signature of the reference method it replaces (file:line in the comments), mutates its
# first argument(s) and is asynchronous on the current CUDA stream.
module B200Dycore

import ClimaCore: Fields, Spaces, Topologies, Quadratures, Geometry
import ClimaComms, CUDA, LinearAlgebra
import ..ClimaAtmos as CA
import ..ClimaAtmos.Parameters as CAP

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
    energy_upwinding::Int32      # 0 none, 1 first_order, 3 vanleer_limiter
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
end
struct CachePtrs
    u_c::Ptr{Cvoid}; u3_f::Ptr{Cvoid}; K_c::Ptr{Cvoid}; T_c::Ptr{Cvoid}; p_c::Ptr{Cvoid}; h_tot_c::Ptr{Cvoid}
end

mutable struct Ctx
    ptr::Ptr{Cvoid}
    keep::Any   # host arrays handed to b200_create (it copies them; kept only until create returns)
end

check(rc, what, ctx = C_NULL) = rc == 0 || error("$what: " * unsafe_string(ccall((:b200_last_error, lib), Cstring, (Ptr{Cvoid},), ctx)))
# VIJFH parent array of a field on the device: (Nv, 4, 4, Nf, Nh), level fastest
dptr(f) = reinterpret(Ptr{Cvoid}, pointer(parent(Fields.field_values(f))))
stream() = reinterpret(Ptr{Cvoid}, CUDA.stream().handle)
# 0 none, 1 first_order, 2 third_order (b200_create rejects it: not built), 3 vanleer_limiter
upw(x) = x == Val(:none) ? Int32(0) : x == Val(:first_order) ? Int32(1) : x == Val(:third_order) ? Int32(2) : Int32(3)

"""
    create(Y, p; approximate_solve_iters = 1) -> Ctx

Called once at the end of `build_cache` (src/cache/cache.jl:165-305): copies geometry, connectivity and parameters
from the live ClimaCore objects into the library.  [UPSTREAM-RECALL] for the ClimaCore accessors.
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
    nh = Topologies.nlocalelems(topo)
    ng = Topologies.nghostelems(topo)
    # horizontal local geometry on the host: ∂x/∂ξ (2×2, local east/north basis), J, latitude, in (j, i) order per element
    lg = Array(parent(Spaces.local_geometry_data(hspace)))             # (Nq, Nq, ncomponents, nh + ng)
    coords = Fields.coordinate_field(hspace)
    lat = Float64.(Array(parent(coords.lat)))[:]
    dxdxi, J2 = extract_dxdxi_J(lg, Nq, nh + ng)                        # Float64 host vectors, see the helper below
    _, w = Quadratures.quadrature_points(Float64, quad)
    D = Quadratures.differentiation_matrix(Float64, quad)              # D[i, k] = l'_k(ξ_i)
    zc = Float64.(Array(parent(Fields.coordinate_field(Spaces.center_space(space)).z))[:, 1, 1, 1, 1])
    zf = Float64.(Array(parent(Fields.coordinate_field(Spaces.face_space(space)).z))[:, 1, 1, 1, 1])
    dzc, dzf = vertical_jacobians(space)                               # ∂z/∂ξ³ of one column at centres / faces
    gll_w = collect(Float64, w); gll_D = collect(Float64, permutedims(D))   # row-major D[i][k] for C
    # Topology2D tables, 0-based, ghost element ids = nh + ghost slot
    faces = Int32[]; for (e1, f1, e2, f2, rev) in Topologies.interior_faces(topo); append!(faces, Int32[e1 - 1, f1 - 1, e2 - 1, f2 - 1, rev]); end
    lv = Int32[]; for (e, v) in topo.local_vertices; append!(lv, Int32[e - 1, v - 1]); end
    lvo = Int32.(topo.local_vertex_offset .- 1)
    nbr, soff, selems, roff, gid = halo_plan(topo)                     # neighbour ranks, send/recv lists, global ids
    params = p.params
    ν₄ᵥ, ν₄ₛ = p.atmos.hyperdiff === nothing ? (0.0, 0.0) : Float64.(CA.ν₄(p.atmos.hyperdiff, Y))   # hyperdiffusion.jl:21-28
    ntr = count(CA.is_tracer_var, propertynames(Y.c))
    dims = Dims(nh, ng, Nv, Nq, sizeof(FT), p.atmos.numerics.deep_atmosphere ? 1 : 0, ntr)
    geo = GeometryC(pointer(dxdxi), pointer(J2), pointer(lat), pointer(gll_w), pointer(gll_D), pointer(zc), pointer(zf),
                    pointer(dzc), pointer(dzf), Float64(CAP.planet_radius(params)), zf[end])
    topo_c = TopologyC(pointer(faces), length(faces) ÷ 5, pointer(lv), pointer(lvo), length(lvo) - 1, length(nbr),
                       pointer(nbr), pointer(soff), pointer(selems), pointer(roff), pointer(gid))
    rs, vs = p.atmos.rayleigh_sponge, p.atmos.viscous_sponge
    vd = p.atmos.vertical_diffusion
    prm = ParamsC(CAP.R_d(params), CAP.cp_d(params), CAP.cv_d(params), CAP.T_0(params), CAP.grav(params), CAP.Omega(params),
                  CAP.p_ref_theta(params), CAP.T_surf_ref(params), CAP.T_min_ref(params), CAP.T_min_sgs(params),
                  Float64(p.dt), ν₄ᵥ, ν₄ₛ, p.atmos.hyperdiff === nothing ? 1.0 : Float64(p.atmos.hyperdiff.divergence_damping_factor),
                  p.atmos.hyperdiff === nothing ? 0 : 1,
                  rs === nothing ? 0 : 1, rs === nothing ? 0.0 : Float64(rs.zd), rs === nothing ? 0.0 : Float64(rs.α_uₕ), rs === nothing ? 0.0 : Float64(rs.α_w),
                  vs === nothing ? 0 : 1, vs === nothing ? 0.0 : Float64(vs.zd), vs === nothing ? 0.0 : Float64(vs.κ₂),
                  upw(p.atmos.numerics.energy_q_tot_upwinding), upw(p.atmos.numerics.tracer_upwinding),
                  p.atmos.radiation_mode isa CA.RRTMGPI.HeldSuarezForcing ? 1 : 0,    # held_suarez.jl
                  CAP.day(params), CAP.σ_b(params), CAP.ΔT_y_dry(params), CAP.T_equator_dry(params), CAP.Δθ_z(params),
                  CAP.T_min_hs(params), CAP.MSLP(params),
                  p.numerics.sem_quasimonotone_limiter === nothing ? 0 : 1,
                  vd === nothing ? 0 : (vd isa CA.VerticalDiffusion ? 1 : 2),                 # types.jl:564-597
                  p.atmos.diff_mode == CA.Implicit() ? 1 : 0,                                # type_getters.jl:131
                  Int32(approximate_solve_iters), CA.disable_momentum_vertical_diffusion(vd) ? 1 : 0,
                  vd isa CA.VerticalDiffusion ? Float64(vd.C_E) : 0.0,
                  vd isa CA.DecayWithHeightDiffusion ? Float64(vd.H) : 1.0,
                  vd isa CA.DecayWithHeightDiffusion ? Float64(vd.D₀) : 0.0,
                  p.numerics.vertical_water_borrowing_limiter === nothing ? 0 : 1)                # cache.jl:213-219
    comms = ClimaComms.context(Y.c)
    rank, nranks = ClimaComms.mypid(comms) - 1, ClimaComms.nprocs(comms)
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
    nranks > 1 && setup_peer_halo!(ctx, comms, roff, ng)   # b200_halo_export / allgather of the 64-byte handles / b200_halo_import
    finalizer(c -> ccall((:b200_destroy, lib), Cint, (Ptr{Cvoid},), c.ptr), ctx)
    return ctx
end

# ---- the hooks (ClimaODEFunction, src/simulation/integrator.jl:215-225) ------------------------------------------
function remaining_tendency!(Yₜ, Yₜ_lim, Y, p, t)            # src/prognostic_equations/remaining_tendency.jl:48
    check(ccall((:b200_t_exp_lim, lib), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}),
                p.b200.ptr, dptr(Yₜ.c), dptr(Yₜ.f), dptr(Yₜ_lim.c), dptr(Yₜ_lim.f), dptr(Y.c), dptr(Y.f), Float64(t), stream()),
          "b200_t_exp_lim")
    return Yₜ
end
implicit_tendency!(Yₜ, Y, p, t) =                             # implicit/implicit_tendency.jl:36
    check(ccall((:b200_t_imp, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}),
                p.b200.ptr, dptr(Yₜ.c), dptr(Yₜ.f), dptr(Y.c), dptr(Y.f), Float64(t), stream()), "b200_t_imp")
correct_implicit_advection_tendency!(Yₜ, Y, p, t) =           # implicit/implicit_tendency.jl:322 (T_post_imp!)
    check(ccall((:b200_t_post_imp, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}),
                p.b200.ptr, dptr(Yₜ.c), dptr(Yₜ.f), dptr(Y.c), dptr(Y.f), Float64(t), stream()), "b200_t_post_imp")
function set_implicit_precomputed_quantities!(Y, p, t)        # cache/precomputed_quantities.jl:698 (cache_imp!)
    pc = p.precomputed
    cp = CachePtrs(dptr(pc.ᶜu), dptr(pc.ᶠu³), dptr(pc.ᶜK), dptr(pc.ᶜT), dptr(pc.ᶜp), dptr(pc.ᶜh_tot))
    check(ccall((:b200_cache_imp, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{CachePtrs}, Ptr{Cvoid}),
                p.b200.ptr, dptr(Y.c), dptr(Y.f), cp, stream()), "b200_cache_imp")
end
function dss!(Y, p, t)                                         # prognostic_equations/constrain_state.jl:59
    fields = Ptr{Cvoid}[dptr(Y.c), dptr(Y.f)]
    ncomp = Int32[size(parent(Y.c), 4), 1]; is_face = Int32[0, 1]; kind = Int32[2, 0]   # kind 2: (ρ, Covariant12 pair, scalars…)
    GC.@preserve fields ncomp is_face kind check(ccall((:b200_dss, lib), Cint,
        (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Int32, Ptr{Cvoid}),
        p.b200.ptr, fields, ncomp, is_face, kind, Int32(2), stream()), "b200_dss")
end
limiters_func!(Y, p, t, ref_Y) =                              # prognostic_equations/limited_tendencies.jl:64 (lim!)
    check(ccall((:b200_lim, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}),
                p.b200.ptr, dptr(Y.c), dptr(Y.f), dptr(ref_Y.c), dptr(ref_Y.f), Float64(t), stream()), "b200_lim")

# ---- Jacobian: the three-method JacobianAlgorithm contract of implicit/jacobian.jl:16-26 ---------------------------
struct B200Jacobian <: CA.JacobianAlgorithm
    approximate_solve_iters::Int     # as ManualSparseJacobian.approximate_solve_iters (manual_sparse_jacobian.jl:45-60); used when diff_mode is Implicit
end
B200Jacobian() = B200Jacobian(1)
CA.jacobian_cache(::B200Jacobian, Y, atmos) = (; ctx = Ref{Ctx}())    # ctx[] = p.b200, set right after build_cache
CA.update_jacobian!(::B200Jacobian, cache, Y, p, dtγ, t) =           # Wfact, implicit/jacobian.jl:74
    check(ccall((:b200_wfact, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Ptr{Cvoid}),
                p.b200.ptr, dptr(Y.c), dptr(Y.f), Float64(dtγ), Float64(t), stream()), "b200_wfact")
CA.invert_jacobian!(::B200Jacobian, cache, ΔY, R) =                  # ldiv!, implicit/jacobian.jl:78
    check(ccall((:b200_ldiv, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                cache.ctx[].ptr, dptr(ΔY.c), dptr(ΔY.f), dptr(R.c), dptr(R.f), stream()), "b200_ldiv")

# ---- optional: the native stepper and the fused implicit stage ----------------------------------------------------
"One ARS343 step of the whole dycore in place (24 launches, CUDA-graph replay); replaces CTS.step! (src/simulation/solve.jl:62,125)."
step_ars343!(Y, p, t; fused = true) =
    check(ccall((:b200_step_ars343, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Int32, Ptr{Cvoid}),
                p.b200.ptr, dptr(Y.c), dptr(Y.f), Float64(t), Int32(fused), stream()), "b200_step_ars343")
"N = U − J(U)⁻¹ R(U) (+ T_post_imp! correction): one Newton iteration of the implicit stage as one kernel."
implicit_stage!(N, U, p, dtγ) =
    check(ccall((:b200_implicit_stage, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}),
                p.b200.ptr, dptr(N.c), dptr(N.f), dptr(U.c), dptr(U.f), Float64(dtγ), stream()), "b200_implicit_stage")

# ---- helpers whose bodies depend on ClimaCore internals [UPSTREAM-RECALL]; signatures fixed by `create` -----------------
"∂x/∂ξ (row-major 2×2 per node, local east/north basis) and J from the LocalGeometry data of the horizontal space."
function extract_dxdxi_J end
"Vertical ∂z/∂ξ³ of one column at centres and faces (FiniteDifferenceSpace local geometry of the extruded space)."
function vertical_jacobians end
"neighbor_ranks, send_offset, send_elems (0-based local ids), recv_offset (0-based ghost slots), elem_gid (Int64) from Topology2D."
function halo_plan end
"b200_halo_export on every rank, ClimaComms.allgather of the 64-byte cudaIpc handles, b200_halo_import."
function setup_peer_halo! end

end # module
