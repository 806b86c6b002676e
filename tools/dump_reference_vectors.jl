# dump_reference_vectors.jl — produce the golden vectors that PIN the oracle (oracle/dycore_oracle.py) and the CUDA path to the
# real ClimaAtmos / ClimaCore / ClimaTimeSteppers stack (SURVEY.md §8c "authoritative oracle plan", VERDICT r1 item 2).
#
# Neither the build container nor the GPU box has julia (profiles/r2_gpu_box_probe.txt), so this script has NOT been executed by
# the builder: it is the committed recipe a maintainer with the pinned environment (.buildkite/Manifest-v1.11.toml, Julia 1.11,
# ClimaAtmos v0.42.7) runs once:
#
#     julia --project=.buildkite tools/dump_reference_vectors.jl /tmp/ref_vectors
#     python tools/ref_vectors_to_npz.py /tmp/ref_vectors tests/golden        # → tests/golden/ref_<case>.npz
#
# after which tests/test_reference_fixtures.py stops being a strict xfail and compares, for every case: the space-filling-curve
# element order and Topology2D tables (bit-exact), the horizontal LocalGeometry, the initial state, every hook's output on the same
# input (cache_imp!, T_exp_T_lim!, T_imp!, Wfact+ldiv!, T_post_imp!, dss!) and Y after one and two CTS steps.
#
# Output format (no HDF5/NPZ dependency on the Python side): one raw little-endian binary file per array plus index.json
# {name: {file, dtype, shape (Julia column-major order)}}.  Parent arrays are VIJFH: (Nv, Nq, Nq, Nf, Nh).
#
# Only public API of the pinned versions is used; the few accessors that are ClimaCore internals are marked [UPSTREAM-RECALL].
import ClimaComms
ClimaComms.@import_required_backends
import ClimaAtmos as CA
import ClimaCore: Fields, Spaces, Topologies, Geometry, Quadratures
import ClimaTimeSteppers as CTS
import LinearAlgebra
import Random

const OUT = length(ARGS) ≥ 1 ? ARGS[1] : "ref_vectors"

# ---- the cases: small enough for the NumPy oracle to replay in seconds; YAML keys as in SURVEY.md Appendix B ----------------------
common = Dict(
    "config" => "sphere", "nh_poly" => 3, "ode_algo" => "ARS343", "hyperdiff" => "Hyperdiffusion",
    "implicit_diffusion" => false, "approximate_linear_solve_iters" => 1, "max_newton_iters_ode" => 1,
    "initial_condition" => "DryBaroclinicWave", "disable_surface_flux_tendency" => true, "deep_atmosphere" => true,
    "output_default_diagnostics" => false, "diagnostics" => [], "dt_save_state_to_disk" => "Inf", "log_progress" => false,
    "device" => "CPUSingleThreaded", "t_end" => "1hours",
)
cases = Dict(
    # shape of configs[0] (numerics_sphere_he6ze10.yml) with the dry baroclinic wave
    "he4ze10_f64" => merge(common, Dict("h_elem" => 4, "z_elem" => 10, "z_max" => 30000.0, "dz_bottom" => 500.0, "dt" => "400secs",
                                        "rayleigh_sponge" => false, "viscous_sponge" => false, "FLOAT_TYPE" => "Float64")),
    "he6ze10_f32" => merge(common, Dict("h_elem" => 6, "z_elem" => 10, "z_max" => 30000.0, "dz_bottom" => 500.0, "dt" => "400secs",
                                        "rayleigh_sponge" => false, "viscous_sponge" => false, "FLOAT_TYPE" => "Float32")),
    # vertical grid and sponges of the he16/he30 ze63 configs (B2/B3), zd = 40 km as toml/longrun_held_suarez.toml
    "he3ze63_f64" => merge(common, Dict("h_elem" => 3, "z_elem" => 63, "z_max" => 60000.0, "dz_bottom" => 30.0, "dt" => "120secs",
                                        "rayleigh_sponge" => true, "viscous_sponge" => true, "FLOAT_TYPE" => "Float64",
                                        "toml" => ["toml/longrun_held_suarez.toml"])),
    # BASELINE configs[2] in small: the 0M-moist baroclinic wave (EquilibriumMicrophysics0M, thermodynamically active ρq_tot).  The
    # precipitation sink of the 0M scheme is a parameterised tendency outside the dycore hooks dumped here (remaining_tendency! contains
    # it through additional_tendency!: the oracle comparison of that hook therefore needs `precip_model`-free settings or a tolerance on
    # ρq_tot/ρe_tot/ρ where q_liq + q_ice > 0; every other hook is unaffected).
    "he4ze10_moist_f64" => merge(common, Dict("h_elem" => 4, "z_elem" => 10, "z_max" => 30000.0, "dz_bottom" => 500.0, "dt" => "400secs",
                                              "rayleigh_sponge" => false, "viscous_sponge" => false, "FLOAT_TYPE" => "Float64",
                                              "initial_condition" => "MoistBaroclinicWave", "microphysics_model" => "0M")),
)

# ---- raw binary writer -------------------------------------------------------------------------------------------------------------
index = Dict{String, Any}()
function put!(case, name, a::AbstractArray)
    a = Array(a)
    file = "$(case)__$(name).bin"
    open(joinpath(OUT, file), "w") do io
        write(io, htol.(a))
    end
    index["$(case)/$(name)"] = Dict("file" => file, "dtype" => string(eltype(a)), "shape" => collect(size(a)))
    return nothing
end
put!(case, name, x::Number) = put!(case, name, [x])
parentarray(f) = Array(parent(Fields.field_values(f)))
function put_state!(case, name, Y)
    put!(case, name * "_c", parentarray(Y.c))   # (Nv, 4, 4, Nf = 4: ρ, uₕ₁, uₕ₂, ρe_tot, Nh)  [prognostic_variables.jl:54-61]
    put!(case, name * "_f", parentarray(Y.f))   # (Nv+1, 4, 4, 1, Nh)
end

function dump_case(case, dict)
    Random.seed!(1234)
    config = CA.AtmosConfig(dict; job_id = "dump_$case")
    simulation = CA.get_simulation(config)
    integrator = simulation.integrator
    Y, p, t = integrator.u, integrator.p, integrator.t
    dt = integrator.dt
    FT = eltype(Y)

    # -- grid: SFC element order, Topology2D tables, horizontal local geometry, quadrature, vertical grid (row c / R2 of SURVEY)
    hspace = Spaces.horizontal_space(axes(Y.c))
    topo = Spaces.topology(hspace)
    put!(case, "elemorder", Int32.(reduce(hcat, [collect(Tuple(ci)) for ci in topo.elemorder])))       # [UPSTREAM-RECALL] field name
    put!(case, "interior_faces", Int32.(reduce(hcat, [collect(f) for f in Topologies.interior_faces(topo)])))  # (5, n): e1 f1 e2 f2 rev
    put!(case, "local_vertices", Int32.(reduce(hcat, [collect(v) for v in topo.local_vertices])))       # (2, n): elem, vert
    put!(case, "local_vertex_offset", Int32.(topo.local_vertex_offset))
    lg = Spaces.local_geometry_data(hspace)
    put!(case, "h_local_geometry", Array(parent(lg)))                                                 # (Nq, Nq, ncomp, Nh): coords, J, WJ, ∂x∂ξ, …
    put!(case, "h_local_geometry_fields", UInt8.(collect(string(propertynames(lg)))))
    coords = Fields.coordinate_field(hspace)
    put!(case, "lat", Array(parent(coords.lat)))
    put!(case, "long", Array(parent(coords.long)))
    quad = Spaces.quadrature_style(hspace)
    ξ, w = Quadratures.quadrature_points(Float64, quad)
    put!(case, "gll_points", collect(ξ)); put!(case, "gll_weights", collect(w))
    put!(case, "gll_D", collect(Quadratures.differentiation_matrix(Float64, quad)))
    put!(case, "z_c", Array(parent(Fields.coordinate_field(Spaces.center_space(axes(Y.c))).z))[:, 1, 1, 1, 1])
    put!(case, "z_f", Array(parent(Fields.coordinate_field(Spaces.face_space(axes(Y.f))).z))[:, 1, 1, 1, 1])
    put!(case, "J_c", parentarray(Fields.local_geometry_field(Y.c).J))
    put!(case, "J_f", parentarray(Fields.local_geometry_field(Y.f).J))
    put!(case, "dt", Float64(float(dt)))
    put!(case, "node_horizontal_length_scale", Float64(Spaces.node_horizontal_length_scale(hspace)))
    if !isnothing(p.atmos.numerics.hyperdiff)
        ν = CA.ν₄(p.atmos.numerics.hyperdiff, Y)                                                       # hyperdiffusion.jl:21-28
        put!(case, "nu4_vorticity", Float64(ν.ν₄_vorticity)); put!(case, "nu4_scalar", Float64(ν.ν₄_scalar))
    end

    # -- initial state and the hooks on it, in the order a CTS implicit stage calls them (integrator.jl:190-225)
    put_state!(case, "Y0", Y)
    U = similar(Y); U .= Y
    CA.set_implicit_precomputed_quantities!(U, p, t)                                                   # cache_imp! (also filters U.f.u₃)
    put_state!(case, "cache_imp_Y", U)
    for name in (:ᶜK, :ᶜT, :ᶜp, :ᶜh_tot, :ᶠu³)
        put!(case, "precomputed_" * String(name), parentarray(getproperty(p.precomputed, name)))
    end
    Yₜ = similar(Y); Yₜ_lim = similar(Y)
    CA.remaining_tendency!(Yₜ, Yₜ_lim, U, p, t)                                                        # T_exp_T_lim!
    put_state!(case, "t_exp", Yₜ); put_state!(case, "t_lim", Yₜ_lim)
    CA.implicit_tendency!(Yₜ, U, p, t)                                                                 # T_imp!
    put_state!(case, "t_imp", Yₜ)
    γ = 0.4358665215084590
    dtγ = FT(float(dt) * γ)
    jac = integrator.cache.newtons_method_cache.j                                                      # [UPSTREAM-RECALL] the Jacobian object CTS holds
    CA.update_jacobian!(jac, U, p, dtγ, t)                                                             # Wfact
    R = similar(Y); Random.seed!(99)
    R.c .= Y.c .* FT(1e-3); R.f .= Y.f .* FT(1e-3) .+ one(FT) .* Geometry.Covariant3Vector(FT(1e-3))   # a reproducible right-hand side
    put_state!(case, "ldiv_R", R)
    ΔY = similar(Y)
    LinearAlgebra.ldiv!(ΔY, jac, R)                                                                    # ldiv!
    put_state!(case, "ldiv_dY", ΔY)
    if p.atmos.numerics.energy_q_tot_upwinding != Val(:none)
        CA.correct_implicit_advection_tendency!(Yₜ, U, p, t)                                           # T_post_imp!
        put_state!(case, "t_post_imp", Yₜ)
    end
    V = similar(Y); V .= Yₜ .* dtγ .+ Y
    CA.dss!(V, p, t)                                                                                   # dss!
    put_state!(case, "dss_in", Yₜ .* dtγ .+ Y); put_state!(case, "dss_out", V)

    # -- Y after one and two full CTS steps (stage order, T_imp formation point, lim!/dss! placement: ADVICE r1 item 2)
    CTS.step!(integrator)
    put_state!(case, "Y1", integrator.u)
    CTS.step!(integrator)
    put_state!(case, "Y2", integrator.u)
    return nothing
end

mkpath(OUT)
for (case, dict) in cases
    @info "dumping $case"
    dump_case(case, dict)
end
# index.json without a JSON dependency
open(joinpath(OUT, "index.json"), "w") do io
    entries = ["  \"$k\": {\"file\": \"$(v["file"])\", \"dtype\": \"$(v["dtype"])\", \"shape\": [$(join(v["shape"], ", "))]}" for (k, v) in index]
    write(io, "{\n" * join(entries, ",\n") * "\n}\n")
end
@info "wrote $(length(index)) arrays to $OUT"
