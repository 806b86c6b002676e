"""Parameter and numerics bundles consumed by the dycore (POD mirrors of the reference's types).

``DycoreParams`` is the subset of ``ClimaAtmosParameters`` (src/parameters/Parameters.jl:541-568,
src/parameters/create_parameters.jl:200-224) the dry/tracer dycore reads.  Default values are the
ClimaParams 1.1.4 defaults [UPSTREAM-RECALL]; in a drop-in deployment they are read from the live
``params`` object and passed across the C-ABI (include/b200_dycore.h: ``b200_params``).
"""
from __future__ import annotations

import dataclasses


@dataclasses.dataclass
class DycoreParams:
    R_d: float = 8.3144598 / 0.02897
    kappa_d: float = 2.0 / 7.0
    T_0: float = 273.16
    grav: float = 9.81
    Omega: float = 7.2921159e-5
    planet_radius: float = 6.371e6
    MSLP: float = 1.01325e5
    p_ref_theta: float = 1.0e5
    T_surf_ref: float = 290.0
    T_min_ref: float = 220.0
    T_min_sgs: float = 150.0  # model_getters.jl:224 fallback value
    # sponges (types.jl:782-866)
    zd_rayleigh: float = 15000.0
    alpha_rayleigh_uh: float = 0.0
    alpha_rayleigh_w: float = 1.0
    zd_viscous: float = 15000.0
    kappa_2_sponge: float = 1.0e6
    # Held–Suarez forcing (held_suarez.jl:188-237; ClimaParams defaults [UPSTREAM-RECALL])
    day: float = 86400.0
    sigma_b: float = 0.7
    dT_y_dry: float = 60.0
    T_equator_dry: float = 315.0
    dtheta_z: float = 10.0
    T_min_hs: float = 200.0
    # vertical diffusion (create_parameters.jl:327-331 vert_diff_params: C_E, H_diffusion → H, D_0_diffusion → D₀;
    # values [UPSTREAM-RECALL], cf. toml/rcemipii_box.toml:61-65, toml/bomex_box_rhoe.toml:1-2)
    C_E: float = 0.0044
    H_diffusion: float = 7000.0
    D_0_diffusion: float = 1.0
    # moist thermodynamics (EquilibriumMicrophysics0M; docs/src/thermodynamics.md:60-150; Thermodynamics.jl 1.3.0 parameter names,
    # ClimaParams defaults [UPSTREAM-RECALL]).  Calorically perfect constituents, reference temperature T_0 = triple point.
    R_v: float = 8.3144598 / 0.01801528
    cp_v: float = 1859.0
    cp_l: float = 4181.0
    cp_i: float = 2100.0
    LH_v0: float = 2.5008e6
    LH_s0: float = 2.8344e6
    T_triple: float = 273.16
    press_triple: float = 611.657
    T_freeze: float = 273.15
    T_icenuc: float = 233.0
    pow_icenuc: float = 1.0

    @property
    def cp_d(self):
        return self.R_d / self.kappa_d

    @property
    def cv_d(self):
        return self.cp_d - self.R_d

    @property
    def cv_v(self):
        return self.cp_v - self.R_v

    @property
    def LH_f0(self):
        return self.LH_s0 - self.LH_v0

    @property
    def e_int_v0(self):  # internal energy of vapour at T_0 (liquid water is the zero)
        return self.LH_v0 - self.R_v * self.T_0

    @property
    def e_int_i0(self):
        return self.LH_f0


@dataclasses.dataclass
class DycoreNumerics:
    """``AtmosNumerics`` + model switches relevant to the path (types.jl:1850, :499-503)."""

    dt: float = 400.0
    hyperdiff: bool = True
    nu4_vorticity_coeff: float = 0.1857  # default_config.yml:30-32
    prandtl_number: float = 0.2  # default_config.yml:33-35 (ν₄_scalar = ν₄_vorticity / Pr)
    divergence_damping_factor: float = 5.0  # default_config.yml:218-220
    rayleigh_sponge: bool = False
    viscous_sponge: bool = False
    energy_upwinding: str = "vanleer_limiter"  # default_config.yml:324-326
    tracer_upwinding: str = "vanleer_limiter"  # default_config.yml:321-323
    apply_sem_quasimonotone_limiter: bool = False  # default_config.yml (Limiters.QuasiMonotoneLimiter in lim!, type_getters.jl:129)
    held_suarez: bool = False
    # default_config.yml:190-198: None | "vertical_water_borrowing" (Limiters.VerticalMassBorrowingLimiter in lim!, cache.jl:216-219)
    tracer_nonnegativity_method: str | None = None
    # vertical diffusion: vert_diff ∈ {None, "VerticalDiffusion", "DecayWithHeightDiffusion"}
    # (default_config.yml:166-168, model_getters.jl:332-358); implicit_diffusion → diff_mode (type_getters.jl:131);
    # approximate_linear_solve_iters (default_config.yml:397-402); momentum diffusion is disabled for Held–Suarez runs
    # (type_getters.jl:46)
    vert_diff: str | None = None
    implicit_diffusion: bool = False
    approximate_linear_solve_iters: int = 1
    disable_momentum_vertical_diffusion: bool = False
    # microphysics_model: None (DryModel) | "0M" (EquilibriumMicrophysics0M: ρq_tot is component 4 of Y.c, thermodynamically active,
    # condensate diagnosed by saturation adjustment; the 0M precipitation sink itself is a parameterised tendency outside the dycore,
    # BASELINE.json configs[2] "dycore + tracer advection only")
    microphysics_model: str | None = None


# ARS343 tableau (Ascher–Ruuth–Spiteri 1997 §2.7), as used by ClimaTimeSteppers' IMEXAlgorithm
def ars343():
    g = 0.4358665215084590
    a42 = 0.5529291480359398
    a43 = a42
    b1 = -3 * g**2 / 2 + 4 * g - 1 / 4
    b2 = 3 * g**2 / 2 - 5 * g + 5 / 4
    a31 = (
        (1 - 9 * g / 2 + 3 * g**2 / 2) * a42
        + (11 / 4 - 21 * g / 2 + 15 * g**2 / 4) * a43
        - 7 / 2
        + 13 * g
        - 9 * g**2 / 2
    )
    a32 = (
        (-1 + 9 * g / 2 - 3 * g**2 / 2) * a42
        + (-11 / 4 + 21 * g / 2 - 15 * g**2 / 4) * a43
        + 4
        - 25 * g / 2
        + 9 * g**2 / 2
    )
    a41 = 1 - a42 - a43
    a_exp = [[0, 0, 0, 0], [g, 0, 0, 0], [a31, a32, 0, 0], [a41, a42, a43, 0]]
    a_imp = [[0, 0, 0, 0], [0, g, 0, 0], [0, (1 - g) / 2, g, 0], [0, b1, b2, g]]
    b_exp = [0, b1, b2, g]
    b_imp = [0, b1, b2, g]
    return a_exp, a_imp, b_exp, b_imp, g
