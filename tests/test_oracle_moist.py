"""CPU tests pinning the MOIST (EquilibriumMicrophysics0M) half of the oracle: the thermodynamic state with saturation adjustment
(precomputed_quantities.jl:735-815), the active ρq_tot transport (implicit_tendency.jl:210-214, :333-338), the moist Jacobian blocks
(manual_sparse_jacobian.jl:653-690, 770-790, 827-831), the water hyperdiffusion / sponge split (hyperdiffusion.jl:148-165, 293-306,
475-484; viscous_sponge.jl:158-199).  Thermodynamics.jl is not vendored: the formulation restated is the one documented in
docs/src/thermodynamics.md:60-150; what pins it here are its own consistency relations."""
import numpy as np
import pytest

from climaatmos_jl_b200 import grid as G, params as prm, setups
from oracle.dycore_oracle import Oracle
from tests.test_oracle_identities import apply_jacobian


def make(he=2, ze=12, zmax=30000.0, dzb=300.0, q_0=0.018, sponge=True, hyperdiff=True, upw="vanleer_limiter", FT=np.float64, z_stretch=True):
    P = prm.DycoreParams(zd_rayleigh=20000.0, zd_viscous=20000.0)
    kw = dict(dz_bottom=dzb) if z_stretch else dict(z_stretch=False)
    g = G.make_sphere_grid(FT=FT, h_elem=he, z_elem=ze, z_max=zmax, radius=P.planet_radius, **kw)
    N = prm.DycoreNumerics(dt=300.0, rayleigh_sponge=sponge, viscous_sponge=sponge, hyperdiff=hyperdiff, energy_upwinding=upw,
                           microphysics_model="0M")
    o = Oracle(g, P, N, FT)
    Yc, Yf = setups.moist_baroclinic_wave(g, P, q_0=q_0)
    return g, P, N, o, Yc, Yf


def apply_jacobian_moist(o, Jm, dc, df):
    outc, outf = apply_jacobian(o, Jm, dc, df)
    x = df[:, 0]
    lo, hi = Jm["rhoq_u3"]
    outc[:, 4] += lo * x[..., :-1] + hi * x[..., 1:]
    pad_lo = lambda a: np.concatenate([0 * a[..., :1], a], -1)
    pad_hi = lambda a: np.concatenate([a, 0 * a[..., :1]], -1)
    lo, hi = Jm["u3_rhoq"]
    outf[:, 0] += lo * pad_lo(dc[:, 4]) + hi * pad_hi(dc[:, 4])
    return outc, outf


def test_saturation_adjustment_solves_the_energy_equation():
    """T, q_liq, q_ice reproduce e_int exactly; unsaturated air keeps q_c = 0; cloudy air sits at q_vap = q_vs(T, ρ) with the
    liquid fraction of the supercooled ramp; ice-only below T_icenuc, liquid-only above T_freeze."""
    g, P, N, o, Yc, Yf = make()
    rng = np.random.default_rng(1234)
    n = 4000
    T_true = rng.uniform(200.0, 310.0, n)
    rho = rng.uniform(2.0e4, 1.02e5, n) / (P.R_d * T_true)
    lam, _ = o.liquid_fraction(T_true)
    qvs = o.q_vap_saturation(T_true, rho, lam)
    qt = qvs * rng.uniform(0.2, 1.8, n)  # from clearly unsaturated to a cloud-water content of 80 % of q_vs
    qc = np.maximum(0.0, qt - qvs)
    e = o.internal_energy(T_true, qt, lam * qc, (1 - lam) * qc)
    T, ql, qi = o.saturation_adjustment(rho, e, qt)
    assert np.abs(T - T_true).max() < 1e-9
    assert np.abs(o.internal_energy(T, qt, ql, qi) - e).max() < 1e-6
    unsat = qt <= qvs
    assert unsat.sum() > 500 and (~unsat).sum() > 500
    assert np.all(ql[unsat] == 0) and np.all(qi[unsat] == 0)
    assert np.allclose((qt - ql - qi)[~unsat], qvs[~unsat], rtol=1e-9)
    cold, warm = (~unsat) & (T_true <= P.T_icenuc), (~unsat) & (T_true >= P.T_freeze)
    assert cold.sum() > 10 and warm.sum() > 10
    assert np.all(ql[cold] == 0) and np.all(qi[warm] == 0)
    # Rankine–Kirchhoff anchors: p_vs(T_triple) = p_triple for both phases; Clausius–Clapeyron slope L(T)/(R_v T²)
    for lam1 in (0.0, 1.0):
        T0 = np.array([P.T_triple])
        assert np.isclose(np.exp(o._ln_pvs(T0, np.array([lam1]))[0])[0], P.press_triple, rtol=1e-14)
    Tt = np.array([285.0])
    one = np.ones(1)
    h = 1e-3
    slope = (o._ln_pvs(Tt + h, one)[0] - o._ln_pvs(Tt - h, one)[0]) / (2 * h)
    L = P.LH_v0 + (P.cp_v - P.cp_l) * (Tt - P.T_0)
    assert np.isclose(slope[0], (L / (P.R_v * Tt * Tt))[0], rtol=1e-8)


def test_moist_cache_is_consistent_and_has_cloudy_points():
    g, P, N, o, Yc, Yf = make(q_0=0.03)  # q_0 raised so that part of the tropical boundary layer is saturated
    pc = o.set_implicit_precomputed_quantities(Yc, Yf)
    rho = Yc[:, 0]
    cloudy = (pc["ql"] + pc["qi"]) > 0
    assert 0 < cloudy.mean() < 0.5
    e_int = Yc[:, 3] / rho - pc["K"] - o.Phi
    assert np.abs(o.internal_energy(pc["T"], pc["qt"], pc["ql"], pc["qi"]) - e_int).max() < 1e-6
    assert np.allclose(pc["p"], rho * pc["Rm"] * pc["T"])
    assert np.allclose(pc["h_tot"], Yc[:, 3] / rho + pc["Rm"] * pc["T"])
    # a drier profile is unsaturated everywhere: no condensate, T is the closed-form inversion
    g, P, N, o, Yc, Yf = make(q_0=0.008)
    pc = o.set_implicit_precomputed_quantities(Yc, Yf)
    assert np.all(pc["ql"] + pc["qi"] == 0)


def test_vanishing_moisture_reproduces_the_dry_model():
    """With ρq_tot ≡ 0 and no hyperdiffusion (whose reference profile q_tot_r(p) is a source of its own) the moist code path is the
    dry one: same tendencies, same Jacobian solve, same step."""
    g, P, N, o, Yc, Yf = make(hyperdiff=False)
    Nd = prm.DycoreNumerics(dt=300.0, rayleigh_sponge=True, viscous_sponge=True, hyperdiff=False)
    od = Oracle(g, P, Nd, np.float64)
    Ycd, Yfd = setups.dry_baroclinic_wave(g, P)
    rng = np.random.default_rng(5)
    Yfd = Yfd + 0.1 * g.dz_f * rng.standard_normal(Yfd.shape)
    od.dss_state(Ycd, Yfd)
    Ycm = np.concatenate([Ycd, np.zeros_like(Ycd[:, :1])], axis=1)
    a, b = od.step(Ycd.copy(), Yfd.copy())
    c, d = o.step(Ycm.copy(), Yfd.copy())
    assert np.abs(c[:, 4]).max() == 0
    for k in range(4):
        assert np.abs(c[:, k] - a[:, k]).max() <= 1e-12 * np.abs(a[:, k]).max()
    assert np.abs(d - b).max() <= 1e-12 * np.abs(b).max()


def test_moist_ldiv_inverts_the_block_matrix_and_matches_a_dense_solve():
    g, P, N, o, Yc, Yf = make()
    rng = np.random.default_rng(7)
    Yf = Yf + 0.2 * g.dz_f * rng.standard_normal(Yf.shape)
    pc = o.set_implicit_precomputed_quantities(Yc, Yf)
    Jm = o.update_jacobian(Yc, Yf, pc, 0.43 * N.dt)
    assert "rhoq_u3" in Jm and "u3_rhoq" in Jm
    Rc, Rf = rng.standard_normal(Yc.shape), rng.standard_normal(Yf.shape)
    dc, df = o.ldiv(Jm, Rc, Rf)
    bc, bf = apply_jacobian_moist(o, Jm, dc, df)
    assert np.abs(bc - Rc).max() < 1e-9 * max(1, np.abs(dc).max()) and np.abs(bf - Rf).max() < 1e-7 * max(1, np.abs(df).max())
    # dense column (independent assembly of the same blocks)
    h, j, i = 3, 1, 2
    M = o.jacobian_dense_column(Jm, h, j, i, ntr=1)
    nv = g.nv
    rhs = np.concatenate([Rc[h, q, j, i] for q in range(5)] + [Rf[h, 0, j, i]])
    x = np.linalg.solve(M, rhs)
    got = np.concatenate([dc[h, q, j, i] for q in range(5)] + [df[h, 0, j, i]])
    assert np.abs(x - got).max() < 1e-9 * np.abs(x).max()


def test_moist_jacobian_matches_finite_differences():
    """Unsaturated smooth state on a fine uniform column: the (u₃, ρq_tot) block — ∂p/∂ρq_tot at fixed ρ, ρe_tot — and the moist κ_m
    in the (u₃, ρ), (u₃, ρe_tot) blocks agree with a centred difference of dtγ·T_imp − Y; the (ρq_tot, u₃) row is linear in u₃."""
    g, P, N, o, Yc, Yf = make(he=2, ze=48, sponge=False, z_stretch=False, q_0=0.008)
    Yf[:, 0, ..., 1:-1] = 0.05 * g.dz_f[1:-1] * np.sin(np.pi * g.z_f[1:-1] / g.z_max)
    dtg = 0.43 * N.dt
    pc = o.set_implicit_precomputed_quantities(Yc, Yf)
    assert np.all(pc["ql"] + pc["qi"] == 0)
    Jm = o.update_jacobian(Yc, Yf, pc, dtg)

    def resid(yc, yf):
        yf = yf.copy()
        p = o.set_implicit_precomputed_quantities(yc, yf)
        tc, tf = o.implicit_tendency(yc, yf, p)
        return dtg * tc - yc, dtg * tf - yf

    k = 6  # low level: q_tot ≈ 1e-2 there
    for comp in (0, 3, 4, "u3"):
        dc, df = np.zeros_like(Yc), np.zeros_like(Yf)
        if comp == "u3":
            df[:, 0, :, :, k] = np.abs(Yf).max()
        else:
            dc[:, comp, :, :, k] = np.abs(Yc[:, comp]).max()
        eps = 1e-6
        rp = resid(Yc + eps * dc, Yf + eps * df)
        rm = resid(Yc - eps * dc, Yf - eps * df)
        fd_f = (rp[1] - rm[1]) / (2 * eps)
        jc, jf = apply_jacobian_moist(o, Jm, dc, df)
        sl = (slice(None), 0, slice(None), slice(None), slice(k - 2, k + 4))
        # columns with q_tot ≈ 0 (poles) sit on the kink of q_tot_nonneg = max(0, ·): a centred difference sees half the slope
        wet = (Yc[:, 4, :, :, k] / Yc[:, 0, :, :, k] > 1e-4)[..., None]
        err = (np.abs(fd_f[sl] - jf[sl]) * wet).max() / np.abs(jf[sl]).max()
        assert err < 2e-2, (comp, err)
        if comp == "u3":
            fd_c = (rp[0] - rm[0]) / (2 * eps)
            for q in (0, 4):
                err = np.abs(fd_c[:, q] - jc[:, q]).max() / np.abs(jc[:, q]).max()
                assert err < 1e-6, (q, err)


def test_water_is_conserved_and_corrected_plus_central_is_upwind():
    g, P, N, o, Yc, Yf = make(he=3)
    rng = np.random.default_rng(11)
    Yf = Yf + 0.2 * g.dz_f * rng.standard_normal(Yf.shape)
    o.dss_state(Yc, Yf)
    pc = o.set_implicit_precomputed_quantities(Yc, Yf)
    tc, tf = o.implicit_tendency(Yc, Yf, pc)
    col = (o.c.J * tc[:, 4]).sum(axis=-1)  # zero-flux boundaries: the column integral of the ρq_tot tendency vanishes
    assert np.abs(col).max() < 1e-12 * (o.c.J * np.abs(tc[:, 4])).sum(axis=-1).max()
    pt, _ = o.correct_implicit_advection_tendency(Yc, Yf, pc)
    up = o.vertical_transport(Yc[:, 0], pc["fu3"], Yc[:, 4] / Yc[:, 0], N.dt, N.energy_upwinding)
    assert np.abs(tc[:, 4] + pt[:, 4] - up).max() < 1e-12 * np.abs(up).max()
    # T_exp_T_lim!: the horizontal terms (advection, hyperdiffusion, sponge) conserve total water and move ρ with it
    Ytc, Ytf, Ylc = o.remaining_tendency(Yc, Yf, pc, with_lim=True)
    tot = Ytc + Ylc
    W = o.c.WJ
    water = (W * tot[:, 4]).sum()
    assert abs(water) < 1e-11 * (W * np.abs(tot[:, 4])).sum()
    assert abs((W * tot[:, 0]).sum()) < 1e-11 * (W * np.abs(tot[:, 0])).sum()
    assert np.abs(Ylc[:, 0]).max() > 0  # the water-mass hyperdiffusion enters Yₜ_lim.ρ


@pytest.mark.parametrize("upw", ["none", "first_order", "vanleer_limiter", "third_order"])
def test_moist_steps_stay_finite_and_conserve(upw):
    g, P, N, o, Yc, Yf = make(he=2, ze=16, q_0=0.025, upw=upw)
    W = o.c.WJ
    m0, w0 = (W * Yc[:, 0]).sum(), (W * Yc[:, 4]).sum()
    for _ in range(3):
        Yc, Yf = o.step(Yc, Yf)
    assert np.isfinite(Yc).all() and np.isfinite(Yf).all()
    assert abs((W * Yc[:, 0]).sum() - m0) / m0 < 1e-12
    assert abs((W * Yc[:, 4]).sum() - w0) / w0 < 1e-12


@pytest.mark.parametrize("name,he,ze,zmax,dzb,dt,sp,lo,hi", [
    ("he4ze10", 4, 10, 30000.0, 500.0, 400.0, False, 1.5e-5, 6e-5), ("he3ze63", 3, 63, 60000.0, 30.0, 120.0, True, 1.5e-5, 6e-5),
    ("he2ze31", 2, 31, 45000.0, 300.0, 200.0, True, 4e-5, 1.6e-4)])
def test_float32_floor_of_the_moist_reference_formulation(name, he, ze, zmax, dzb, dt, sp, lo, hi):
    """The reference's formulas, literally, in Float32 against themselves in Float64 from the same Float32 moist state: after one step
    ρ, uₕ, ρe_tot, ρq_tot agree to ≤ 5e-6, u₃ only to 2.4e-5 … 7.9e-5 (measured: he4ze10 3.0e-5, he3ze63 2.9e-5, he2ze31 7.9e-5).  The
    moist branch has no T_min_sgs floor, so the unbalanced top of the dry tests is absent and u₃ is a smaller field; this is the floor
    the Float32 u₃ bar of tests/test_gpu_moist.py is derived from."""
    P = prm.DycoreParams(zd_rayleigh=0.66 * zmax, zd_viscous=0.66 * zmax)
    g = G.make_sphere_grid(FT=np.float32, h_elem=he, z_elem=ze, z_max=zmax, dz_bottom=dzb, radius=P.planet_radius)
    N = prm.DycoreNumerics(dt=dt, rayleigh_sponge=sp, viscous_sponge=sp, microphysics_model="0M")
    Yc, Yf = setups.moist_baroclinic_wave(g, P, q_0=0.03)
    a = Oracle(g, P, N, np.float64).step(Yc.astype(np.float64), Yf.astype(np.float64))
    b = Oracle(g, P, N, np.float32).step(Yc.copy(), Yf.copy())
    rel = lambda x, y: np.linalg.norm((x.astype(np.float64) - y).ravel()) / np.linalg.norm(y.ravel())
    for k in range(5):
        assert rel(b[0][:, k], a[0][:, k]) < 5e-6, (k, rel(b[0][:, k], a[0][:, k]))
    assert lo < rel(b[1], a[1]) < hi, rel(b[1], a[1])
