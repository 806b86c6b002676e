"""CPU tests pinning the oracle (oracle/dycore_oracle.py) with the structural identities the
reference itself tests or documents — numerical golden data from the reference do
not exist for this path ("parity unpinned")."""
import os
import numpy as np
import pytest

from climaatmos_jl_b200 import grid as G, params as prm, setups
from oracle.dycore_oracle import Oracle


@pytest.fixture(scope="module")
def case():
    P = prm.DycoreParams(zd_rayleigh=20000.0, zd_viscous=20000.0)
    g = G.make_sphere_grid(FT=np.float64, h_elem=3, z_elem=12, z_max=30000.0, dz_bottom=300.0, radius=P.planet_radius)
    N = prm.DycoreNumerics(dt=300.0, rayleigh_sponge=True, viscous_sponge=True)
    o = Oracle(g, P, N, np.float64)
    Yc, Yf = setups.dry_baroclinic_wave(g, P)
    rng = np.random.default_rng(1234)  # reference convention Random.seed!(1234), ci_driver.jl:17-18
    Yc = Yc * (1 + 1e-3 * rng.standard_normal(Yc.shape))
    Yf = 0.3 * g.dz_f * rng.standard_normal(Yf.shape)
    o.dss_state(Yc, Yf)
    return g, P, N, o, Yc, Yf, rng


def test_operator_identities(case):
    """curl∘grad = 0, div∘curl = 0 and weak = −adjoint(strong) (docs/src/discretization.md:76-94)."""
    g, P, N, o, Yc, Yf, rng = case
    f = rng.standard_normal(Yc[:, 0].shape)
    gx, gy = o.grad(f)
    assert np.abs(o.curl3(gx, gy, o.c)).max() < 1e-12 * np.abs(gx).max() / o.c.J.min()
    c1, c2 = o.curl12(f, o.c)
    assert np.abs(o.div(c1, c2, o.c)).max() * o.c.J.max() < 1e-9 * np.abs(f).max()
    # discrete integration by parts, element by element: Σ WJ (φ wdiv(u) + u·grad φ) = 0
    u1, u2 = rng.standard_normal(f.shape), rng.standard_normal(f.shape)
    phi = rng.standard_normal(f.shape)
    px, py = o.grad(phi)
    lhs = (o.c.WJ * (phi * o.wdiv(u1, u2, o.c) + u1 * px + u2 * py)).sum(axis=(1, 2))
    scale = (o.c.WJ * np.abs(u1 * px)).sum(axis=(1, 2))
    assert np.abs(lhs / scale).max() < 1e-12
    # weak gradient / curl adjointness
    w1, w2 = o.wgrad(phi, o.c)
    lhs = (o.c.WJ * (w1 * u1 + w2 * u2 + phi * o.div(u1, u2, o.c))).sum(axis=(1, 2))
    assert np.abs(lhs / scale).max() < 1e-12
    a1, a2 = rng.standard_normal(f.shape), rng.standard_normal(f.shape)
    lhs = (o.c.WJ * (phi * o.wcurl3(a1, a2, o.c))).sum(axis=(1, 2)) - (o.c.WJ * (c_dot(o.curl12(phi, o.c), (a1, a2)))).sum(axis=(1, 2))
    assert np.abs(lhs / (o.c.WJ * np.abs(phi * a1 / o.c.J)).sum(axis=(1, 2))).max() < 1e-10


def c_dot(a, b):
    return a[0] * b[0] + a[1] * b[1]


def test_split_divergence_unit_tracer(case):
    """χ ≡ 1 consistency (test/prognostic_equations/tracer_mass_consistency_tests.jl:52-84):
    split_divₕ(ρu, 1) == wdivₕ(ρu) and every vertical_transport flavour with χ = 1 equals ∂ₜρ."""
    g, P, N, o, Yc, Yf, rng = case
    rho, u1, u2 = Yc[:, 0], Yc[:, 1], Yc[:, 2]
    c1, c2 = o.ct12(u1, u2, o.c)
    one = np.ones_like(rho)
    a = o.split_div(rho * c1, rho * c2, one, o.c)
    b = o.wdiv(rho * c1, rho * c2, o.c)
    assert np.abs(a - b).max() <= 100 * np.finfo(float).eps * np.abs(b).max()
    pc = o.set_implicit_precomputed_quantities(Yc, Yf)
    mass = -o.advdiv_f2c(o.interp_c2f(rho * o.c.J) / o.f.J * pc["fu3"])
    for up in ("none", "first_order", "third_order", "vanleer_limiter"):
        t = o.vertical_transport(rho, pc["fu3"], one, N.dt, up)
        assert np.abs(t - mass).max() <= 100 * np.finfo(float).eps * np.abs(mass).max()


def test_impenetrability_and_cache(case):
    """|ᶠu³| = 0 at both boundaries after cache_imp! (test/prognostic_equations/advection_tests.jl:46-58)."""
    g, P, N, o, Yc, Yf, rng = case
    Yf2 = Yf.copy()
    Yf2[..., 0] = 1.0
    Yf2[..., -1] = -2.0
    pc = o.set_implicit_precomputed_quantities(Yc, Yf2)
    assert np.all(pc["fu3"][..., 0] == 0) and np.all(pc["fu3"][..., -1] == 0)
    assert np.all(Yf2[..., 0] == 0) and np.all(Yf2[..., -1] == 0)
    assert np.all(pc["T"] >= P.T_min_sgs) and np.allclose(pc["p"], Yc[:, 0] * P.R_d * pc["T"])


def test_corrected_plus_central_is_upwind(case):
    """test/prognostic_equations/correct_implicit_advection_tests.jl:40-67."""
    g, P, N, o, Yc, Yf, rng = case
    pc = o.set_implicit_precomputed_quantities(Yc, Yf)
    corr, _ = o.correct_implicit_advection_tendency(Yc, Yf, pc)
    cen = o.vertical_transport(Yc[:, 0], pc["fu3"], pc["h_tot"], N.dt, "none")
    up = o.vertical_transport(Yc[:, 0], pc["fu3"], pc["h_tot"], N.dt, "vanleer_limiter")
    assert np.allclose(corr[:, 3] + cen, up, rtol=0, atol=1e-12 * np.abs(up).max())
    assert np.all(corr[:, :3] == 0)
    # van Leer face values are bounded by the neighbouring cell values (monotone)
    h = pc["h_tot"]
    fv = o.lin_vanleer(pc["fu3"], h, N.dt)[..., 1:-1] / np.where(pc["fu3"][..., 1:-1] == 0, 1, pc["fu3"][..., 1:-1])
    lo, hi = np.minimum(h[..., :-1], h[..., 1:]), np.maximum(h[..., :-1], h[..., 1:])
    ok = (pc["fu3"][..., 1:-1] == 0) | ((fv >= lo - 1e-9 * np.abs(hi)) & (fv <= hi + 1e-9 * np.abs(hi)))
    assert ok.all()


def test_sponge_profiles(case):
    """test/parameterized_tendencies/sponge.jl:44-80: β = 0 below zd, → coefficient at z_max."""
    g, P, N, o, Yc, Yf, rng = case
    z = np.array([0.0, P.zd_rayleigh, 0.5 * (P.zd_rayleigh + g.z_max), g.z_max])
    b = o.beta_rayleigh(z, 1.0)
    assert b[0] == 0 and b[1] == 0 and np.isclose(b[2], 0.5) and np.isclose(b[3], 1.0)
    bv = o.beta_viscous(z)
    assert bv[0] == 0 and np.isclose(bv[3], P.kappa_2_sponge)


def test_mass_and_vertical_telescoping(case):
    """Σ WJ ρₜ = 0 for the explicit horizontal mass flux (per element) and Σ_k J ρₜ = 0 per column
    for the implicit vertical flux (conservation to round-off, .buildkite/ci_driver.jl:176-195)."""
    g, P, N, o, Yc, Yf, rng = case
    pc = o.set_implicit_precomputed_quantities(Yc, Yf)
    tc, tf = o.implicit_tendency(Yc, Yf, pc)
    col = (o.c.J * tc[:, 0]).sum(axis=-1)
    assert np.abs(col).max() < 1e-12 * (o.c.J * np.abs(tc[:, 0])).sum(axis=-1).max()
    col = (o.c.J * tc[:, 3]).sum(axis=-1)
    assert np.abs(col).max() < 1e-12 * (o.c.J * np.abs(tc[:, 3])).sum(axis=-1).max()
    ec, ef = o.remaining_tendency(Yc, Yf, pc)
    tot = (o.c.WJ * ec[:, 0]).sum()
    assert abs(tot) < 1e-11 * (o.c.WJ * np.abs(ec[:, 0])).sum()


def test_dss_properties(case):
    g, P, N, o, Yc, Yf, rng = case
    a = rng.standard_normal(Yc[:, 0].shape)
    integ0 = (o.c.WJ * a).sum()
    b = a.copy()
    o.weighted_dss([("scalar", [b])])
    assert abs((o.c.WJ * b).sum() - integ0) < 1e-10 * (o.c.WJ * np.abs(a)).sum()  # integral preserved
    c = b.copy()
    o.weighted_dss([("scalar", [c])])
    assert np.abs(c - b).max() < 1e-13  # idempotent
    off, mem = o.dss_offs, o.dss_mem
    for n in range(0, len(off) - 1, 7):
        vals = np.array([b[e, j, i] for e, i, j in mem[off[n]:off[n + 1]]])
        assert np.abs(vals - vals[0]).max() < 1e-13  # continuous
    # a constant Cartesian vector field, expressed in covariant components, is continuous: the
    # vector DSS must leave it unchanged everywhere, including the elements touching the poles
    lat, lon = np.radians(g.lat), np.radians(g.lon)
    east = np.stack([-np.sin(lon), np.cos(lon), 0 * lon], -1)
    north = np.stack([-np.sin(lat) * np.cos(lon), -np.sin(lat) * np.sin(lon), np.cos(lat)], -1)
    vec = np.array([0.3, -1.1, 0.7])
    uu, vv = east @ vec, north @ vec
    u1 = (g.dxdxi[..., 0, 0] * uu + g.dxdxi[..., 1, 0] * vv)[..., None] * np.ones(g.nv)
    u2 = (g.dxdxi[..., 0, 1] * uu + g.dxdxi[..., 1, 1] * vv)[..., None] * np.ones(g.nv)
    v1, v2 = u1.copy(), u2.copy()
    o.weighted_dss([("c12", [v1, v2])])
    assert np.abs(v1 - u1).max() < 1e-9 * np.abs(u1).max() and np.abs(v2 - u2).max() < 1e-9 * np.abs(u2).max()


def apply_jacobian(o, Jm, dc, df):
    """J·ΔY with the block structure of manual_sparse_jacobian.jl (scalars and uₕ diagonal = −I)."""
    outc, outf = -dc.copy(), np.zeros_like(df)
    x = df[:, 0]
    for idx, key in ((0, "rho_u3"), (3, "rhoe_u3")):
        lo, hi = Jm[key]
        outc[:, idx] += lo * x[..., :-1] + hi * x[..., 1:]
    pad_lo = lambda a: np.concatenate([0 * a[..., :1], a], -1)
    pad_hi = lambda a: np.concatenate([a, 0 * a[..., :1]], -1)
    l, d, u = Jm["u3_u3"]
    r = d * x
    r[..., 1:] += l[..., 1:] * x[..., :-1]
    r[..., :-1] += u[..., :-1] * x[..., 1:]
    for key, a in (("u3_rho", dc[:, 0]), ("u3_rhoe", dc[:, 3])):
        lo, hi = Jm[key]
        r += lo * pad_lo(a) + hi * pad_hi(a)
    for k in range(2):
        lo, hi = Jm["u3_uh"][k]
        r += lo * pad_lo(dc[:, 1 + k]) + hi * pad_hi(dc[:, 1 + k])
    outf[:, 0] = r
    return outc, outf


def test_ldiv_inverts_the_block_matrix(case):
    g, P, N, o, Yc, Yf, rng = case
    pc = o.set_implicit_precomputed_quantities(Yc, Yf)
    Jm = o.update_jacobian(Yc, Yf, pc, 0.43 * N.dt)
    Rc, Rf = rng.standard_normal(Yc.shape), rng.standard_normal(Yf.shape)
    dc, df = o.ldiv(Jm, Rc, Rf)
    bc, bf = apply_jacobian(o, Jm, dc, df)
    assert np.abs(bc - Rc).max() < 1e-9 * max(1, np.abs(dc).max()) and np.abs(bf - Rf).max() < 1e-7 * max(1, np.abs(df).max())


def test_analytic_jacobian_matches_finite_differences():
    """The (u₃, ·) rows linearise the Exner-form PGF exactly in the continuum
    (manual_sparse_jacobian.jl:786-813 comment): on a smooth state and a fine uniform column the
    analytic blocks agree with a centred finite difference of dtγ·T_imp(Y) − Y to O(Δz²); the
    (ρ, u₃) row is linear in u₃ and agrees to round-off."""
    P = prm.DycoreParams()
    g = G.make_sphere_grid(FT=np.float64, h_elem=2, z_elem=48, z_max=30000.0, z_stretch=False, radius=P.planet_radius)
    N = prm.DycoreNumerics(dt=300.0)
    o = Oracle(g, P, N, np.float64)
    Yc, Yf = setups.dry_baroclinic_wave(g, P)
    Yf[:, 0, ..., 1:-1] = 0.05 * g.dz_f[1:-1] * np.sin(np.pi * g.z_f[1:-1] / g.z_max)  # smooth w ≠ 0
    dtg = 0.43 * N.dt
    pc = o.set_implicit_precomputed_quantities(Yc, Yf)
    Jm = o.update_jacobian(Yc, Yf, pc, dtg)

    def resid(yc, yf):
        yf = yf.copy()
        p = o.set_implicit_precomputed_quantities(yc, yf)
        tc, tf = o.implicit_tendency(yc, yf, p)
        return dtg * tc - yc, dtg * tf - yf

    k = 24
    for comp in (0, 1, 2, 3, "u3"):
        dc, df = np.zeros_like(Yc), np.zeros_like(Yf)
        if comp == "u3":
            df[:, 0, :, :, k] = np.abs(Yf).max()
        else:
            dc[:, comp, :, :, k] = np.abs(Yc[:, comp]).max()
        eps = 1e-6
        rp = resid(Yc + eps * dc, Yf + eps * df)
        rm = resid(Yc - eps * dc, Yf - eps * df)
        fd_f = (rp[1] - rm[1]) / (2 * eps)
        jc, jf = apply_jacobian(o, Jm, dc, df)
        sl = (slice(None), 0, slice(None), slice(None), slice(k - 2, k + 4))
        err = np.abs(fd_f[sl] - jf[sl]).max() / np.abs(jf[sl]).max()
        assert err < 2e-2, (comp, err)
        if comp == "u3":
            fd_c = (rp[0] - rm[0]) / (2 * eps)
            err = np.abs(fd_c[:, 0] - jc[:, 0]).max() / np.abs(jc[:, 0]).max()
            assert err < 1e-6, err


def test_steps_stay_finite_and_hydrostatic_column_is_steady():
    P = prm.DycoreParams()
    g = G.make_sphere_grid(FT=np.float64, h_elem=2, z_elem=20, z_max=30000.0, dz_bottom=500.0, radius=P.planet_radius)
    N = prm.DycoreNumerics(dt=400.0)
    o = Oracle(g, P, N, np.float64)
    Yc, Yf = setups.dry_baroclinic_wave(g, P, perturb=False)
    m0 = (o.c.WJ * Yc[:, 0]).sum()
    e0 = (o.c.WJ * Yc[:, 3]).sum()
    for _ in range(5):
        Yc, Yf = o.step(Yc, Yf)
    assert np.isfinite(Yc).all() and np.isfinite(Yf).all()
    # conservation (ci_driver.jl:176-195): mass to round-off; energy up to the T-floor/sponge-free residual
    assert abs((o.c.WJ * Yc[:, 0]).sum() - m0) / m0 < 1e-13
    assert abs((o.c.WJ * Yc[:, 3]).sum() - e0) / abs(e0) < 1e-6
    # the balanced zonal flow stays close to its initial state (steady-state check, ci_driver.jl:123-173)
    assert np.abs(Yf[:, 0] / g.dz_f).max() < 0.5


def test_threaded_step_is_bitwise_identical():
    """The multi-threaded CPU arm (element chunks on a thread pool, used by bench.py --impl reference) gives
    bitwise the same step as the serial oracle."""
    from concurrent.futures import ThreadPoolExecutor

    P = prm.DycoreParams(zd_rayleigh=20000.0, zd_viscous=20000.0)
    g = G.make_sphere_grid(FT=np.float32, h_elem=3, z_elem=12, z_max=30000.0, dz_bottom=300.0, radius=P.planet_radius)
    N = prm.DycoreNumerics(dt=300.0, rayleigh_sponge=True, viscous_sponge=True)
    o = Oracle(g, P, N, np.float32)
    Yc, Yf = setups.dry_baroclinic_wave(g, P)
    a = o.step(Yc.copy(), Yf.copy())
    with ThreadPoolExecutor(4) as pool:
        b = o.step(Yc.copy(), Yf.copy(), pool=pool, nchunks=5)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_golden_regression():
    """Oracle output pinned by a committed fixture (tests/golden/make_golden.py) so that later edits
    to the oracle cannot silently change what the CUDA path is compared against."""
    path = os.path.join(os.path.dirname(__file__), "golden", "oracle_step_he2_ze8_f64.npz")
    d = np.load(path)
    from tests.golden.make_golden import run_case

    Yc, Yf = run_case()
    for k in range(4):
        assert np.linalg.norm(Yc[:, k] - d["Yc"][:, k]) <= 1e-12 * np.linalg.norm(d["Yc"][:, k])
    assert np.linalg.norm(Yf - d["Yf"]) <= 1e-11 * np.linalg.norm(d["Yf"])


def test_golden_regression_vertical_diffusion_and_limiters():
    """Second committed fixture: implicit VerticalDiffusion (2 solver iterations) + one tracer + the vertical mass-borrowing limiter."""
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_step_vdiff_he2_ze8_f64.npz"))
    from tests.golden.make_golden import run_case_vdiff

    Yc, Yf = run_case_vdiff()
    for k in range(5):
        assert np.linalg.norm(Yc[:, k] - d["Yc"][:, k]) <= 1e-12 * np.linalg.norm(d["Yc"][:, k])
    assert np.linalg.norm(Yf - d["Yf"]) <= 1e-11 * np.linalg.norm(d["Yf"])


def _with_tracers(case, nq=2):
    g, P, N, o, Yc, Yf, rng = case
    zz = np.broadcast_to(g.z_c, Yc[:, 0].shape)
    chis = [np.ones_like(zz), 0.5 * (1 + np.sin(np.radians(g.lat[..., None])) * np.cos(np.radians(g.lon[..., None]))) * np.exp(-zz / 8000.0)]
    extra = [(Yc[:, 0] * chi)[:, None] for chi in chis[:nq]]
    Yq = np.concatenate([Yc] + extra, axis=1)
    o.dss_state(Yq, Yf)
    return Yq


def test_tracer_tendencies_unit_tracer_and_conservation(case):
    """Tracer-carrying state (ρχ appended to Y.c): the reference's χ ≡ 1 consistency test on the FULL tendencies
    (test/prognostic_equations/tracer_mass_consistency_tests.jl:52-84): Yₜ_lim.ρχ == horizontal ∂ₜρ and the explicit vertical
    transport of ρχ == the implicit ∂ₜρ; the flux-form tracer tendencies conserve global tracer mass; the dry components do not
    see the tracers; T_imp leaves passive tracers alone and ldiv! applies the −I fallback block to them."""
    g, P, N, o, Yc, Yf, rng = case
    Yq = _with_tracers(case)
    pc = o.set_implicit_precomputed_quantities(Yq, Yf)
    tc, tf, lc = o.remaining_tendency(Yq, Yf, pc, with_lim=True)
    tc0, tf0 = o.remaining_tendency(Yq[:, :4].copy(), Yf, pc)
    assert np.array_equal(tc[:, :4], tc0) and np.array_equal(tf, tf0)
    assert np.all(lc[:, :4] == 0)
    ic, _ = o.implicit_tendency(Yq, Yf, pc)
    eps = np.finfo(float).eps
    # viscous sponge and hyperdiffusion of χ ≡ 1 vanish identically, so the limited part is the horizontal mass tendency
    h_mass = tc[:, 0]
    assert np.abs(lc[:, 4] - h_mass).max() <= 1e3 * eps * np.abs(h_mass).max()
    assert np.abs(tc[:, 4] - ic[:, 0]).max() <= 1e3 * eps * np.abs(ic[:, 0]).max()
    assert np.all(ic[:, 4:] == 0)
    # global conservation of the second tracer: Σ WJ·(ρχ)ₜ = 0 for the horizontal (DSS-continuous state) and, per column,
    # Σ_k J·(ρχ)ₜ = 0 for the vertical flux divergence
    hor = lc[:, 5]
    assert abs((o.c.WJ * hor).sum()) < 1e-10 * (o.c.WJ * np.abs(hor)).sum()
    N2 = prm.DycoreNumerics(dt=N.dt, rayleigh_sponge=False, viscous_sponge=False, hyperdiff=False)
    o2 = Oracle(g, P, N2, np.float64)
    t2, _, l2 = o2.remaining_tendency(Yq, Yf, pc, with_lim=True)
    col = (o2.c.J * t2[:, 5]).sum(axis=-1)
    assert np.abs(col).max() < 1e-12 * (o2.c.J * np.abs(t2[:, 5])).sum(axis=-1).max()
    # ldiv!: passive tracers only have the −I block
    J = o.update_jacobian(Yq, Yf, pc, 0.4358665215084590 * N.dt)
    Rc, Rf = rng.standard_normal(Yq.shape), rng.standard_normal(Yf.shape)
    dc, df = o.ldiv(J, Rc, Rf)
    assert np.array_equal(dc[:, 4:], -Rc[:, 4:])


def test_step_with_tracers_leaves_dry_components_alone(case):
    """A full ARS343 step with passive tracers: the dry components do not depend on their presence, tracers stay finite and
    positive.  (χ ≡ 1 is NOT preserved by a step: ρ is advected vertically inside the implicit solve, ρχ explicitly at the stage
    state — the consistency the reference tests is the one of the tendencies, checked above.)"""
    g, P, N, o, Yc, Yf, rng = case
    Yq = _with_tracers(case)
    a, af = o.step(Yq.copy(), Yf.copy())
    b, bf = o.step(Yq[:, :4].copy(), Yf.copy())
    assert np.abs(a[:, :4] - b).max() <= 1e-12 * np.abs(b).max() and np.abs(af - bf).max() <= 1e-10 * max(np.abs(bf).max(), 1e-30)
    assert np.isfinite(a).all() and (a[:, 4] > 0).all()


def test_quasimonotone_limiter_properties(case):
    """lim! with the SEM quasi-monotone limiter (limited_tendencies.jl:64-122; ClimaCore Limiters.QuasiMonotoneLimiter): per
    (element, level) slab the tracer mass Σ WJ ρχ is conserved to round-off, χ ends inside the neighbour bounds (relaxed to the slab
    mean), a state inside its own bounds is untouched, and the limiter is idempotent."""
    g, P, N, o, Yc, Yf, rng = case
    N2 = prm.DycoreNumerics(dt=N.dt, apply_sem_quasimonotone_limiter=True)
    ol = Oracle(g, P, N2, np.float64)
    zz = np.broadcast_to(g.z_c, Yc[:, 0].shape)
    chi = (np.abs(g.lat[..., None]) < 30.0) * (zz < 12000.0) * 1.0
    ref = np.concatenate([Yc, (Yc[:, 0] * chi)[:, None]], axis=1)
    Y = ref.copy()
    Y[:, 4] += 0.3 * ref[:, 0] * rng.standard_normal(chi.shape)
    m0 = (ol.c.WJ * Y[:, 4]).sum(axis=(1, 2))
    rho_m = (ol.c.WJ * Y[:, 0]).sum(axis=(1, 2))
    qmin, qmax = ol.limiter_bounds(ref[:, 4], ref[:, 0])
    assert (qmin <= (ref[:, 4] / ref[:, 0]).min(axis=(1, 2))).all()  # neighbour bounds contain the element's own
    ol.limiters_func(Y, ref)
    m1 = (ol.c.WJ * Y[:, 4]).sum(axis=(1, 2))
    assert np.abs(m1 - m0).max() <= 1e-13 * np.abs(m0).max()
    q = Y[:, 4] / Y[:, 0]
    lo, hi = np.minimum(qmin, m0 / rho_m), np.maximum(qmax, m0 / rho_m)
    assert (q >= lo[:, None, None, :] - 1e-14).all() and (q <= hi[:, None, None, :] + 1e-14).all()
    Y2 = Y.copy()
    ol.limiters_func(Y2, ref)
    assert np.abs(Y2 - Y).max() <= 1e-15 * np.abs(Y).max()
    Y3 = ref.copy()
    ol.limiters_func(Y3, ref)
    assert np.array_equal(Y3, ref)
    # disabled limiter: the reference's no-op
    Y4 = Y.copy()
    o.limiters_func(Y4, ref)
    assert np.array_equal(Y4, Y)


def test_step_with_limiter_reduces_overshoots(case):
    """ARS343 step with T_lim kept apart and lim! between the limited and the unlimited increments (CTS update_stage!): the
    overshoot of a step-function tracer is smaller than without the limiter and the global tracer mass is unchanged."""
    g, P, N, o, Yc, Yf, rng = case
    zz = np.broadcast_to(g.z_c, Yc[:, 0].shape)
    chi = (np.abs(g.lat[..., None]) < 30.0) * (zz < 12000.0) * 1.0
    Yq = np.concatenate([Yc, (Yc[:, 0] * chi)[:, None]], axis=1)
    o.dss_state(Yq, Yf)
    N2 = prm.DycoreNumerics(dt=N.dt, rayleigh_sponge=True, viscous_sponge=True, apply_sem_quasimonotone_limiter=True)
    ol = Oracle(g, P, N2, np.float64)
    tr = []
    a, af = ol.step(Yq.copy(), Yf.copy(), trace=tr)
    b, bf = o.step(Yq.copy(), Yf.copy())
    assert tr.count("lim") == 4  # stages 2-4 and the final update
    assert np.abs(a[:, :4] - b[:, :4]).max() <= 1e-12 * np.abs(b[:, :4]).max()  # the dry components do not see the limiter
    over = lambda y: max((y[:, 4] / y[:, 0]).max() - 1.0, -(y[:, 4] / y[:, 0]).min())
    assert over(a) < 0.5 * over(b)
    ma, mb = (ol.c.WJ * a[:, 4]).sum(), (o.c.WJ * b[:, 4]).sum()
    assert abs(ma - mb) <= 1e-9 * abs(mb)


def _balanced_state_residual(he, ze, deep=False):
    P = prm.DycoreParams()
    g = G.make_sphere_grid(FT=np.float64, h_elem=he, z_elem=ze, z_max=30000.0, dz_bottom=30000.0 / ze, radius=P.planet_radius,
                           deep_atmosphere=deep)
    o = Oracle(g, P, prm.DycoreNumerics(dt=100.0, hyperdiff=False), np.float64)
    Yc, Yf = setups.dry_baroclinic_wave(g, P, perturb=False)
    pc = o.set_implicit_precomputed_quantities(Yc, Yf)
    tc, tf = o.remaining_tendency(Yc, Yf, pc)
    ic, if_ = o.implicit_tendency(Yc, Yf, pc)
    tot_c, tot_f = tc + ic, tf + if_
    o.dss_state(tot_c, tot_f)  # the tendency as it acts on the continuous state
    c1, c2 = o.ct12(tot_c[:, 1], tot_c[:, 2], o.c)
    l2 = lambda a, w: np.sqrt((w * a**2).sum() / w.sum())
    uh = l2(np.sqrt(np.abs(tot_c[:, 1] * c1 + tot_c[:, 2] * c2)), o.c.WJ)       # |∂ₜuₕ| in m s⁻²
    rho = l2(tot_c[:, 0], o.c.WJ) / l2(Yc[:, 0], o.c.WJ)
    w = (tot_f[:, 0] / g.dz_f)[..., 1:-1]                                       # ∂ₜw in m s⁻² at interior faces
    return uh, rho, l2(w, o.f.WJ[..., 1:-1]), w, g


def test_convergence_to_the_analytic_steady_state():
    """Independent of any recalled implementation detail: the unperturbed baroclinic-wave state (Ullrich et al. 2014;
    src/setups/DryBaroclinicWave.jl:88-155) is an exact steady solution of the governing equations, so the total discrete tendency
    T_exp + T_imp evaluated on it is pure truncation error.  It must vanish at the design orders: ≥ 3rd order in the element size for
    the Nq = 4 spectral-element operators (horizontal momentum, mass), 2nd order in Δz for the staggered finite differences
    (vertical momentum: discrete hydrostatic balance).  This pins signs, metric terms, the Coriolis and pressure-gradient
    formulations and the implicit/explicit split of the oracle against the PDE itself."""
    uh2, r2, _, _, _ = _balanced_state_residual(2, 30)
    uh4, r4, w30, _, _ = _balanced_state_residual(4, 30)
    uh8, r8, _, _, _ = _balanced_state_residual(8, 30)
    assert uh4 < uh2 / 6 and uh8 < uh4 / 6, (uh2, uh4, uh8)      # measured 7.8 per halving
    assert r4 < r2 / 5 and r8 < r4 / 5, (r2, r4, r8)
    assert uh8 < 3e-6 and r8 < 1e-8                               # m s⁻² and s⁻¹: six orders below f·u and u/a
    _, _, w15, _, _ = _balanced_state_residual(4, 15)
    _, _, w60, _, _ = _balanced_state_residual(4, 60)
    assert w30 < w15 / 3.5 and w60 < w30 / 3.5, (w15, w30, w60)   # measured 3.8–3.9 per halving
    # Deep atmosphere: the reference's deep initial state is balanced for g(r) = g·(a/r)² while the model's Φ = g·z (cache.jl:179-183);
    # the vertical residual must therefore be exactly that gravity mismatch (a consistency check of the deep metric terms).
    _, _, _, w, g = _balanced_state_residual(4, 60, deep=True)
    a = g.radius
    expect = -prm.DycoreParams().grav * (1.0 - (a / (a + g.z_f[1:-1])) ** 2)
    assert np.abs(w - expect).max() < 4e-3                         # Coriolis/metric and truncation terms of the deep state
    assert np.abs(w[..., -1] - expect[-1]).max() < 0.05 * abs(expect[-1])



@pytest.mark.parametrize("name,he,ze,zmax,dzb,dt,sponge", [("he4ze10", 4, 10, 30000.0, 500.0, 400.0, False),
                                                          ("he3ze63", 3, 63, 60000.0, 30.0, 120.0, True)])
def test_float32_floor_of_the_reference_formulation(name, he, ze, zmax, dzb, dt, sponge):
    """The number behind the Float32 u₃ tolerance of the GPU tests.  The oracle restates the reference's formulas literally; run in
    Float32 (what ClimaAtmos computes with FLOAT_TYPE Float32) against itself in Float64 on the same Float32 initial state, one step:
    ρ, uₕ, ρe_tot agree to ≤ 5e-6, but u₃ — produced by the cancellation ᶠgradᵥΦ − ᶠgradᵥΦ_r + cp_d ᶠinterp(θ′) ᶠgradᵥΠ
    (implicit_tendency.jl:292-293) of level values up to 6e5 J/kg — only to 1.4e-5 … 2.7e-5.  A Float32 implementation that follows the
    reference formula cannot be closer than that to the Float64 result, and two such implementations differ from each other by the
    same amount; the CUDA Float32 kernels use the difference form (common.cuh pgf_diff) and land BELOW this floor."""
    P = prm.DycoreParams(zd_rayleigh=0.66 * zmax, zd_viscous=0.66 * zmax)
    g = G.make_sphere_grid(FT=np.float32, h_elem=he, z_elem=ze, z_max=zmax, dz_bottom=dzb, radius=P.planet_radius)
    N = prm.DycoreNumerics(dt=dt, rayleigh_sponge=sponge, viscous_sponge=sponge)
    Yc0, Yf0 = setups.dry_baroclinic_wave(g, P)
    Yc0, Yf0 = Yc0.astype(np.float32), Yf0.astype(np.float32)
    c32, f32 = Oracle(g, P, N, np.float32).step(Yc0.copy(), Yf0.copy())
    c64, f64 = Oracle(g, P, N, np.float64).step(Yc0.astype(np.float64), Yf0.astype(np.float64))
    rel = lambda a, b: np.linalg.norm((a.astype(np.float64) - b).ravel()) / np.linalg.norm(b.ravel())
    for k in range(4):
        assert rel(c32[:, k], c64[:, k]) < 5e-6, k
    gap = rel(f32, f64)
    assert 1.0e-5 < gap < 4e-5, gap  # the floor: above the 1e-5 bar, which is why the bar needs the difference form



def test_third_order_upwinding_reconstruction():
    """ᶠupwind3 (abbreviations.jl:229-240): on a uniform column the interior stencil and the two one-sided closures are finite-volume
    reconstructions — from the cell averages of a quadratic profile they return its face values exactly (third-order accuracy);
    the interior value is the upwind-biased one; and `corrected + central == upwind` holds for :third_order like for the other
    schemes (correct_implicit_advection_tests.jl:40-67)."""
    P = prm.DycoreParams()
    g = G.make_sphere_grid(FT=np.float64, h_elem=2, z_elem=9, z_max=9000.0, dz_bottom=1000.0, radius=P.planet_radius)  # dz_bottom = z_max/z_elem: uniform
    assert np.allclose(np.diff(g.z_f), 1000.0)
    N = prm.DycoreNumerics(dt=100.0, energy_upwinding="third_order")
    o = Oracle(g, P, N, np.float64)
    zc, zf = g.z_c / 1000.0, g.z_f / 1000.0  # h = 1
    q2 = 0.13
    quad = lambda z: 2.0 - 0.7 * z + q2 * z * z
    cell_avg = quad(zc) + 2 * q2 / 24.0  # average of a quadratic over a cell of width 1 = midpoint value + h² q''/24
    a = np.broadcast_to(cell_avg, (g.nelems, 4, 4, g.nv)).copy()
    rng = np.random.default_rng(3)
    v = rng.standard_normal((g.nelems, 4, 4, g.nv + 1))
    rec = o.upwind3(v, a)[..., 1:-1] / v[..., 1:-1]
    assert np.abs(rec - quad(zf)[1:-1]).max() < 1e-12  # every interior face, both closures included, either sign of v
    # interior: upwind-biased
    b = rng.standard_normal(a.shape)
    f3 = o.upwind3(v, b)[..., 2:-2]
    amm, am, ap, app = b[..., :-3], b[..., 1:-2], b[..., 2:-1], b[..., 3:]
    vf = v[..., 2:-2]
    up = np.where(vf >= 0, (-amm + 5 * am + 2 * ap) / 6, (2 * am + 5 * ap - app) / 6)
    assert np.abs(f3 - vf * up).max() < 1e-13
    # corrected + central == upwind (T_post_imp! contract)
    Yc, Yf = setups.dry_baroclinic_wave(g, P)
    Yf = 0.3 * g.dz_f * rng.standard_normal(Yf.shape)
    pc = o.set_implicit_precomputed_quantities(Yc, Yf)
    corr, _ = o.correct_implicit_advection_tendency(Yc, Yf, pc)
    cen = o.vertical_transport(Yc[:, 0], pc["fu3"], pc["h_tot"], N.dt, "none")
    upw = o.vertical_transport(Yc[:, 0], pc["fu3"], pc["h_tot"], N.dt, "third_order")
    assert np.abs(corr[:, 3] + cen - upw).max() <= 1e-10 * np.abs(upw).max()
