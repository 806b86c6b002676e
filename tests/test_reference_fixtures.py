"""Pinning the oracle to the REAL reference (VERDICT r1 "pin the oracle").

`tools/dump_reference_vectors.jl` (run with the pinned ClimaAtmos v0.42.7 / ClimaCore 0.15.1 / ClimaTimeSteppers 0.10.6 environment)
dumps, for three small configurations, the space-filling-curve element order, the Topology2D tables, the horizontal LocalGeometry,
the initial state, every hook's output and Y after one and two `CTS.step!`; `tools/ref_vectors_to_npz.py` packs them into
`tests/golden/ref_<case>.npz`.  Neither this container nor the GPU box has julia (profiles/r2_gpu_box_probe.txt), so the fixtures
cannot be produced by the builder: until a maintainer commits them every test here is a STRICT xfail — it fails on the missing
fixture, and the day the fixture exists it must pass (an unexpected pass without the comparison running is impossible: the
comparison is the test body).  Until then the oracle header and DESIGN.md §5 say "parity unpinned".

What is compared (tolerances: the north-star bars — 1e-12 Float64 / 1e-5 Float32 rel-L2 per field; index tables bit-exact):
  * grid: `topo.elemorder` against grid.py's space-filling curve, `interior_faces` / `local_vertices` against grid.py's Topology2D,
    lat/long/J of the horizontal LocalGeometry, GLL points / weights / D, vertical levels, ν₄;
  * the analytic initial state (src/setups/DryBaroclinicWave.jl) against climaatmos_jl_b200/setups.py;
  * cache_imp!, T_exp_T_lim!, T_imp!, Wfact + ldiv!, T_post_imp!, dss! on the reference's own inputs;
  * Y after one and two steps (pins the CTS stage order the fused stepper bakes in).
"""
import os

import numpy as np
import pytest

from climaatmos_jl_b200 import grid as G, params as prm, setups
from oracle.dycore_oracle import Oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {
    # name: (h_elem, z_elem, z_max, dz_bottom, dt, sponges, FT, zd)
    "he4ze10_f64": (4, 10, 30000.0, 500.0, 400.0, False, np.float64, None),
    "he6ze10_f32": (6, 10, 30000.0, 500.0, 400.0, False, np.float32, None),
    "he3ze63_f64": (3, 63, 60000.0, 30.0, 120.0, True, np.float64, 40000.0),
    "he4ze10_moist_f64": (4, 10, 30000.0, 500.0, 400.0, False, np.float64, None),  # EquilibriumMicrophysics0M, MoistBaroclinicWave
}


def fixture(case):
    return os.path.join(GOLDEN, f"ref_{case}.npz")


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    n = np.linalg.norm(b.ravel())
    return np.linalg.norm((a - b).ravel()) / (n if n > 0 else 1.0)


def strict_xfail_without(case):
    return pytest.mark.xfail(not os.path.exists(fixture(case)), strict=True,
                             reason=f"tests/golden/ref_{case}.npz absent: produce it with tools/dump_reference_vectors.jl on a machine with the "
                                    "pinned Julia environment (no julia in the build container or on the GPU box) — parity unpinned until then")


def setup(case):
    he, ze, zmax, dzb, dt, sp, FT, zd = CASES[case]
    R = np.load(fixture(case))
    kw = dict(zd_rayleigh=zd, zd_viscous=zd) if zd else {}
    P = prm.DycoreParams(**kw)
    g = G.make_sphere_grid(FT=FT, h_elem=he, z_elem=ze, z_max=zmax, dz_bottom=dzb, radius=P.planet_radius)
    N = prm.DycoreNumerics(dt=dt, rayleigh_sponge=sp, viscous_sponge=sp, microphysics_model="0M" if "moist" in case else None)
    return R, P, g, N, FT, (1e-12 if FT == np.float64 else 1e-5)


def state(R, name):
    return np.ascontiguousarray(R[name + "_c"]), np.ascontiguousarray(R[name + "_f"])


@pytest.mark.parametrize("case", [pytest.param(c, marks=strict_xfail_without(c)) for c in CASES])
def test_grid_tables_and_geometry_match_climacore(case):
    R, P, g, N, FT, tol = setup(case)
    # space-filling-curve order: ClimaCore stores CartesianIndex (i, j, panel) per element, 1-based
    order = np.asarray(R["elemorder"]).reshape(-1, 3) - 1
    ours = np.asarray(g.topology.elemorder)  # (nelems, 3): (ex, ey, panel) of every element in SFC order (grid.py spacefillingcurve)
    assert np.array_equal(order, ours), "space-filling-curve element order differs from ClimaCore's"
    faces = np.asarray(R["interior_faces"]).reshape(-1, 5).copy()
    faces[:, :4] -= 1
    assert np.array_equal(faces, np.asarray(g.topology.interior_faces)), "interior_faces table differs (bit-exact contract)"
    lv = np.asarray(R["local_vertices"]).reshape(-1, 2) - 1
    assert np.array_equal(lv, np.asarray(g.topology.local_vertices))
    assert np.array_equal(np.asarray(R["local_vertex_offset"]).ravel() - 1, np.asarray(g.topology.local_vertex_offset))
    assert np.allclose(R["gll_weights"].ravel(), g.wq, rtol=1e-14) and np.allclose(R["gll_D"].reshape(4, 4).T, g.D, rtol=1e-13, atol=1e-14)
    assert np.allclose(R["z_c"].ravel(), g.z_c, rtol=1e-6 if FT == np.float32 else 1e-13)
    assert np.allclose(R["z_f"].ravel(), g.z_f, rtol=1e-6 if FT == np.float32 else 1e-13, atol=1e-9)
    assert rel(R["lat"][:, 0], g.lat) < (1e-6 if FT == np.float32 else 1e-13)
    o = Oracle(g, P, N, FT)
    assert rel(R["J_c"][:, 0], o.c.J) < max(tol, 1e-13) and rel(R["J_f"][:, 0], o.f.J) < max(tol, 1e-13)
    assert abs(float(R["nu4_vorticity"].ravel()[0]) / float(o.nu4_vort) - 1) < 1e-6


@pytest.mark.parametrize("case", [pytest.param(c, marks=strict_xfail_without(c)) for c in CASES])
def test_initial_state_and_hooks_match_the_reference(case):
    R, P, g, N, FT, tol = setup(case)
    o = Oracle(g, P, N, FT)
    Yc0, Yf0 = setups.dry_baroclinic_wave(g, P)
    rc, rf = state(R, "Y0")
    for k in range(4):
        assert rel(Yc0[:, k], rc[:, k]) < tol, ("initial state", k)
    # hooks on the REFERENCE's state (so that differences in the initial state do not leak into the hook comparison)
    Uc, Uf = rc.astype(FT).copy(), rf.astype(FT).copy()
    pc = o.set_implicit_precomputed_quantities(Uc, Uf)
    cc, cf = state(R, "cache_imp_Y")
    assert rel(Uf, cf) < tol
    for key, name in (("K", "ᶜK"), ("T", "ᶜT"), ("p", "ᶜp"), ("h_tot", "ᶜh_tot")):
        assert rel(pc[key], R["precomputed_" + name][:, 0]) < tol, name
    tc, tf = o.remaining_tendency(Uc, Uf, pc)
    gc, gf = state(R, "t_exp")
    ttol = tol if FT == np.float64 else 2e-4  # cancelling tendencies in Float32 (tests/test_gpu_parity.py)
    for k in range(4):
        assert rel(tc[:, k], gc[:, k]) < ttol, ("T_exp", k)
    assert rel(tf, gf) < ttol
    ic, if_ = o.implicit_tendency(Uc, Uf, pc)
    gc, gf = state(R, "t_imp")
    assert rel(ic[:, 0], gc[:, 0]) < ttol and rel(ic[:, 3], gc[:, 3]) < ttol and rel(if_, gf) < ttol
    dtg = FT(N.dt * 0.4358665215084590)
    Jm = o.update_jacobian(Uc, Uf, pc, dtg)
    Rc, Rf = state(R, "ldiv_R")
    dc, df = o.ldiv(Jm, Rc.astype(FT), Rf.astype(FT))
    gc, gf = state(R, "ldiv_dY")
    for k in range(4):
        assert rel(dc[:, k], gc[:, k]) < max(tol, 1e-11), ("ldiv", k)
    assert rel(df, gf) < max(tol, 1e-11)
    if "t_post_imp_c" in R:
        pc_, _ = o.correct_implicit_advection_tendency(Uc, Uf, pc)
        assert rel(pc_[:, 3], R["t_post_imp_c"][:, 3]) < ttol
    ic_, if2 = state(R, "dss_in")
    ic_, if2 = ic_.astype(FT).copy(), if2.astype(FT).copy()
    o.dss_state(ic_, if2)
    gc, gf = state(R, "dss_out")
    assert rel(ic_, gc) < max(tol, 1e-13) and rel(if2, gf) < max(tol, 1e-13)


@pytest.mark.parametrize("case", [pytest.param(c, marks=strict_xfail_without(c)) for c in CASES])
def test_one_and_two_steps_match_the_reference(case):
    R, P, g, N, FT, tol = setup(case)
    o = Oracle(g, P, N, FT)
    Yc, Yf = [a.astype(FT) for a in state(R, "Y0")]
    for name in ("Y1", "Y2"):
        Yc, Yf = o.step(Yc, Yf)
        rc, rf = state(R, name)
        for k in range(4):
            assert rel(Yc[:, k], rc[:, k]) < tol, (name, k)
        assert rel(Yf, rf) < (tol if FT == np.float64 else 4e-5), name  # Float32 u₃: the floor of the literal formulation
