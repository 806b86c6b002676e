"""GPU parity tests for vertical diffusion and the approximate arrowhead solve — kernels_vdiff.cuh
(k_vdiff_tend, k_vdiff_jac, k_ldiv_diff) through the C-ABI against the NumPy oracle.

Run on a B200 at the end of round 1 (profiles/r1_vdiff_gpu_pytest.log, profiles/r1_vdiff_gpu_quickcheck.log: Float64 ≤ 1e-13,
Float32 ≤ 1e-6 on the centre fields).  Tolerances as in
tests/test_gpu_parity.py (Float64 1e-11; Float32 1e-5 state, u₃ 5e-5 on this 10-level grid with diffusion / 5e-4 cancelling tendencies)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from climaatmos_jl_b200 import dycore, params as prm
from oracle.dycore_oracle import Oracle

torch = pytest.importorskip("torch")


def rel(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    n = np.linalg.norm(b.ravel())
    d = np.linalg.norm((a - b).ravel())
    return d / n if n > 0 else d


def make(FT, vd, implicit, ntr=1, iters=2, deep=True, rad=None, ze=10):
    P = prm.DycoreParams(D_0_diffusion=60.0, H_diffusion=5000.0, C_E=0.0044)
    tracers = [lambda lat, lon, z: 1e-2 * (1 + 0.5 * np.cos(z / 900.0) * np.cos(np.radians(lat))) + 0 * lon][:ntr]
    sim = dycore.AtmosSimulation(FT=FT, h_elem=3, z_elem=ze, z_max=30000.0, dz_bottom=500.0, dt=200.0, params=P, tracers=tracers,
                                 vert_diff=vd, implicit_diffusion=implicit, approximate_linear_solve_iters=iters,
                                 deep_atmosphere=deep, rad=rad)
    o = Oracle(sim.grid, P, sim.numerics, FT)
    Yc0, Yf0 = sim.Y.cpu()
    rng = np.random.default_rng(1234)
    Yc = (Yc0.astype(np.float64) * (1 + 1e-3 * rng.standard_normal(Yc0.shape))).astype(FT)
    Yf = (0.5 * sim.grid.dz_f * rng.standard_normal(Yf0.shape)).astype(FT)
    Yf[..., 0] = 0
    Yf[..., -1] = 0
    return sim, o, Yc, Yf, rng


@pytest.mark.parametrize("FT", [np.float64, np.float32])
@pytest.mark.parametrize("vd", ["DecayWithHeightDiffusion", "VerticalDiffusion"])
@pytest.mark.parametrize("deep", [True, False])
def test_implicit_diffusion_hooks_match_oracle(FT, vd, deep):
    """T_imp! (+ diffusion), Wfact (+ diffusion blocks) and ldiv! (ApproximateBlockArrowheadIterativeSolve, 0–3 iterations)."""
    sim, o, Yc, Yf, rng = make(FT, vd, True, deep=deep)
    Y = sim.to_device(Yc, Yf)
    oc, of = Yc.copy(), Yf.copy()
    pc = o.set_implicit_precomputed_quantities(oc, of)
    Yt = Y.zeros_like()
    sim.implicit_tendency(Yt, Y)
    tc, tf = o.implicit_tendency(oc, of, pc)
    gc, gf = Yt.cpu()
    t64 = FT == np.float64
    for k, lim in ((0, 1e-5), (1, 5e-4), (2, 5e-4), (3, 1e-4), (4, 1e-4)):
        assert rel(gc[:, k], tc[:, k]) <= (1e-11 if t64 else lim), f"t_imp comp {k}: {rel(gc[:, k], tc[:, k]):.3e}"
    assert rel(gf, tf) <= (1e-11 if t64 else 5e-4)
    dtg = sim.dt * 0.4358665215
    sim.update_jacobian(Y, dtg)
    Jm = o.update_jacobian(oc, of, pc, dtg)
    Rc = (rng.standard_normal(Yc.shape) * np.abs(Yc).mean(axis=(0, 2, 3, 4), keepdims=True) * 1e-3).astype(FT)
    Rf = rng.standard_normal(Yf.shape).astype(FT)
    R = sim.to_device(Rc, Rf)
    for iters in (2, 0, 1, 3):
        sim.close()
        sim, o, _, _, _ = make(FT, vd, True, iters=iters, deep=deep)
        sim.update_jacobian(Y, dtg)
        dY = R.zeros_like()
        sim.ldiv(dY, R)
        o.N.approximate_linear_solve_iters = iters
        dc, df = o.ldiv(Jm, Rc, Rf)
        gc, gf = dY.cpu()
        for k in range(5):
            e = rel(gc[:, k], dc[:, k])
            assert e <= (1e-10 if t64 else 5e-5), f"ldiv iters={iters} comp {k}: {e:.3e}"
        assert rel(gf, df) <= (1e-10 if t64 else 5e-5), f"ldiv iters={iters} u3"
    sim.close()


@pytest.mark.parametrize("FT", [np.float64, np.float32])
@pytest.mark.parametrize("vd", ["DecayWithHeightDiffusion", "VerticalDiffusion"])
def test_explicit_diffusion_joins_t_exp(FT, vd):
    """diff_mode Explicit: the diffusion tendency is part of T_exp_T_lim! (remaining_tendency.jl:185-195); T_imp!/ldiv! unchanged."""
    sim, o, Yc, Yf, rng = make(FT, vd, False)
    Y = sim.to_device(Yc, Yf)
    sim.dss(Y)
    oc, of = Yc.copy(), Yf.copy()
    o.dss_state(oc, of)
    pc = o.set_implicit_precomputed_quantities(oc, of)
    Yt, Yl = Y.zeros_like(), Y.zeros_like()
    sim.remaining_tendency(Yt, Yl, Y)
    tc, tf, lc = o.remaining_tendency(oc, of, pc, with_lim=True)
    gc, gf = Yt.cpu()
    gl, _ = Yl.cpu()
    t64 = FT == np.float64
    for k, lim in ((0, 1e-5), (1, 5e-4), (2, 5e-4), (3, 1e-4)):
        assert rel(gc[:, k], tc[:, k]) <= (1e-11 if t64 else lim), f"t_exp comp {k}: {rel(gc[:, k], tc[:, k]):.3e}"
    assert rel(gc[:, 4] + gl[:, 4], tc[:, 4] + lc[:, 4]) <= (1e-11 if t64 else 1e-4)
    assert rel(gf, tf) <= (1e-11 if t64 else 5e-4)
    sim.close()


def test_momentum_diffusion_disabled_for_held_suarez():
    """rad = held_suarez sets disable_momentum_vertical_diffusion (type_getters.jl:46): uₕ rows fall back to −I."""
    FT = np.float64
    sim, o, Yc, Yf, rng = make(FT, "DecayWithHeightDiffusion", True, ntr=0, rad="held_suarez")
    assert o.N.disable_momentum_vertical_diffusion
    Y = sim.to_device(Yc, Yf)
    pc = o.set_implicit_precomputed_quantities(Yc.copy(), Yf.copy())
    dtg = 80.0
    sim.update_jacobian(Y, dtg)
    Jm = o.update_jacobian(Yc, Yf, pc, dtg)
    Rc, Rf = rng.standard_normal(Yc.shape), rng.standard_normal(Yf.shape)
    R = sim.to_device(Rc, Rf)
    dY = R.zeros_like()
    sim.ldiv(dY, R)
    dc, df = o.ldiv(Jm, Rc, Rf)
    gc, gf = dY.cpu()
    assert np.array_equal(gc[:, 1:3], -Rc[:, 1:3])
    for k in (0, 3):
        assert rel(gc[:, k], dc[:, k]) <= 1e-10
    assert rel(gf, df) <= 1e-10
    sim.close()


@pytest.mark.parametrize("FT", [np.float64, np.float32])
@pytest.mark.parametrize("implicit", [False, True])
def test_step_with_vertical_diffusion_matches_oracle(FT, implicit):
    """One ARS343 step (fused entry point; with implicit diffusion the implicit stages run through the hook sequence)."""
    sim, o, _, _, _ = make(FT, "DecayWithHeightDiffusion", implicit)
    Yc0, Yf0 = sim.Y.cpu()
    sim.step(fused=True)
    torch.cuda.synchronize()
    gc, gf = sim.Y.cpu()
    o64 = Oracle(sim.grid, sim.params, sim.numerics, np.float64)
    oc, of = o64.step(Yc0.astype(np.float64), Yf0.astype(np.float64))
    t64 = FT == np.float64
    for k in range(5):
        e = rel(gc[:, k], oc[:, k])
        assert e <= (1e-11 if t64 else 1e-5), f"step comp {k}: {e:.3e}"
    assert rel(gf, of) <= (1e-11 if t64 else 5e-5)
    # literal hook-by-hook step agrees with the fused entry point
    sim2, _, _, _, _ = make(FT, "DecayWithHeightDiffusion", implicit)
    sim2.step(fused=False)
    torch.cuda.synchronize()
    hc, hf = sim2.Y.cpu()
    assert rel(hc, gc) <= (1e-12 if t64 else 1e-5)
    sim.close()
    sim2.close()


@pytest.mark.parametrize("FT", [np.float64, np.float32])
def test_vertical_mass_borrowing_limiter_matches_oracle(FT):
    """lim! with tracer_nonnegativity_method: vertical_water_borrowing (limited_tendencies.jl:95-121) through b200_lim, and a step."""
    P = prm.DycoreParams()
    tr = [lambda lat, lon, z: 4e-4 + 1e-3 * np.cos(z / 700.0) * np.cos(3 * np.radians(lat)) + 0 * lon]
    sim = dycore.AtmosSimulation(FT=FT, h_elem=3, z_elem=10, z_max=30000.0, dz_bottom=500.0, dt=200.0, params=P, tracers=tr,
                                 tracer_nonnegativity_method="vertical_water_borrowing")
    o = Oracle(sim.grid, P, sim.numerics, FT)
    Yc, Yf = sim.Y.cpu()
    assert (Yc[:, 4] < 0).any()
    ref = Yc.copy()
    o.limiters_func(ref, Yc)
    Y = sim.to_device(Yc, Yf)
    sim.limiters_func(Y, 0.0, Y)
    gc, _ = Y.cpu()
    assert (gc[:, 4] >= 0).all()
    assert np.array_equal(gc[:, :4], Yc[:, :4])
    assert rel(gc[:, 4], ref[:, 4]) <= (1e-14 if FT == np.float64 else 1e-6)
    Yc0, Yf0 = sim.Y.cpu()
    sim.step(fused=True)
    torch.cuda.synchronize()
    gc, gf = sim.Y.cpu()
    o64 = Oracle(sim.grid, P, sim.numerics, np.float64)
    oc, of = o64.step(Yc0.astype(np.float64), Yf0.astype(np.float64))
    for k in range(5):
        assert rel(gc[:, k], oc[:, k]) <= (1e-11 if FT == np.float64 else (1e-5 if k < 4 else 1e-4)), k
    sim.close()


@pytest.mark.parametrize("FT", [np.float64, np.float32])
@pytest.mark.parametrize("ze,zmax,dzb", [(10, 30000.0, 500.0), (63, 60000.0, 30.0), (2, 10000.0, 5000.0)])
def test_dry_hook_kernels_at_column_height_limits(FT, ze, zmax, dzb):
    """k_t_imp2, k_wfact2 + k_ldiv2, k_t_post_imp2: T_imp!, Wfact + ldiv!, T_post_imp! against the oracle for nv = 2, 10, 63, and the hook-by-hook
    step against the fused step."""
    P = prm.DycoreParams(zd_rayleigh=0.66 * zmax, zd_viscous=0.66 * zmax)
    sim = dycore.AtmosSimulation(FT=FT, h_elem=3, z_elem=ze, z_max=zmax, dz_bottom=dzb, dt=150.0, rayleigh_sponge=True, viscous_sponge=True, params=P)
    o = Oracle(sim.grid, P, sim.numerics, FT)
    Yc0, Yf0 = sim.Y.cpu()
    rng = np.random.default_rng(1234)
    Yc = (Yc0.astype(np.float64) * (1 + 1e-3 * rng.standard_normal(Yc0.shape))).astype(FT)
    Yf = (0.5 * sim.grid.dz_f * rng.standard_normal(Yf0.shape)).astype(FT)
    Yf[..., 0] = 0
    Yf[..., -1] = 0
    Y = sim.to_device(Yc, Yf)
    pc = o.set_implicit_precomputed_quantities(Yc.copy(), Yf.copy())
    t64 = FT == np.float64
    Yt = Y.zeros_like()
    sim.implicit_tendency(Yt, Y)
    tc, tf = o.implicit_tendency(Yc, Yf, pc)
    gc, gf = Yt.cpu()
    assert rel(gc[:, 0], tc[:, 0]) <= (1e-11 if t64 else 1e-5) and rel(gc[:, 3], tc[:, 3]) <= (1e-11 if t64 else 1e-4)
    assert rel(gf, tf) <= (1e-11 if t64 else 5e-4)
    dtg = sim.dt * 0.4358665215
    sim.update_jacobian(Y, dtg)
    Jm = o.update_jacobian(Yc, Yf, pc, dtg)
    Rc = (rng.standard_normal(Yc.shape) * np.abs(Yc).mean(axis=(0, 2, 3, 4), keepdims=True) * 1e-3).astype(FT)
    Rf = rng.standard_normal(Yf.shape).astype(FT)
    R = sim.to_device(Rc, Rf)
    dY = R.zeros_like()
    sim.ldiv(dY, R)
    dc, df = o.ldiv(Jm, Rc, Rf)
    gc, gf = dY.cpu()
    for k in range(4):
        assert rel(gc[:, k], dc[:, k]) <= (1e-10 if t64 else 5e-5), k
    assert rel(gf, df) <= (1e-10 if t64 else 5e-5)
    sim.correct_implicit_advection_tendency(Yt, Y)
    pcc, _ = o.correct_implicit_advection_tendency(Yc, Yf, pc)
    gc, gf = Yt.cpu()
    assert rel(gc[:, 3], pcc[:, 3]) <= (1e-10 if t64 else 5e-4)
    sim.Y = sim.to_device(Yc0, Yf0)
    sim.step(fused=False)
    hc, hf = sim.Y.cpu()
    sim.Y = sim.to_device(Yc0, Yf0)
    sim.step(fused=True)
    fc, ff = sim.Y.cpu()
    assert rel(hc, fc) <= (1e-12 if t64 else 1e-5)
    sim.close()


@pytest.mark.parametrize("FT", [np.float64, np.float32])
@pytest.mark.parametrize("vd", ["DecayWithHeightDiffusion", "VerticalDiffusion"])
def test_fused_implicit_diffusion_stage(FT, vd):
    """k_imp_stage_diff: b200_implicit_stage with implicit diffusion as one kernel against the oracle's hook sequence, and the fused,
    graph-replayed step against the oracle and against the hook-by-hook step."""
    sim, o, Yc, Yf, rng = make(FT, vd, True)
    U = sim.to_device(Yc, Yf)
    N = U.zeros_like()
    dtg = sim.dt * 0.4358665215
    sim.implicit_stage(N, U, dtg)
    torch.cuda.synchronize()
    oc, of = Yc.astype(np.float64), Yf.astype(np.float64)
    o64 = Oracle(sim.grid, sim.params, sim.numerics, np.float64)
    o64._implicit_stage_local(oc, of, dtg, lambda s: None)
    gc, gf = N.cpu()
    t64 = FT == np.float64
    for k in range(5):
        assert rel(gc[:, k], oc[:, k]) <= (1e-12 if t64 else 2e-6), k
    assert rel(gf, of) <= (1e-10 if t64 else 5e-5)
    Yc0, Yf0 = sim.Y.cpu()
    for _ in range(3):  # eager step, graph capture, graph replay
        sim.Y = sim.to_device(Yc0, Yf0)
        sim.step(fused=True)
        torch.cuda.synchronize()
        gc, gf = sim.Y.cpu()
        if _ == 0:
            oc, of = o64.step(Yc0.astype(np.float64), Yf0.astype(np.float64))
        for k in range(5):
            assert rel(gc[:, k], oc[:, k]) <= (1e-11 if t64 else 1e-5), (_, k)
        assert rel(gf, of) <= (1e-11 if t64 else 5e-5)
    sim.close()
