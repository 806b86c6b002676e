"""The MOIST (EquilibriumMicrophysics0M) instantiations of the product kernels, executed on the CPU by the CTA emulator of tests/emu/
(see tests/test_kernels_cpu_emulation.py) and compared with the oracle's moist path: k5_imp_stage<…, MOIST> (fused stage and LDIV
mode), k5_exp_a<…, MOIST>, k5_tracer_a / the tracer parts of k7_exp_c with the active ρq_tot (water terms: part 1 of k7_exp_c), and the hook kernels k_cache_imp, k_t_imp2,
k_t_post_imp2 in a moist context.  The states carry cloudy points (q_0 raised), so the saturation-adjustment branch runs.
Test infrastructure only."""
import ctypes as C

import numpy as np
import pytest

from climaatmos_jl_b200 import grid as G, params as prm, setups
from oracle.dycore_oracle import Oracle
from tests.test_kernels_cpu_emulation import HG_N, HG_GI11, HG_GI12, HG_GI22, _full_hgeo, emu, emu5, emux, rel  # noqa: F401 (fixtures)

UPW = {"none": 0, "first_order": 1, "third_order": 2, "vanleer_limiter": 3}
p = lambda a: a.ctypes.data_as(C.c_void_p)


def moist_par(P):
    return np.array([P.R_v, P.cp_v, P.cp_l, P.cp_i, P.LH_v0, P.LH_s0, P.T_triple, P.press_triple, P.T_freeze, P.T_icenuc, P.pow_icenuc])


def make(deep, sponge, upw, ze, dzb, ntr=0, q_0=0.03, seed=5, rayleigh=None):
    P = prm.DycoreParams(zd_rayleigh=12000.0, zd_viscous=12000.0)
    g = G.make_sphere_grid(FT=np.float64, h_elem=2, z_elem=ze, z_max=30000.0, dz_bottom=dzb, radius=P.planet_radius, deep_atmosphere=deep)
    N = prm.DycoreNumerics(dt=250.0, rayleigh_sponge=sponge if rayleigh is None else rayleigh, viscous_sponge=sponge, energy_upwinding=upw,
                           microphysics_model="0M")
    o = Oracle(g, P, N, np.float64)
    Yc, Yf = setups.moist_baroclinic_wave(g, P, q_0=q_0)
    rng = np.random.default_rng(seed)
    Yc = Yc * (1 + 1e-3 * rng.standard_normal(Yc.shape))
    Yf = 0.3 * g.dz_f * rng.standard_normal(Yf.shape)
    Yc = np.ascontiguousarray(np.concatenate([Yc] + [Yc[:, :1] * 1e-2 * (1 + 0.5 * rng.random(Yc[:, :1].shape)) for _ in range(ntr)], axis=1))
    return P, g, N, o, Yc, np.ascontiguousarray(Yf), rng


def level_tables(g, P, o, deep, rayleigh, viscous):
    nv = g.nv
    s_c = (g.radius + g.z_c) / g.radius if deep else np.ones(nv)
    s_f = (g.radius + g.z_f) / g.radius if deep else np.ones(nv + 1)
    pad = lambda a: np.concatenate([np.asarray(a, dtype=np.float64), np.zeros(64 - len(a))])
    phic = P.grav * g.z_c
    dphif = np.zeros(nv + 1)
    dphif[1:-1] = phic[1:] - phic[:-1]
    z0 = np.zeros(nv + 1)
    return np.stack([pad(1 / s_c**2), pad(1 / s_f**2), pad(s_f), pad(g.dz_c), pad(g.dz_f), pad(s_c**2 * g.dz_c), pad(1 / (s_c**2 * g.dz_c)),
                     pad(1 / g.dz_f**2), pad(phic), pad(dphif),
                     pad(o.beta_rayleigh(g.z_f, P.alpha_rayleigh_w) if rayleigh else z0), pad(o.beta_rayleigh(g.z_c, P.alpha_rayleigh_uh) if rayleigh else z0[:-1]),
                     pad(o.beta_viscous(g.z_c) if viscous else z0[:-1]), pad(o.beta_viscous(g.z_f) if viscous else z0)])


@pytest.mark.parametrize("upw,rayleigh,deep,ze,dzb,ntr", [
    ("vanleer_limiter", True, True, 12, 400.0, 0), ("first_order", False, False, 12, 400.0, 1), ("none", False, True, 12, 400.0, 0),
    ("third_order", True, True, 63, 30.0, 0), ("vanleer_limiter", False, True, 5, 3000.0, 2),
])
def test_emulated_moist_implicit_stage_and_ldiv_match_oracle(emu5, upw, rayleigh, deep, ze, dzb, ntr):
    P, g, N, o, Yc, Yf, rng = make(deep, False, upw, ze, dzb, ntr, rayleigh=rayleigh)
    nh, ncf, nv = Yc.shape[0], Yc.shape[1], g.nv
    dtg = 0.4358665215084590 * N.dt
    vl = level_tables(g, P, o, deep, rayleigh, False)[:11]
    A = g.dxdxi
    Ginv = np.linalg.inv(np.einsum("...ab,...ac->...bc", A, A))
    hgeo = np.zeros((nh, HG_N, 16))
    hgeo[:, HG_GI11], hgeo[:, HG_GI12], hgeo[:, HG_GI22] = (Ginv[..., a, b].reshape(nh, 16) for a, b in ((0, 0), (0, 1), (1, 1)))
    sc = np.array([P.R_d, P.cp_d, P.cv_d, P.T_0, P.p_ref_theta, P.T_surf_ref, P.T_min_ref, P.T_min_sgs, N.dt, float(rayleigh), dtg, UPW[upw], ncf])
    mp = moist_par(P)
    try:
        assert emu5.emu_set_moist(p(mp), None) == 0
        Uc, Uf = Yc.copy(), Yf.copy()
        pc0 = o.set_implicit_precomputed_quantities(Yc.copy(), Yf.copy())
        assert ((pc0["ql"] + pc0["qi"]) > 0).mean() > 0.005  # cloudy points present: the Newton branch of the adjustment runs
        o._implicit_stage_local(Uc, Uf, dtg, lambda s: None)
        for fn in (emu5.emu_imp5, emu5.emu_imp8):  # shared-memory slab layout / warp-per-column-pair layout
            Nc, Nf = np.zeros_like(Yc), np.zeros_like(Yf)
            assert fn(nh, nv, p(sc), p(vl), p(hgeo), p(Yc), p(Yf), p(Nc), p(Nf)) == 0
            for k in range(ncf):
                assert rel(Nc[:, k], Uc[:, k]) < 1e-12, (k, rel(Nc[:, k], Uc[:, k]))
            for k in (0, 3, 4):
                assert rel(Nc[:, k] - Yc[:, k], Uc[:, k] - Yc[:, k]) < 1e-8, (k, rel(Nc[:, k] - Yc[:, k], Uc[:, k] - Yc[:, k]))
            assert rel(Nf, Uf) < 1e-10
        # LDIV mode on the Wfact snapshot
        Yf0 = Yf.copy()
        Yf0[..., 0] = 0
        Yf0[..., -1] = 0
        pc = o.set_implicit_precomputed_quantities(Yc.copy(), Yf0.copy())
        Jm = o.update_jacobian(Yc, Yf0, pc, dtg)
        Rc = np.ascontiguousarray(rng.standard_normal(Yc.shape) * np.abs(Yc).mean(axis=(0, 2, 3, 4), keepdims=True) * 1e-3)
        Rf = np.ascontiguousarray(rng.standard_normal(Yf.shape))
        dc, df = o.ldiv(Jm, Rc, Rf)
        for fn in (emu5.emu_ldiv5, emu5.emu_ldiv8):
            dYc, dYf = np.zeros_like(Yc), np.zeros_like(Yf)
            assert fn(nh, nv, p(sc), p(vl), p(hgeo), p(Yc), p(Yf0), p(Rc), p(Rf), p(dYc), p(dYf)) == 0
            for k in range(ncf):
                assert rel(dYc[:, k], dc[:, k]) < 1e-11, ("ldiv", k, rel(dYc[:, k], dc[:, k]))
            assert rel(dYf, df) < 1e-11
        # T_imp! and T_post_imp! of the hook path in the warp-per-column-pair layout
        Ytc, Ytf, Ypc, Ypf = np.full_like(Yc, 7.0), np.full_like(Yf, 7.0), np.full_like(Yc, 7.0), np.full_like(Yf, 7.0)
        assert emu5.emu_hooks8(nh, nv, p(sc), p(vl), p(hgeo), p(Yc), p(Yf0), p(Ytc), p(Ytf), p(Ypc), p(Ypf)) == 0
        tc, tf = o.implicit_tendency(Yc, Yf0, pc)
        for k in (0, 3, 4):
            assert rel(Ytc[:, k], tc[:, k]) < 1e-11, ("k8_t_imp", k, rel(Ytc[:, k], tc[:, k]))
        assert np.abs(Ytc[:, 1:3]).max() == 0 and not np.any(Ytc[:, 5:])
        assert rel(Ytf, tf) < 1e-9
        if upw != "none":
            qc, qf = o.correct_implicit_advection_tendency(Yc, Yf0, pc)
            for k in (3, 4):
                assert rel(Ypc[:, k], qc[:, k]) < 1e-10, ("k8_t_post_imp", k, rel(Ypc[:, k], qc[:, k]))
            assert np.abs(Ypc[:, :3]).max() == 0 and np.abs(Ypf).max() == 0
    finally:
        emu5.emu_set_moist(None, None)


@pytest.mark.parametrize("deep,sponge,ze,dzb,ntr", [(True, True, 12, 400.0, 0), (False, False, 12, 400.0, 1), (True, True, 63, 30.0, 0)])
def test_emulated_moist_explicit_tendency_kernels_match_oracle(emux, deep, sponge, ze, dzb, ntr):
    """T_exp_T_lim! of a moist context, kernel by kernel: k5_exp_a<MOIST> (+ k5_tracer_a) against the oracle's element-local
    `_rt_pre` / `_tracer_pre` / `_tracer_laplacians`, then k7_exp_c (incl. its water terms and tracer parts) against `_rt_post` / `_tracer_post`."""
    P, g, N, o, Yc, Yf, rng = make(deep, sponge, "vanleer_limiter", ze, dzb, ntr, seed=17)
    Yf[..., 0] = 0
    Yf[..., -1] = 0
    nh, ncf, nv = Yc.shape[0], Yc.shape[1], g.nv
    vl = level_tables(g, P, o, deep, sponge, sponge)
    hgeo = _full_hgeo(g, P, deep)
    sc = np.array([P.R_d, P.cp_d, P.cv_d, P.T_0, P.p_ref_theta, P.T_surf_ref, P.T_min_ref, P.T_min_sgs, N.dt, float(o.nu4_vort), float(o.nu4_scalar),
                   N.divergence_damping_factor, 1, float(sponge), float(sponge), 3, ncf, 3])
    Dm, wq = np.ascontiguousarray(g.D, dtype=np.float64), np.ascontiguousarray(g.wq, dtype=np.float64)
    mp = moist_par(P)
    Hw = np.zeros((nh, 16, nv))
    try:
        assert emux.emu_set_moist(p(mp), p(Hw)) == 0
        Ytc, Ytf, H, Ylc = np.zeros_like(Yc), np.zeros_like(Yf), np.zeros_like(Yc), np.zeros_like(Yc)
        assert emux.emu_exp5(0, nh, nv, p(sc), p(vl), p(Dm), p(wq), p(hgeo), p(Yc), p(Yf), p(Ytc), p(Ytf), p(H), p(Ylc)) == 0
        assert emux.emu_exp5(2, nh, nv, p(sc), p(vl), p(Dm), p(wq), p(hgeo), p(Yc), p(Yf), p(Ytc), p(Ytf), p(H), p(Ylc)) == 0
        pc = o.set_implicit_precomputed_quantities(Yc.copy(), Yf.copy())
        assert ((pc["ql"] + pc["qi"]) > 0).mean() > 0.005
        tc, tf, L = o._rt_pre(Yc, Yf, pc)
        lc = np.zeros_like(Yc)
        o._tracer_pre(tc, lc, Yc, Yf, pc)
        Lq = o._tracer_laplacians(Yc, pc)
        for k in range(ncf):
            ref = tc[:, k]
            if np.abs(ref).max() == 0:
                assert np.abs(Ytc[:, k]).max() == 0, k
            else:
                assert rel(Ytc[:, k], ref) < 1e-10, ("exp_a tendency", k, rel(Ytc[:, k], ref))
        for k in range(4, ncf):
            assert rel(Ylc[:, k], lc[:, k]) < 1e-10, ("T_lim", k)
        assert rel(Ytf, tf) < 1e-9
        for k in range(4):
            assert rel(H[:, k], L[k]) < 1e-10, ("laplacian", k, rel(H[:, k], L[k]))
        for k in range(4, ncf):
            assert rel(H[:, k], Lq[k - 4]) < 1e-9, ("tracer laplacian", k, rel(H[:, k], Lq[k - 4]))
        rho = Yc[:, 0]
        heff = rho * o.h_eff_plus_Phi(pc["T"], pc["qt"], pc["ql"], pc["qi"])
        assert rel(Hw.reshape(heff.shape), heff) < 1e-12
        # apply phase on the (un-DSSed) ∇² fields
        Hin = np.ascontiguousarray(np.stack(list(L) + list(Lq), axis=1))
        o._rt_post(tc, tf, Yc, L, pc, Lq[0])
        o._tracer_post(lc, Yc, Lq)
        # k7_exp_c parts 0-2 (part 1 carries the water terms of the moist context: ρ, ρq_tot of Yₜ_lim and the enthalpy flux), then its tracer parts
        assert emux.emu_exp5(1, nh, nv, p(sc), p(vl), p(Dm), p(wq), p(hgeo), p(Yc), p(Yf), p(Ytc), p(Ytf), p(Hin), p(Ylc)) == 0
        assert emux.emu_exp5(3, nh, nv, p(sc), p(vl), p(Dm), p(wq), p(hgeo), p(Yc), p(Yf), p(Ytc), p(Ytf), p(Hin), p(Ylc)) == 0
        for k in range(ncf):
            if np.abs(tc[:, k]).max() > 0:
                assert rel(Ytc[:, k], tc[:, k]) < 1e-10, ("apply", k, rel(Ytc[:, k], tc[:, k]))
        for k in [0] + list(range(4, ncf)):
            assert rel(Ylc[:, k], lc[:, k]) < 1e-9, ("apply T_lim", k, rel(Ylc[:, k], lc[:, k]))
        assert np.abs(Ylc[:, 1:4]).max() == 0
        assert rel(Ytf, tf) < 1e-9
    finally:
        emux.emu_set_moist(None, None)


@pytest.mark.parametrize("upw,deep", [("vanleer_limiter", True), ("first_order", False), ("third_order", True)])
def test_emulated_moist_hook_kernels_match_oracle(emu, upw, deep):
    """k_cache_imp, k_t_imp2 and k_t_post_imp2 in a moist context against the oracle's cache_imp!, T_imp! and T_post_imp!."""
    P, g, N, o, Yc, Yf, rng = make(deep, False, upw, 12, 400.0, 1, rayleigh=True)
    nh, ncf, nv = Yc.shape[0], Yc.shape[1], g.nv
    vl = level_tables(g, P, o, deep, True, False)[:11]
    hgeo = _full_hgeo(g, P, deep)
    sc = np.zeros(18)
    sc[:10] = [P.R_d, P.cp_d, P.cv_d, P.T_0, P.p_ref_theta, P.T_surf_ref, P.T_min_ref, P.T_min_sgs, N.dt, 1.0]
    sc[14], sc[16] = 0.4358665215084590 * N.dt, UPW[upw]
    mp = moist_par(P)
    z = lambda *s: np.zeros(s)
    Kc, Tc, pcc, hc = z(nh, 16, nv), z(nh, 16, nv), z(nh, 16, nv), z(nh, 16, nv)
    Ytc, Ytf, Ypc, Ypf = np.zeros_like(Yc), np.zeros_like(Yf), np.zeros_like(Yc), np.zeros_like(Yf)
    dummy = np.zeros(8)
    try:
        assert emu.emu_set_moist(p(mp), None) == 0
        Yf_k = Yf.copy()
        assert emu.emu_hooks(nh, nv, ncf, p(sc), p(vl), p(hgeo), p(Yc), p(Yf_k), None, None, p(Kc), p(Tc), p(pcc), p(hc), p(Ytc), p(Ytf), p(dummy),
                             None, None, p(Ypc), p(Ypf), None, None) == 0
        Yf_o = Yf.copy()
        pc = o.set_implicit_precomputed_quantities(Yc, Yf_o)
        assert np.array_equal(Yf_k, Yf_o)
        sh = pc["T"].shape
        assert rel(Tc.reshape(sh), pc["T"]) < 1e-13 and rel(pcc.reshape(sh), pc["p"]) < 1e-13 and rel(hc.reshape(sh), pc["h_tot"]) < 1e-13
        tc, tf = o.implicit_tendency(Yc, Yf_o, pc)
        for k in (0, 3, 4):
            assert rel(Ytc[:, k], tc[:, k]) < 1e-11, ("t_imp", k, rel(Ytc[:, k], tc[:, k]))
        assert np.abs(Ytc[:, 1:3]).max() == 0 and not np.any(Ytc[:, 5:])
        assert rel(Ytf, tf) < 1e-9
        qc, qf = o.correct_implicit_advection_tendency(Yc, Yf_o, pc)
        for k in (3, 4):
            assert rel(Ypc[:, k], qc[:, k]) < 1e-10, ("t_post_imp", k, rel(Ypc[:, k], qc[:, k]))
        assert np.abs(Ypc[:, :3]).max() == 0 and np.abs(Ypf).max() == 0
    finally:
        emu.emu_set_moist(None, None)
