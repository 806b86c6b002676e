"""The CUDA kernel sources of climaatmos.jl_b200/csrc, executed on the CPU: tests/emu/ compiles the kernel headers UNCHANGED with g++
against a stub cuda_runtime.h and runs each CTA as 256 cooperative fibers on one host thread — a fiber runs to its next barrier and
yields; __syncthreads(), per-warp barriers and an exchange slot per lane for warp shuffles / votes / __syncwarp (tests/emu/cuda_runtime.h) —
and the results are compared with the oracle (Float64
instantiations; a Float32 build for the slab / quarter-element kernels).  Covered: k5_exp_a, k_dss2, k7_exp_c, k5_imp_stage, k_axpy_dss (every kernel of the benchmarked step), k5_tracer_a and the tracer parts of k7_exp_c, the dry hook kernels (k_cache_imp, k_t_imp2, k_wfact2, k_t_post_imp2; ldiv! = k5_imp_stage in LDIV mode), and
the vertical-diffusion / limiter kernels of kernels_vdiff.cuh.

Test infrastructure only: it checks indexing, phase structure and arithmetic of the very code that runs on the B200, not timing and
not the packed-Float32 PTX specialisations.  It was written when the round's GPU budget was spent and is how the kernels of
kernels_vdiff.cuh were debugged before their GPU runs."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from climaatmos_jl_b200 import grid as G, params as prm, setups
from oracle.dycore_oracle import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))
HG_N, HG_GI11, HG_GI12, HG_GI22 = 23, 2, 3, 4


@pytest.fixture(scope="module")
def emu():
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = os.path.join(HERE, "emu", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libemu_vdiff.so")
    src = os.path.join(HERE, "emu", "emu_vdiff.cpp")
    csrc = os.path.join(os.path.dirname(HERE), "climaatmos.jl_b200", "csrc")
    subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", "-shared", "-fPIC", "-fvisibility=hidden", "-I", os.path.join(HERE, "emu"), "-I", csrc, src, "-o", so],
                   check=True)
    return C.CDLL(so)


@pytest.fixture(scope="module")
def emu32():
    """The same harness compiled with -DEMU_FT=float: the Float32 instantiations of the kernels."""
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = os.path.join(HERE, "emu", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libemu_vdiff_f32.so")
    csrc = os.path.join(os.path.dirname(HERE), "climaatmos.jl_b200", "csrc")
    subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", "-shared", "-fPIC", "-fvisibility=hidden", "-DEMU_FT=float", "-I", os.path.join(HERE, "emu"),
                    "-I", csrc, os.path.join(HERE, "emu", "emu_vdiff.cpp"), "-o", so], check=True)
    return C.CDLL(so)


def run_case(emu, vd, deep, dm, iters, ntr, rayleigh=False, ze=12, dzb=400.0):
    P = prm.DycoreParams(D_0_diffusion=150.0, H_diffusion=5000.0, C_E=0.0044, zd_rayleigh=12000.0)
    g = G.make_sphere_grid(FT=np.float64, h_elem=2, z_elem=ze, z_max=30000.0, dz_bottom=dzb, radius=P.planet_radius, deep_atmosphere=deep)
    N = prm.DycoreNumerics(dt=250.0, vert_diff=vd, implicit_diffusion=True, approximate_linear_solve_iters=iters,
                           disable_momentum_vertical_diffusion=dm, rayleigh_sponge=rayleigh)
    o = Oracle(g, P, N, np.float64)
    Yc, Yf = setups.dry_baroclinic_wave(g, P)
    rng = np.random.default_rng(1234)
    Yc = Yc * (1 + 1e-3 * rng.standard_normal(Yc.shape))
    Yf = 0.3 * g.dz_f * rng.standard_normal(Yf.shape)
    Yf[..., 0] = 0
    Yf[..., -1] = 0
    Yc = np.concatenate([Yc] + [Yc[:, :1] * 1e-2 * (1 + 0.5 * rng.random(Yc[:, :1].shape)) for _ in range(ntr)], axis=1)
    Yc, Yf = np.ascontiguousarray(Yc), np.ascontiguousarray(Yf)
    nh, ncf, nv = Yc.shape[0], Yc.shape[1], g.nv
    dtg = 0.4358665215084590 * N.dt
    # context constants as capi.cu:create_geo builds them
    s_c = (g.radius + g.z_c) / g.radius if deep else np.ones(nv)
    s_f = (g.radius + g.z_f) / g.radius if deep else np.ones(nv + 1)
    pad = lambda a: np.concatenate([np.asarray(a, dtype=np.float64), np.zeros(64 - len(a))])
    phic = P.grav * g.z_c
    dphif = np.zeros(nv + 1)
    dphif[1:-1] = phic[1:] - phic[:-1]
    brw = o.beta_rayleigh(g.z_f, P.alpha_rayleigh_w) if rayleigh else np.zeros(nv + 1)
    vl = np.stack([pad(1 / s_c**2), pad(1 / s_f**2), pad(s_f), pad(g.dz_c), pad(g.dz_f), pad(s_c**2 * g.dz_c), pad(1 / (s_c**2 * g.dz_c)),
                   pad(1 / g.dz_f**2), pad(phic), pad(dphif), pad(brw)])
    A = g.dxdxi
    Ginv = np.linalg.inv(np.einsum("...ab,...ac->...bc", A, A))
    hgeo = np.zeros((nh, HG_N, 16))
    hgeo[:, HG_GI11] = Ginv[..., 0, 0].reshape(nh, 16)
    hgeo[:, HG_GI12] = Ginv[..., 0, 1].reshape(nh, 16)
    hgeo[:, HG_GI22] = Ginv[..., 1, 1].reshape(nh, 16)
    kdec = pad(P.D_0_diffusion * np.exp(-(g.z_c - g.z_f[0]) / P.H_diffusion))
    mode = {"VerticalDiffusion": 1, "DecayWithHeightDiffusion": 2}[vd]
    sc = np.array([P.R_d, P.cp_d, P.cv_d, P.T_0, P.p_ref_theta, P.T_surf_ref, P.T_min_ref, P.T_min_sgs, N.dt, float(rayleigh), mode,
                   0 if dm else 1, iters, P.C_E * g.dz_c[0] / 2, dtg, 0, 0, 0])
    Rc = rng.standard_normal(Yc.shape) * np.abs(Yc).mean(axis=(0, 2, 3, 4), keepdims=True) * 1e-3
    Rf = rng.standard_normal(Yf.shape)
    Ytc = np.zeros_like(Yc)
    jac = np.zeros((nh, 15, 16, nv + 1))
    jacd = np.zeros((nh, 2, 16, nv + 1))
    dYc, dYf = np.zeros_like(Yc), np.zeros_like(Yf)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = emu.emu_vdiff(nh, nv, ncf, p(sc), p(vl), p(hgeo), p(kdec), p(Yc), p(Yf), p(Rc), p(Rf), p(Ytc), p(jac), p(jacd), p(dYc), p(dYf))
    assert rc == 0
    # oracle
    pc = o.set_implicit_precomputed_quantities(Yc[:, :4].copy(), Yf.copy())
    ot = np.zeros_like(Yc)
    o.vertical_diffusion_boundary_layer_tendency(ot, Yc, pc)
    Jm = o.update_jacobian(Yc, Yf, pc, dtg)
    oc, of = o.ldiv(Jm, Rc, Rf)
    return (Ytc, dYc, dYf), (ot, oc, of)


def rel(a, b):
    n = np.linalg.norm(b.ravel())
    return np.linalg.norm((a - b).ravel()) / n if n > 0 else np.linalg.norm(a.ravel())


@pytest.mark.parametrize("vd,deep,dm,iters,ntr,rayleigh", [
    ("DecayWithHeightDiffusion", True, False, 2, 1, False),
    ("VerticalDiffusion", True, False, 2, 1, True),
    ("DecayWithHeightDiffusion", False, False, 1, 2, False),
    ("DecayWithHeightDiffusion", True, True, 0, 0, False),
    ("VerticalDiffusion", False, True, 3, 1, False),
])
def test_emulated_vdiff_kernels_match_oracle(emu, vd, deep, dm, iters, ntr, rayleigh):
    (gt, gc, gf), (ot, oc, of) = run_case(emu, vd, deep, dm, iters, ntr, rayleigh)
    ncf = gt.shape[1]
    for k in range(ncf):
        if k == 0 or (dm and k in (1, 2)):
            assert np.all(gt[:, k] == 0) and np.all(ot[:, k] == 0)
        else:
            assert rel(gt[:, k], ot[:, k]) < 1e-11, ("tend", k, rel(gt[:, k], ot[:, k]))
    for k in range(ncf):
        assert rel(gc[:, k], oc[:, k]) < 1e-10, ("ldiv", k, rel(gc[:, k], oc[:, k]))
    assert rel(gf, of) < 1e-10, ("ldiv u3", rel(gf, of))


def test_emulated_vertical_mass_borrowing_kernel_matches_oracle(emu):
    """k_lim_vborrow (lim!, vertical_water_borrowing) on the CPU emulator against the oracle, bit for bit (same operation order)."""
    P = prm.DycoreParams()
    g = G.make_sphere_grid(FT=np.float64, h_elem=2, z_elem=12, z_max=30000.0, dz_bottom=400.0, radius=P.planet_radius)
    N = prm.DycoreNumerics(dt=100.0, tracer_nonnegativity_method="vertical_water_borrowing")
    o = Oracle(g, P, N, np.float64)
    Yc, Yf = setups.dry_baroclinic_wave(g, P)
    rng = np.random.default_rng(7)
    chi = [1e-3 * rng.standard_normal(Yc[:, 0].shape) + 4e-4, 1e-3 * rng.standard_normal(Yc[:, 0].shape) - 2e-4]
    Yc = np.ascontiguousarray(np.concatenate([Yc] + [(Yc[:, 0] * c)[:, None] for c in chi], axis=1))
    ref = Yc.copy()
    o.limiters_func(ref, Yc)
    got = Yc.copy()
    dzc = np.concatenate([g.dz_c, np.zeros(64 - g.nv)])
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert emu.emu_vborrow(Yc.shape[0], g.nv, Yc.shape[1], p(dzc), p(got)) == 0
    assert (got[:, 4:] >= 0).all() and (Yc[:, 4:] < 0).any()
    assert np.array_equal(got[:, :4], Yc[:, :4])
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("upw,rayleigh,deep", [("vanleer_limiter", True, True), ("first_order", False, False), ("none", False, True), ("third_order", True, True)])
def test_emulated_dry_hook_kernels_match_oracle(emu, upw, rayleigh, deep):
    """k_cache_imp, k_t_imp2, k_wfact2, k_t_post_imp2 (quarter element per CTA) on the CPU emulator
    against the oracle's cache_imp! / T_imp! / Wfact + ldiv! / T_post_imp! (Float64)."""
    P = prm.DycoreParams(zd_rayleigh=12000.0)
    g = G.make_sphere_grid(FT=np.float64, h_elem=2, z_elem=12, z_max=30000.0, dz_bottom=400.0, radius=P.planet_radius, deep_atmosphere=deep)
    N = prm.DycoreNumerics(dt=250.0, rayleigh_sponge=rayleigh, energy_upwinding=upw)
    o = Oracle(g, P, N, np.float64)
    Yc, Yf = setups.dry_baroclinic_wave(g, P)
    rng = np.random.default_rng(99)
    Yc = np.ascontiguousarray(Yc * (1 + 1e-3 * rng.standard_normal(Yc.shape)))
    Yf = np.ascontiguousarray(0.3 * g.dz_f * rng.standard_normal(Yf.shape))  # boundary faces non-zero: cache_imp! must filter them
    nh, ncf, nv = Yc.shape[0], Yc.shape[1], g.nv
    dtg = 0.4358665215084590 * N.dt
    s_c = (g.radius + g.z_c) / g.radius if deep else np.ones(nv)
    s_f = (g.radius + g.z_f) / g.radius if deep else np.ones(nv + 1)
    pad = lambda a: np.concatenate([np.asarray(a, dtype=np.float64), np.zeros(64 - len(a))])
    phic = P.grav * g.z_c
    dphif = np.zeros(nv + 1)
    dphif[1:-1] = phic[1:] - phic[:-1]
    brw = o.beta_rayleigh(g.z_f, P.alpha_rayleigh_w) if rayleigh else np.zeros(nv + 1)
    vl = np.stack([pad(1 / s_c**2), pad(1 / s_f**2), pad(s_f), pad(g.dz_c), pad(g.dz_f), pad(s_c**2 * g.dz_c), pad(1 / (s_c**2 * g.dz_c)),
                   pad(1 / g.dz_f**2), pad(phic), pad(dphif), pad(brw)])
    A = g.dxdxi
    Ginv = np.linalg.inv(np.einsum("...ab,...ac->...bc", A, A))
    hgeo = np.zeros((nh, HG_N, 16))
    hgeo[:, HG_GI11], hgeo[:, HG_GI12], hgeo[:, HG_GI22] = (Ginv[..., a, b].reshape(nh, 16) for a, b in ((0, 0), (0, 1), (1, 1)))
    sc = np.array([P.R_d, P.cp_d, P.cv_d, P.T_0, P.p_ref_theta, P.T_surf_ref, P.T_min_ref, P.T_min_sgs, N.dt, float(rayleigh), 0, 0, 0, 0, dtg,
                   0, {"none": 0, "first_order": 1, "third_order": 2, "vanleer_limiter": 3}[upw], 0])
    Rc = rng.standard_normal(Yc.shape) * np.abs(Yc).mean(axis=(0, 2, 3, 4), keepdims=True) * 1e-3
    Rf = rng.standard_normal(Yf.shape)
    z4 = lambda: np.zeros((nh, 16, nv))
    Kc, Tc, pc_, hc = z4(), z4(), z4(), z4()
    Ytc, Ytf, dYc, dYf, Ypc, Ypf, Sc, Sf = (np.zeros_like(a) for a in (Yc, Yf, Yc, Yf, Yc, Yf, Yc, Yf))
    jac = np.zeros((nh, 15, 16, nv + 1))
    gYf = Yf.copy()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert emu.emu_hooks(nh, nv, ncf, p(sc), p(vl), p(hgeo), p(Yc), p(gYf), p(Rc), p(Rf), p(Kc), p(Tc), p(pc_), p(hc), p(Ytc), p(Ytf), p(jac),
                         p(dYc), p(dYf), p(Ypc), p(Ypf), p(Sc), p(Sf)) == 0
    oc, of = Yc.copy(), Yf.copy()
    pc = o.set_implicit_precomputed_quantities(oc, of)
    assert np.array_equal(gYf, of) and np.all(gYf[..., 0] == 0) and np.all(gYf[..., -1] == 0)
    sh = Yc[:, 0].shape
    for got, key in ((Kc, "K"), (Tc, "T"), (pc_, "p"), (hc, "h_tot")):
        assert rel(got.reshape(sh), pc[key]) < 1e-13, key
    tc, tf = o.implicit_tendency(oc, of, pc)
    assert rel(Ytc[:, 0], tc[:, 0]) < 1e-12 and rel(Ytc[:, 3], tc[:, 3]) < 1e-12 and rel(Ytf, tf) < 1e-11
    assert np.all(Ytc[:, 1:3] == 0)
    # Wfact planes (k_wfact2: what the implicit-diffusion solve and b200_debug_jacobian consume) against the oracle's blocks; ldiv! of
    # the dry path is k5_imp_stage<…, LDIV> (test_emulated_default_fused_implicit_stage_matches_oracle)
    Jm = o.update_jacobian(oc, of, pc, dtg)
    for k, w in ((3, Jm["u3_rho"][0]), (4, Jm["u3_rho"][1]), (5, Jm["u3_rhoe"][0]), (6, Jm["u3_rhoe"][1]), (7, Jm["u3_uh"][0][0]), (10, Jm["u3_uh"][1][1])):
        assert rel(jac[:, k].reshape(nh, 4, 4, nv + 1), w) < 1e-11, k
    for k, w in ((11, Jm["rho_u3"][0]), (12, Jm["rho_u3"][1]), (13, Jm["rhoe_u3"][0]), (14, Jm["rhoe_u3"][1])):
        assert rel(jac[:, k].reshape(nh, 4, 4, nv + 1)[..., :-1], w) < 1e-11, k
    pc_c, pc_f = o.correct_implicit_advection_tendency(oc, of, pc)
    if upw != "none":  # with :none the host returns zeros without launching the kernel (capi.cu: impl_t_post)
        assert rel(Ypc[:, 3], pc_c[:, 3]) < 1e-10
    assert np.all(Ypc[:, :3] == 0) and np.all(Ypf == 0)


@pytest.mark.parametrize("ze,dzb", [(2, 15000.0), (3, 5000.0), (63, 30.0)])
def test_emulated_vdiff_kernels_at_the_column_height_limits(emu, ze, dzb):
    """Minimum (2, 3 levels) and maximum (63 levels = 64 faces, the LV = 64 limit of the slab layout) column heights."""
    (gt, gc, gf), (ot, oc, of) = run_case(emu, "VerticalDiffusion", True, False, 2, 1, ze=ze, dzb=dzb)
    for k in range(1, gt.shape[1]):
        assert rel(gt[:, k], ot[:, k]) < 1e-10, ("tend", k)
    for k in range(gc.shape[1]):
        assert rel(gc[:, k], oc[:, k]) < 1e-9, ("ldiv", k)
    assert rel(gf, of) < 1e-9


@pytest.mark.parametrize("vd,deep,dm,iters,ntr,upw,rayleigh,ze,dzb", [
    ("DecayWithHeightDiffusion", True, False, 2, 1, "vanleer_limiter", True, 12, 400.0),
    ("VerticalDiffusion", False, False, 1, 2, "first_order", False, 12, 400.0),
    ("DecayWithHeightDiffusion", True, False, 2, 1, "third_order", True, 12, 400.0),
    ("VerticalDiffusion", True, True, 3, 0, "none", False, 12, 400.0),
    ("DecayWithHeightDiffusion", True, False, 2, 1, "vanleer_limiter", True, 63, 30.0),
    ("DecayWithHeightDiffusion", True, False, 0, 1, "vanleer_limiter", False, 2, 15000.0),
])
def test_emulated_fused_implicit_diffusion_stage_matches_oracle(emu, emu5, vd, deep, dm, iters, ntr, upw, rayleigh, ze, dzb):
    """k_imp_stage_diff: cache_imp! → Wfact → T_imp! → ldiv! (approximate arrowhead iteration) → U −= ΔU → cache_imp! →
    T_post_imp! with implicit vertical diffusion in ONE kernel, against the oracle's hook sequence (Float64).  The input state has
    non-zero u₃ on the boundary faces: the kernel must treat them as zero like cache_imp!."""
    P = prm.DycoreParams(D_0_diffusion=150.0, H_diffusion=5000.0, C_E=0.0044, zd_rayleigh=12000.0)
    g = G.make_sphere_grid(FT=np.float64, h_elem=2, z_elem=ze, z_max=30000.0, dz_bottom=dzb, radius=P.planet_radius, deep_atmosphere=deep)
    N = prm.DycoreNumerics(dt=250.0, vert_diff=vd, implicit_diffusion=True, approximate_linear_solve_iters=iters,
                           disable_momentum_vertical_diffusion=dm, rayleigh_sponge=rayleigh, energy_upwinding=upw)
    o = Oracle(g, P, N, np.float64)
    Yc, Yf = setups.dry_baroclinic_wave(g, P)
    rng = np.random.default_rng(4321)
    Yc = Yc * (1 + 1e-3 * rng.standard_normal(Yc.shape))
    Yf = 0.3 * g.dz_f * rng.standard_normal(Yf.shape)
    Yc = np.ascontiguousarray(np.concatenate([Yc] + [Yc[:, :1] * 1e-2 * (1 + 0.5 * rng.random(Yc[:, :1].shape)) for _ in range(ntr)], axis=1))
    Yf = np.ascontiguousarray(Yf)
    nh, ncf, nv = Yc.shape[0], Yc.shape[1], g.nv
    dtg = 0.4358665215084590 * N.dt
    s_c = (g.radius + g.z_c) / g.radius if deep else np.ones(nv)
    s_f = (g.radius + g.z_f) / g.radius if deep else np.ones(nv + 1)
    pad = lambda a: np.concatenate([np.asarray(a, dtype=np.float64), np.zeros(64 - len(a))])
    phic = P.grav * g.z_c
    dphif = np.zeros(nv + 1)
    dphif[1:-1] = phic[1:] - phic[:-1]
    brw = o.beta_rayleigh(g.z_f, P.alpha_rayleigh_w) if rayleigh else np.zeros(nv + 1)
    vl = np.stack([pad(1 / s_c**2), pad(1 / s_f**2), pad(s_f), pad(g.dz_c), pad(g.dz_f), pad(s_c**2 * g.dz_c), pad(1 / (s_c**2 * g.dz_c)),
                   pad(1 / g.dz_f**2), pad(phic), pad(dphif), pad(brw)])
    A = g.dxdxi
    Ginv = np.linalg.inv(np.einsum("...ab,...ac->...bc", A, A))
    hgeo = np.zeros((nh, HG_N, 16))
    hgeo[:, HG_GI11], hgeo[:, HG_GI12], hgeo[:, HG_GI22] = (Ginv[..., a, b].reshape(nh, 16) for a, b in ((0, 0), (0, 1), (1, 1)))
    kdec = pad(P.D_0_diffusion * np.exp(-(g.z_c - g.z_f[0]) / P.H_diffusion))
    mode = {"VerticalDiffusion": 1, "DecayWithHeightDiffusion": 2}[vd]
    sc = np.array([P.R_d, P.cp_d, P.cv_d, P.T_0, P.p_ref_theta, P.T_surf_ref, P.T_min_ref, P.T_min_sgs, N.dt, float(rayleigh), mode,
                   0 if dm else 1, iters, P.C_E * g.dz_c[0] / 2, dtg, 0, {"none": 0, "first_order": 1, "third_order": 2, "vanleer_limiter": 3}[upw]])
    Nc, Nf = np.zeros_like(Yc), np.zeros_like(Yf)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert emu.emu_stage_diff(nh, nv, ncf, p(sc), p(vl), p(hgeo), p(kdec), p(Yc), p(Yf), p(Nc), p(Nf)) == 0
    Uc, Uf = Yc.copy(), Yf.copy()
    o._implicit_stage_local(Uc, Uf, dtg, lambda s: None)
    assert np.abs(Uc - Yc).max() > 0
    for k in range(ncf):
        assert rel(Nc[:, k], Uc[:, k]) < 1e-12, (k, rel(Nc[:, k], Uc[:, k]))
        assert rel(Nc[:, k] - Yc[:, k], Uc[:, k] - Yc[:, k]) < 1e-8 or np.abs(Uc[:, k] - Yc[:, k]).max() == 0, ("increment", k)
    assert rel(Nf, Uf) < 1e-10
    # the warp-per-column-pair version of the same stage (k8_imp_stage_diff, kernels_imp8d.cuh: what b200_implicit_stage launches)
    sc5 = np.array([P.R_d, P.cp_d, P.cv_d, P.T_0, P.p_ref_theta, P.T_surf_ref, P.T_min_ref, P.T_min_sgs, N.dt, float(rayleigh), dtg,
                    {"none": 0, "first_order": 1, "third_order": 2, "vanleer_limiter": 3}[upw], ncf, mode, 0 if dm else 1, iters, P.C_E * g.dz_c[0] / 2, 5])
    for layout in (8,):
        sc5[17] = layout
        N5c, N5f = np.zeros_like(Yc), np.zeros_like(Yf)
        assert emu5.emu_imp5d(nh, nv, p(sc5), p(vl), p(hgeo), p(kdec), p(Yc), p(Yf), p(N5c), p(N5f)) == 0
        for k in range(ncf):
            assert rel(N5c[:, k], Uc[:, k]) < 1e-12, (layout, k, rel(N5c[:, k], Uc[:, k]))
            assert rel(N5c[:, k] - Yc[:, k], Uc[:, k] - Yc[:, k]) < 1e-8 or np.abs(Uc[:, k] - Yc[:, k]).max() == 0, (layout, "increment", k, rel(N5c[:, k] - Yc[:, k], Uc[:, k] - Yc[:, k]))
        assert rel(N5f, Uf) < 1e-10, (layout, "u3", rel(N5f, Uf))


@pytest.mark.parametrize("vd,upw", [("DecayWithHeightDiffusion", "vanleer_limiter"), ("VerticalDiffusion", "first_order")])
def test_float32_instantiations_of_the_vdiff_kernels(emu32, vd, upw):
    """Float32 builds of the vertical-diffusion kernels (k_vdiff_jac → k_ldiv_diff, k_vdiff_tend2, k_imp_stage_diff) on the CPU emulator
    against the Float64 oracle evaluated on the same Float32-rounded inputs: the errors must sit inside the tolerances those GPU tests use
    (ldiv! 5e-5; fused stage 2e-6 on the centre fields, 2e-4 on u₃)."""
    F = np.float32
    P = prm.DycoreParams(D_0_diffusion=150.0, H_diffusion=5000.0, C_E=0.0044)
    g = G.make_sphere_grid(FT=np.float64, h_elem=2, z_elem=12, z_max=30000.0, dz_bottom=400.0, radius=P.planet_radius)
    N = prm.DycoreNumerics(dt=250.0, vert_diff=vd, implicit_diffusion=True, approximate_linear_solve_iters=2, energy_upwinding=upw)
    o = Oracle(g, P, N, np.float64)
    Yc, Yf = setups.dry_baroclinic_wave(g, P)
    rng = np.random.default_rng(11)
    Yc = Yc * (1 + 1e-3 * rng.standard_normal(Yc.shape))
    Yf = 0.3 * g.dz_f * rng.standard_normal(Yf.shape)
    Yf[..., 0] = 0
    Yf[..., -1] = 0
    Yc = np.concatenate([Yc, Yc[:, :1] * 1e-2 * (1 + 0.5 * rng.random(Yc[:, :1].shape))], axis=1)
    Yc32, Yf32 = np.ascontiguousarray(Yc, dtype=F), np.ascontiguousarray(Yf, dtype=F)
    Yc, Yf = Yc32.astype(np.float64), Yf32.astype(np.float64)
    nh, ncf, nv = Yc.shape[0], Yc.shape[1], g.nv
    dtg = 0.4358665215084590 * N.dt
    s_c, s_f = (g.radius + g.z_c) / g.radius, (g.radius + g.z_f) / g.radius
    pad = lambda a: np.concatenate([np.asarray(a, dtype=np.float64), np.zeros(64 - len(a))])
    phic = P.grav * g.z_c
    dphif = np.zeros(nv + 1)
    dphif[1:-1] = (phic.astype(F)[1:] - phic.astype(F)[:-1]).astype(np.float64)  # as capi.cu: difference of the rounded values
    vl = np.stack([pad(1 / s_c**2), pad(1 / s_f**2), pad(s_f), pad(g.dz_c), pad(g.dz_f), pad(s_c**2 * g.dz_c), pad(1 / (s_c**2 * g.dz_c)),
                   pad(1 / g.dz_f**2), pad(phic), pad(dphif), pad(np.zeros(nv + 1))]).astype(F)
    A = g.dxdxi
    Ginv = np.linalg.inv(np.einsum("...ab,...ac->...bc", A, A))
    hgeo = np.zeros((nh, HG_N, 16), dtype=F)
    hgeo[:, HG_GI11], hgeo[:, HG_GI12], hgeo[:, HG_GI22] = (Ginv[..., a, b].reshape(nh, 16) for a, b in ((0, 0), (0, 1), (1, 1)))
    kdec = pad(P.D_0_diffusion * np.exp(-(g.z_c - g.z_f[0]) / P.H_diffusion)).astype(F)
    mode = {"VerticalDiffusion": 1, "DecayWithHeightDiffusion": 2}[vd]
    sc = np.array([P.R_d, P.cp_d, P.cv_d, P.T_0, P.p_ref_theta, P.T_surf_ref, P.T_min_ref, P.T_min_sgs, N.dt, 0.0, mode, 1, 2,
                   P.C_E * g.dz_c[0] / 2, dtg, 2, {"first_order": 1, "vanleer_limiter": 3}[upw], 2])
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    # fused stage
    Nc, Nf = np.zeros_like(Yc32), np.zeros_like(Yf32)
    assert emu32.emu_stage_diff(nh, nv, ncf, p(sc), p(vl), p(hgeo), p(kdec), p(Yc32), p(Yf32), p(Nc), p(Nf)) == 0
    Uc, Uf = Yc.copy(), Yf.copy()
    o._implicit_stage_local(Uc, Uf, dtg, lambda s: None)
    errs = [rel(Nc[:, k].astype(np.float64), Uc[:, k]) for k in range(ncf)] + [rel(Nf.astype(np.float64), Uf)]
    assert max(errs[:ncf]) < 2e-6 and errs[-1] < 2e-4, errs
    # Wfact planes (k_wfact2, k_vdiff_jac) → ldiv! (k_ldiv_diff) with the tendency kernel k_vdiff_tend2
    sc2 = sc.copy()
    sc2[16] = 0
    Rc = (rng.standard_normal(Yc.shape) * np.abs(Yc).mean(axis=(0, 2, 3, 4), keepdims=True) * 1e-3).astype(F)
    Rf = rng.standard_normal(Yf.shape).astype(F)
    Ytc, dYc, dYf = np.zeros_like(Yc32), np.zeros_like(Yc32), np.zeros_like(Yf32)
    jac, jacd = np.zeros((nh, 15, 16, nv + 1), dtype=F), np.zeros((nh, 2, 16, nv + 1), dtype=F)
    assert emu32.emu_vdiff(nh, nv, ncf, p(sc2), p(vl), p(hgeo), p(kdec), p(Yc32), p(Yf32), p(Rc), p(Rf), p(Ytc), p(jac), p(jacd), p(dYc), p(dYf)) == 0
    pc = o.set_implicit_precomputed_quantities(Yc[:, :4].copy(), Yf.copy())
    Jm = o.update_jacobian(Yc, Yf, pc, dtg)
    oc, of = o.ldiv(Jm, Rc.astype(np.float64), Rf.astype(np.float64))
    e2 = [rel(dYc[:, k].astype(np.float64), oc[:, k]) for k in range(ncf)] + [rel(dYf.astype(np.float64), of)]
    assert max(e2) < 5e-5, e2
    ot = np.zeros_like(Yc)
    o.vertical_diffusion_boundary_layer_tendency(ot, Yc, pc)
    e3 = [rel(Ytc[:, k].astype(np.float64), ot[:, k]) for k in range(1, ncf)]
    assert max(e3) < 1e-4, e3


@pytest.fixture(scope="module")
def emu5():
    """tests/emu/emu_imp5.cpp: the DEFAULT fused implicit-stage kernel k5_imp_stage (Float64 instantiation, PCR solve) on the CTA emulator."""
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = os.path.join(HERE, "emu", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libemu_imp5.so")
    csrc = os.path.join(os.path.dirname(HERE), "climaatmos.jl_b200", "csrc")
    subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", "-shared", "-fPIC", "-fvisibility=hidden", "-I", os.path.join(HERE, "emu"), "-I", csrc,
                    os.path.join(HERE, "emu", "emu_imp5.cpp"), "-o", so], check=True)
    return C.CDLL(so)


@pytest.mark.parametrize("upw,rayleigh,deep,ze,dzb,ntr", [
    ("vanleer_limiter", True, True, 12, 400.0, 0), ("first_order", False, False, 12, 400.0, 1), ("none", False, True, 12, 400.0, 0),
    ("vanleer_limiter", True, True, 63, 30.0, 0), ("vanleer_limiter", False, True, 2, 15000.0, 0), ("vanleer_limiter", False, True, 5, 3000.0, 2),
    ("third_order", True, True, 12, 400.0, 0), ("third_order", False, True, 3, 8000.0, 1), ("third_order", False, False, 63, 30.0, 0), ("third_order", False, True, 2, 15000.0, 0),
])
def test_emulated_default_fused_implicit_stage_matches_oracle(emu5, upw, rayleigh, deep, ze, dzb, ntr):
    """k5_imp_stage<double, PCR> — the implicit-stage kernel of the benchmarked step (packed row layout, parallel cyclic reduction) — run on
    the CPU from its unchanged source against the oracle's cache_imp! → Wfact → T_imp! → ldiv! → U −= ΔU → cache_imp! → T_post_imp!."""
    P = prm.DycoreParams(zd_rayleigh=12000.0)
    g = G.make_sphere_grid(FT=np.float64, h_elem=2, z_elem=ze, z_max=30000.0, dz_bottom=dzb, radius=P.planet_radius, deep_atmosphere=deep)
    N = prm.DycoreNumerics(dt=250.0, rayleigh_sponge=rayleigh, energy_upwinding=upw)
    o = Oracle(g, P, N, np.float64)
    Yc, Yf = setups.dry_baroclinic_wave(g, P)
    rng = np.random.default_rng(5)
    Yc = Yc * (1 + 1e-3 * rng.standard_normal(Yc.shape))
    Yf = 0.3 * g.dz_f * rng.standard_normal(Yf.shape)  # non-zero boundary faces: the kernel filters them on load
    Yc = np.ascontiguousarray(np.concatenate([Yc] + [Yc[:, :1] * 1e-2 * (1 + 0.5 * rng.random(Yc[:, :1].shape)) for _ in range(ntr)], axis=1))
    Yf = np.ascontiguousarray(Yf)
    nh, ncf, nv = Yc.shape[0], Yc.shape[1], g.nv
    dtg = 0.4358665215084590 * N.dt
    s_c = (g.radius + g.z_c) / g.radius if deep else np.ones(nv)
    s_f = (g.radius + g.z_f) / g.radius if deep else np.ones(nv + 1)
    pad = lambda a: np.concatenate([np.asarray(a, dtype=np.float64), np.zeros(64 - len(a))])
    phic = P.grav * g.z_c
    dphif = np.zeros(nv + 1)
    dphif[1:-1] = phic[1:] - phic[:-1]
    brw = o.beta_rayleigh(g.z_f, P.alpha_rayleigh_w) if rayleigh else np.zeros(nv + 1)
    vl = np.stack([pad(1 / s_c**2), pad(1 / s_f**2), pad(s_f), pad(g.dz_c), pad(g.dz_f), pad(s_c**2 * g.dz_c), pad(1 / (s_c**2 * g.dz_c)),
                   pad(1 / g.dz_f**2), pad(phic), pad(dphif), pad(brw)])
    A = g.dxdxi
    Ginv = np.linalg.inv(np.einsum("...ab,...ac->...bc", A, A))
    hgeo = np.zeros((nh, HG_N, 16))
    hgeo[:, HG_GI11], hgeo[:, HG_GI12], hgeo[:, HG_GI22] = (Ginv[..., a, b].reshape(nh, 16) for a, b in ((0, 0), (0, 1), (1, 1)))
    sc = np.array([P.R_d, P.cp_d, P.cv_d, P.T_0, P.p_ref_theta, P.T_surf_ref, P.T_min_ref, P.T_min_sgs, N.dt, float(rayleigh), dtg,
                   {"none": 0, "first_order": 1, "third_order": 2, "vanleer_limiter": 3}[upw], ncf])
    Nc, Nf = np.zeros_like(Yc), np.zeros_like(Yf)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert emu5.emu_imp5(nh, nv, p(sc), p(vl), p(hgeo), p(Yc), p(Yf), p(Nc), p(Nf)) == 0
    Uc, Uf = Yc.copy(), Yf.copy()
    o._implicit_stage_local(Uc, Uf, dtg, lambda s: None)
    for k in range(ncf):
        assert rel(Nc[:, k], Uc[:, k]) < 1e-12, (k, rel(Nc[:, k], Uc[:, k]))
    assert rel(Nc[:, 0] - Yc[:, 0], Uc[:, 0] - Yc[:, 0]) < 1e-8 and rel(Nc[:, 3] - Yc[:, 3], Uc[:, 3] - Yc[:, 3]) < 1e-8
    assert rel(Nf, Uf) < 1e-10
    # the warp-per-column-pair version (k8_imp_stage, kernels_imp8.cuh: shuffle neighbours and shuffle PCR, no shared memory)
    N8c, N8f = np.zeros_like(Yc), np.zeros_like(Yf)
    assert emu5.emu_imp8(nh, nv, p(sc), p(vl), p(hgeo), p(Yc), p(Yf), p(N8c), p(N8f)) == 0
    for k in range(ncf):
        assert rel(N8c[:, k], Uc[:, k]) < 1e-12, ("k8", k, rel(N8c[:, k], Uc[:, k]))
    assert rel(N8c[:, 0] - Yc[:, 0], Uc[:, 0] - Yc[:, 0]) < 1e-8 and rel(N8c[:, 3] - Yc[:, 3], Uc[:, 3] - Yc[:, 3]) < 1e-8
    assert rel(N8f, Uf) < 1e-10, ("k8 u3", rel(N8f, Uf))
    # ldiv! of the hook path = the same kernel in LDIV mode on the state Wfact saw (b200_wfact keeps a snapshot): ΔY = J(Y, dtγ)⁻¹ R for a
    # random right-hand side with ALL components non-zero (R_uₕ ≠ 0 exercises the (u₃, uₕ) blocks, R_ρχ the tracer fallback block)
    Yf0 = Yf.copy()
    Yf0[..., 0] = 0
    Yf0[..., -1] = 0  # Wfact is called on a filtered state (cache_imp! precedes it)
    pc = o.set_implicit_precomputed_quantities(Yc[:, :4].copy(), Yf0.copy())
    Jm = o.update_jacobian(Yc, Yf0, pc, dtg)
    Rc = np.ascontiguousarray(rng.standard_normal(Yc.shape) * np.abs(Yc).mean(axis=(0, 2, 3, 4), keepdims=True) * 1e-3)
    Rf = np.ascontiguousarray(rng.standard_normal(Yf.shape))
    dc, df = o.ldiv(Jm, Rc, Rf)
    dYc, dYf = np.zeros_like(Yc), np.zeros_like(Yf)
    assert emu5.emu_ldiv5(nh, nv, p(sc), p(vl), p(hgeo), p(Yc), p(Yf0), p(Rc), p(Rf), p(dYc), p(dYf)) == 0
    for k in range(ncf):
        assert rel(dYc[:, k], dc[:, k]) < 1e-11, ("ldiv", k, rel(dYc[:, k], dc[:, k]))
    assert rel(dYf, df) < 1e-11, ("ldiv u3", rel(dYf, df))
    dYc, dYf = np.zeros_like(Yc), np.zeros_like(Yf)
    assert emu5.emu_ldiv8(nh, nv, p(sc), p(vl), p(hgeo), p(Yc), p(Yf0), p(Rc), p(Rf), p(dYc), p(dYf)) == 0
    for k in range(ncf):
        assert rel(dYc[:, k], dc[:, k]) < 1e-11, ("ldiv8", k, rel(dYc[:, k], dc[:, k]))
    assert rel(dYf, df) < 1e-11, ("ldiv8 u3", rel(dYf, df))
    # T_imp! and T_post_imp! of the hook path in the same layout (k8_t_imp, k8_t_post_imp)
    Ytc, Ytf, Ypc, Ypf = np.full_like(Yc, 7.0), np.full_like(Yf, 7.0), np.full_like(Yc, 7.0), np.full_like(Yf, 7.0)
    assert emu5.emu_hooks8(nh, nv, p(sc), p(vl), p(hgeo), p(Yc), p(Yf0), p(Ytc), p(Ytf), p(Ypc), p(Ypf)) == 0
    tc, tf = o.implicit_tendency(Yc, Yf0, pc)
    for k in (0, 3):
        assert rel(Ytc[:, k], tc[:, k]) < 1e-11, ("k8_t_imp", k, rel(Ytc[:, k], tc[:, k]))
    assert np.abs(Ytc[:, 1:3]).max() == 0 and not np.any(Ytc[:, 4:])
    assert rel(Ytf, tf) < 1e-9, ("k8_t_imp u3", rel(Ytf, tf))
    if upw != "none":
        qc, qf = o.correct_implicit_advection_tendency(Yc, Yf0, pc)
        assert rel(Ypc[:, 3], qc[:, 3]) < 1e-10, ("k8_t_post_imp", rel(Ypc[:, 3], qc[:, 3]))
        assert np.abs(Ypc[:, :3]).max() == 0 and not np.any(Ypc[:, 4:]) and np.abs(Ypf).max() == 0


@pytest.fixture(scope="module")
def emux():
    """tests/emu/emu_exp5.cpp: k5_exp_a / k7_exp_c (Float64 instantiations) on the CTA emulator with emulated warp shuffles."""
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = os.path.join(HERE, "emu", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libemu_exp5.so")
    csrc = os.path.join(os.path.dirname(HERE), "climaatmos.jl_b200", "csrc")
    subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", "-shared", "-fPIC", "-fvisibility=hidden", "-I", os.path.join(HERE, "emu"), "-I", csrc,
                    os.path.join(HERE, "emu", "emu_exp5.cpp"), "-o", so], check=True)
    return C.CDLL(so)


def _full_hgeo(g, P, deep):
    """hgeo[h][HG_N][16] as capi.cu:create_geo fills the components the element kernels stage (J2 … COS2)."""
    nh = g.J2.shape[0]
    A = g.dxdxi.reshape(nh, 16, 2, 2)
    a00, a01, a10, a11 = A[..., 0, 0], A[..., 0, 1], A[..., 1, 0], A[..., 1, 1]
    gc11, gc12, gc22 = a00 * a00 + a10 * a10, a00 * a01 + a10 * a11, a01 * a01 + a11 * a11
    det, dA = gc11 * gc22 - gc12 * gc12, a00 * a11 - a01 * a10
    lat = np.radians(g.lat.reshape(nh, 16))
    fv, fw = 2 * P.Omega * np.cos(lat), 2 * P.Omega * np.sin(lat)
    ai01, ai11 = -a01 / dA, a00 / dA
    J2 = g.J2.reshape(nh, 16)
    hg = np.zeros((nh, HG_N, 16))
    for k, val in enumerate([J2, 1 / J2, gc22 / det, -gc12 / det, gc11 / det, gc11, gc12, gc22, ai01 * fv if deep else 0 * fv,
                             ai11 * fv if deep else 0 * fv, fw, np.sin(lat) ** 2, np.cos(lat) ** 2]):
        hg[:, k] = val
    return hg


@pytest.mark.parametrize("deep,sponge,ze,dzb", [(True, True, 12, 400.0), (False, False, 12, 400.0), (True, True, 63, 30.0), (True, False, 3, 8000.0)])
def test_emulated_explicit_tendency_kernels_match_oracle(emux, deep, sponge, ze, dzb):
    """k5_exp_a and k7_exp_c — the two explicit-tendency kernels of the benchmarked step (packed row layout, ξ² contractions by warp
    shuffles) — run on the CPU from their unchanged source (Float64 instantiation): the pre-DSS tendencies and ∇² fields against the
    oracle's element-local `_rt_pre`, the hyperdiffusion apply against `_rt_post` on the same ∇² fields."""
    P = prm.DycoreParams(zd_rayleigh=12000.0, zd_viscous=12000.0)
    g = G.make_sphere_grid(FT=np.float64, h_elem=2, z_elem=ze, z_max=30000.0, dz_bottom=dzb, radius=P.planet_radius, deep_atmosphere=deep)
    N = prm.DycoreNumerics(dt=250.0, rayleigh_sponge=sponge, viscous_sponge=sponge)
    o = Oracle(g, P, N, np.float64)
    Yc, Yf = setups.dry_baroclinic_wave(g, P)
    rng = np.random.default_rng(17)
    Yc = np.ascontiguousarray(Yc * (1 + 1e-3 * rng.standard_normal(Yc.shape)))
    Yf = 0.3 * g.dz_f * rng.standard_normal(Yf.shape)
    Yf[..., 0] = 0
    Yf[..., -1] = 0
    Yf = np.ascontiguousarray(Yf)
    nh, nv = Yc.shape[0], g.nv
    s_c = (g.radius + g.z_c) / g.radius if deep else np.ones(nv)
    s_f = (g.radius + g.z_f) / g.radius if deep else np.ones(nv + 1)
    pad = lambda a: np.concatenate([np.asarray(a, dtype=np.float64), np.zeros(64 - len(a))])
    phic = P.grav * g.z_c
    dphif = np.zeros(nv + 1)
    dphif[1:-1] = phic[1:] - phic[:-1]
    z0 = np.zeros(nv + 1)
    vl = np.stack([pad(1 / s_c**2), pad(1 / s_f**2), pad(s_f), pad(g.dz_c), pad(g.dz_f), pad(s_c**2 * g.dz_c), pad(1 / (s_c**2 * g.dz_c)),
                   pad(1 / g.dz_f**2), pad(phic), pad(dphif),
                   pad(o.beta_rayleigh(g.z_f, P.alpha_rayleigh_w) if sponge else z0), pad(o.beta_rayleigh(g.z_c, P.alpha_rayleigh_uh) if sponge else z0[:-1]),
                   pad(o.beta_viscous(g.z_c) if sponge else z0[:-1]), pad(o.beta_viscous(g.z_f) if sponge else z0)])
    hgeo = _full_hgeo(g, P, deep)
    sc = np.array([P.R_d, P.cp_d, P.cv_d, P.T_0, P.p_ref_theta, P.T_surf_ref, P.T_min_ref, P.T_min_sgs, N.dt, float(o.nu4_vort), float(o.nu4_scalar),
                   N.divergence_damping_factor, 1, float(sponge), float(sponge), 3, 4, 3])
    Dm, wq = np.ascontiguousarray(g.D, dtype=np.float64), np.ascontiguousarray(g.wq, dtype=np.float64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    Ytc, Ytf, H = np.zeros_like(Yc), np.zeros_like(Yf), np.zeros_like(Yc)
    assert emux.emu_exp5(0, nh, nv, p(sc), p(vl), p(Dm), p(wq), p(hgeo), p(Yc), p(Yf), p(Ytc), p(Ytf), p(H), None) == 0
    pc = o.set_implicit_precomputed_quantities(Yc.copy(), Yf.copy())
    tc, tf, L = o._rt_pre(Yc, Yf, pc)
    for k in range(4):
        assert rel(Ytc[:, k], tc[:, k]) < 1e-10, ("exp_a tendency", k, rel(Ytc[:, k], tc[:, k]))
    assert rel(Ytf, tf) < 1e-9, ("exp_a u3 tendency", rel(Ytf, tf))
    for k in range(4):
        assert rel(H[:, k], L[k]) < 1e-10, ("exp_a laplacian", k, rel(H[:, k], L[k]))
    # hyperdiffusion apply on the (un-DSSed) ∇² fields: any H is a valid input for an element-local comparison
    Hin = np.ascontiguousarray(np.stack(L, axis=1))
    o._rt_post(tc, tf, Yc, L)
    assert emux.emu_exp5(1, nh, nv, p(sc), p(vl), p(Dm), p(wq), p(hgeo), p(Yc), p(Yf), p(Ytc), p(Ytf), p(Hin), None) == 0
    for k in range(4):
        assert rel(Ytc[:, k], tc[:, k]) < 1e-10, ("exp_c", k, rel(Ytc[:, k], tc[:, k]))
    assert rel(Ytf, tf) < 1e-9, ("exp_c u3", rel(Ytf, tf))


@pytest.mark.parametrize("tupw,sponge,ze,dzb", [("vanleer_limiter", True, 12, 400.0), ("first_order", False, 12, 400.0), ("none", True, 63, 30.0),
                                                ("vanleer_limiter", False, 3, 8000.0), ("third_order", True, 12, 400.0), ("third_order", False, 3, 8000.0),
                                                ("third_order", False, 4, 6000.0)])
def test_emulated_tracer_kernels_match_oracle(emux, tupw, sponge, ze, dzb):
    """k5_tracer_a / parts 3.. of k7_exp_c (passive tracers: horizontal advection into Yₜ_lim, ∇²χ, explicit vertical transport with tracer_upwinding,
    viscous sponge; tracer hyperdiffusion) on the CPU emulator against the oracle's `_tracer_pre`, `_tracer_laplacians`, `_tracer_post`."""
    P = prm.DycoreParams(zd_rayleigh=12000.0, zd_viscous=12000.0)
    g = G.make_sphere_grid(FT=np.float64, h_elem=2, z_elem=ze, z_max=30000.0, dz_bottom=dzb, radius=P.planet_radius)
    N = prm.DycoreNumerics(dt=250.0, rayleigh_sponge=sponge, viscous_sponge=sponge, tracer_upwinding=tupw)
    o = Oracle(g, P, N, np.float64)
    Yc, Yf = setups.dry_baroclinic_wave(g, P)
    rng = np.random.default_rng(23)
    Yc = Yc * (1 + 1e-3 * rng.standard_normal(Yc.shape))
    Yf = 0.3 * g.dz_f * rng.standard_normal(Yf.shape)
    Yf[..., 0] = 0
    Yf[..., -1] = 0
    chis = [1e-2 * (1 + 0.5 * rng.random(Yc[:, 0].shape)), 0.5 * (1 + np.sin(np.radians(g.lat))[..., None] * np.exp(-o.c.z / 8000.0))]
    Yc = np.ascontiguousarray(np.concatenate([Yc] + [(Yc[:, 0] * c)[:, None] for c in chis], axis=1))
    Yf = np.ascontiguousarray(Yf)
    nh, nv, ncf = Yc.shape[0], g.nv, Yc.shape[1]
    s_c, s_f = (g.radius + g.z_c) / g.radius, (g.radius + g.z_f) / g.radius
    pad = lambda a: np.concatenate([np.asarray(a, dtype=np.float64), np.zeros(64 - len(a))])
    phic = P.grav * g.z_c
    dphif = np.zeros(nv + 1)
    dphif[1:-1] = phic[1:] - phic[:-1]
    z0 = np.zeros(nv + 1)
    vl = np.stack([pad(1 / s_c**2), pad(1 / s_f**2), pad(s_f), pad(g.dz_c), pad(g.dz_f), pad(s_c**2 * g.dz_c), pad(1 / (s_c**2 * g.dz_c)),
                   pad(1 / g.dz_f**2), pad(phic), pad(dphif),
                   pad(o.beta_rayleigh(g.z_f, P.alpha_rayleigh_w) if sponge else z0), pad(o.beta_rayleigh(g.z_c, P.alpha_rayleigh_uh) if sponge else z0[:-1]),
                   pad(o.beta_viscous(g.z_c) if sponge else z0[:-1]), pad(o.beta_viscous(g.z_f) if sponge else z0)])
    hgeo = _full_hgeo(g, P, True)
    sc = np.array([P.R_d, P.cp_d, P.cv_d, P.T_0, P.p_ref_theta, P.T_surf_ref, P.T_min_ref, P.T_min_sgs, N.dt, float(o.nu4_vort), float(o.nu4_scalar),
                   N.divergence_damping_factor, 1, float(sponge), float(sponge), 3, ncf, {"none": 0, "first_order": 1, "third_order": 2, "vanleer_limiter": 3}[tupw]])
    Dm, wq = np.ascontiguousarray(g.D, dtype=np.float64), np.ascontiguousarray(g.wq, dtype=np.float64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    Ytc, Ylc, H, Ytf = np.zeros_like(Yc), np.zeros_like(Yc), np.zeros_like(Yc), np.zeros_like(Yf)
    assert emux.emu_exp5(2, nh, nv, p(sc), p(vl), p(Dm), p(wq), p(hgeo), p(Yc), p(Yf), p(Ytc), p(Ytf), p(H), p(Ylc)) == 0
    pc = o.set_implicit_precomputed_quantities(Yc[:, :4].copy(), Yf.copy())
    tc, lc = np.zeros_like(Yc), np.zeros_like(Yc)
    o._tracer_pre(tc, lc, Yc, Yf, pc)
    Lq = o._tracer_laplacians(Yc)
    for k, q in enumerate(range(4, ncf)):
        assert rel(Ylc[:, q], lc[:, q]) < 1e-10, ("T_lim", q, rel(Ylc[:, q], lc[:, q]))
        assert rel(Ytc[:, q], tc[:, q]) < 1e-9, ("T_exp", q, rel(Ytc[:, q], tc[:, q]))
        assert rel(H[:, q], Lq[k]) < 1e-10, ("laplacian", q)
    Hin = np.zeros_like(Yc)
    for k, q in enumerate(range(4, ncf)):
        Hin[:, q] = Lq[k]
    o._tracer_post(lc, Yc, Lq)
    assert emux.emu_exp5(3, nh, nv, p(sc), p(vl), p(Dm), p(wq), p(hgeo), p(Yc), p(Yf), p(Ytc), p(Ytf), p(Hin), p(Ylc)) == 0
    for q in range(4, ncf):
        assert rel(Ylc[:, q], lc[:, q]) < 1e-10, ("T_lim after hyperdiffusion", q, rel(Ylc[:, q], lc[:, q]))


@pytest.fixture(scope="module")
def emud():
    """tests/emu/emu_dss.cpp: k_dss2 (Float64 instantiation, single-rank path) on the CTA emulator."""
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = os.path.join(HERE, "emu", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libemu_dss.so")
    csrc = os.path.join(os.path.dirname(HERE), "climaatmos.jl_b200", "csrc")
    subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", "-shared", "-fPIC", "-fvisibility=hidden", "-I", os.path.join(HERE, "emu"), "-I", csrc,
                    os.path.join(HERE, "emu", "emu_dss.cpp"), "-o", so], check=True)
    return C.CDLL(so)


@pytest.mark.parametrize("he,ze", [(2, 5), (3, 12), (4, 63)])
def test_emulated_state_dss_matches_oracle(emud, he, ze):
    """k_dss2 — the weighted DSS of the state (ρ and ρe_tot as scalars, uₕ as a Covariant12 vector summed in the local physical basis, u₃
    on faces) with the per-node records of capi.cu:create_geo — on the CPU emulator against the oracle's Spaces.weighted_dss! restatement,
    including the pole and cube-corner elements (3-member vertices)."""
    HG_DSSW, HG_A00, HG_AI00 = 13, 14, 18
    P = prm.DycoreParams()
    g = G.make_sphere_grid(FT=np.float64, h_elem=he, z_elem=ze, z_max=30000.0, dz_bottom=30.0 if ze == 63 else 500.0, radius=P.planet_radius)
    o = Oracle(g, P, prm.DycoreNumerics(), np.float64)
    Yc, Yf = setups.dry_baroclinic_wave(g, P)
    rng = np.random.default_rng(31)
    Yc = np.ascontiguousarray(Yc * (1 + 1e-2 * rng.standard_normal(Yc.shape)))  # discontinuous across elements
    Yf = np.ascontiguousarray(0.3 * g.dz_f * rng.standard_normal(Yf.shape))
    nh, nv = Yc.shape[0], g.nv
    offs, mem = G.dss_node_csr(g.topology, 4)
    off = np.ascontiguousarray(offs, dtype=np.int32)
    m32 = np.ascontiguousarray(mem[:, 0] * 16 + mem[:, 2] * 4 + mem[:, 1], dtype=np.int32)  # (elem, i, j) → elem·16 + j·4 + i
    A = g.dxdxi.reshape(nh, 16, 2, 2)
    dA = A[..., 0, 0] * A[..., 1, 1] - A[..., 0, 1] * A[..., 1, 0]
    hgeo = np.zeros((nh, HG_N, 16))
    hgeo[:, HG_DSSW] = o.dss_w.reshape(nh, 16)
    hgeo[:, HG_A00], hgeo[:, HG_A00 + 1], hgeo[:, HG_A00 + 2], hgeo[:, HG_A00 + 3] = A[..., 0, 0], A[..., 0, 1], A[..., 1, 0], A[..., 1, 1]
    hgeo[:, HG_AI00], hgeo[:, HG_AI00 + 1] = A[..., 1, 1] / dA, -A[..., 0, 1] / dA
    hgeo[:, HG_AI00 + 2], hgeo[:, HG_AI00 + 3] = -A[..., 1, 0] / dA, A[..., 0, 0] / dA
    gc, gf = Yc.copy(), Yf.copy()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert emud.emu_dss_state(nh, nv, len(off) - 1, p(off), p(m32), p(hgeo), p(gc), p(gf)) == 0
    oc, of = Yc.copy(), Yf.copy()
    o.dss_state(oc, of)
    assert np.abs(oc - Yc).max() > 0
    for k in range(4):
        assert rel(gc[:, k], oc[:, k]) < 1e-14, (k, rel(gc[:, k], oc[:, k]))
    assert rel(gf, of) < 1e-14
    # interior nodes untouched, and the result is continuous: a second DSS changes nothing beyond round-off
    assert np.array_equal(gc[:, :, 1:3, 1:3], Yc[:, :, 1:3, 1:3])
    g2c, g2f = gc.copy(), gf.copy()
    assert emud.emu_dss_state(nh, nv, len(off) - 1, p(off), p(m32), p(hgeo), p(g2c), p(g2f)) == 0
    assert rel(g2c, gc) < 1e-14 and rel(g2f, gf) < 1e-14


@pytest.mark.parametrize("he,ze,ntr,dmask", [(2, 5, 0, 0), (3, 12, 1, 0b001), (2, 63, 0, 0b011)])
def test_emulated_fused_increment_dss_matches_oracle(emud, he, ze, ntr, dmask):
    """k_axpy_dss<3 terms> — the stage increment U = u + Σ cₖTₖ (terms flagged in dmask enter as cₖ·(Tₖ − u): the stage-solution form),
    the u₃ boundary filter and the state DSS in one kernel — on the CPU emulator against NumPy increment + filter + the oracle's dss!."""
    HG_DSSW, HG_A00, HG_AI00 = 13, 14, 18
    P = prm.DycoreParams()
    g = G.make_sphere_grid(FT=np.float64, h_elem=he, z_elem=ze, z_max=30000.0, dz_bottom=30.0 if ze == 63 else 500.0, radius=P.planet_radius)
    o = Oracle(g, P, prm.DycoreNumerics(), np.float64)
    Yc, Yf = setups.dry_baroclinic_wave(g, P)
    rng = np.random.default_rng(37)
    Yc = np.concatenate([Yc] + [Yc[:, :1] * 1e-2 * (1 + 0.5 * rng.random(Yc[:, :1].shape)) for _ in range(ntr)], axis=1)
    base_c = np.ascontiguousarray(Yc * (1 + 1e-2 * rng.standard_normal(Yc.shape)))
    base_f = np.ascontiguousarray(0.3 * g.dz_f * rng.standard_normal(Yf.shape))
    nh, nv, ncf = Yc.shape[0], g.nv, Yc.shape[1]
    T = []
    for k in range(3):
        if dmask >> k & 1:  # a stage solution close to the base state
            T.append((np.ascontiguousarray(base_c * (1 + 1e-3 * rng.standard_normal(Yc.shape))), np.ascontiguousarray(base_f + 0.01 * g.dz_f * rng.standard_normal(Yf.shape))))
        else:  # a tendency
            T.append((np.ascontiguousarray(1e-4 * base_c * rng.standard_normal(Yc.shape)), np.ascontiguousarray(1e-3 * g.dz_f * rng.standard_normal(Yf.shape))))
    coef = np.array([0.37 if dmask & 1 else 120.0, -0.21 if dmask & 2 else 55.0, 80.0])
    offs, mem = G.dss_node_csr(g.topology, 4)
    off = np.ascontiguousarray(offs, dtype=np.int32)
    m32 = np.ascontiguousarray(mem[:, 0] * 16 + mem[:, 2] * 4 + mem[:, 1], dtype=np.int32)
    A = g.dxdxi.reshape(nh, 16, 2, 2)
    dA = A[..., 0, 0] * A[..., 1, 1] - A[..., 0, 1] * A[..., 1, 0]
    hgeo = np.zeros((nh, HG_N, 16))
    hgeo[:, HG_DSSW] = o.dss_w.reshape(nh, 16)
    hgeo[:, HG_A00], hgeo[:, HG_A00 + 1], hgeo[:, HG_A00 + 2], hgeo[:, HG_A00 + 3] = A[..., 0, 0], A[..., 0, 1], A[..., 1, 0], A[..., 1, 1]
    hgeo[:, HG_AI00], hgeo[:, HG_AI00 + 1] = A[..., 1, 1] / dA, -A[..., 0, 1] / dA
    hgeo[:, HG_AI00 + 2], hgeo[:, HG_AI00 + 3] = -A[..., 1, 0] / dA, A[..., 0, 0] / dA
    out_c, out_f = np.full_like(base_c, np.nan), np.full_like(base_f, np.nan)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert emud.emu_axpy_dss3(nh, nv, ncf, len(off) - 1, p(off), p(m32), p(hgeo), p(base_c), p(base_f), p(T[0][0]), p(T[0][1]), p(T[1][0]), p(T[1][1]),
                              p(T[2][0]), p(T[2][1]), p(coef), C.c_uint(dmask), p(out_c), p(out_f)) == 0
    Uc, Uf = base_c.copy(), base_f.copy()
    for k in range(3):
        Uc += coef[k] * ((T[k][0] - base_c) if dmask >> k & 1 else T[k][0])
        Uf += coef[k] * ((T[k][1] - base_f) if dmask >> k & 1 else T[k][1])
    Uf[..., 0] = 0
    Uf[..., -1] = 0
    o.dss_state(Uc, Uf)
    assert np.isfinite(out_c).all() and np.isfinite(out_f).all()  # every point written exactly once
    for k in range(ncf):
        assert rel(out_c[:, k], Uc[:, k]) < 1e-13, (k, rel(out_c[:, k], Uc[:, k]))
    assert rel(out_f, Uf) < 1e-12
    assert np.all(out_f[..., 0] == 0) and np.all(out_f[..., -1] == 0)


@pytest.mark.parametrize("deep,sponge,he,ze,dzb", [(True, True, 2, 12, 400.0), (False, False, 3, 5, 3000.0)])
def test_emulated_t_exp_composite_matches_oracle(emux, emud, deep, sponge, he, ze, dzb):
    """T_exp_T_lim! as the library launches it — k5_exp_a → k_dss2 of the ∇² fields → k7_exp_c — entirely on the CPU emulator, against the
    oracle's remaining_tendency! (which includes its own weighted DSS of the ∇² fields)."""
    HG_DSSW, HG_A00, HG_AI00 = 13, 14, 18
    P = prm.DycoreParams(zd_rayleigh=12000.0, zd_viscous=12000.0)
    g = G.make_sphere_grid(FT=np.float64, h_elem=he, z_elem=ze, z_max=30000.0, dz_bottom=dzb, radius=P.planet_radius, deep_atmosphere=deep)
    N = prm.DycoreNumerics(dt=250.0, rayleigh_sponge=sponge, viscous_sponge=sponge)
    o = Oracle(g, P, N, np.float64)
    Yc, Yf = setups.dry_baroclinic_wave(g, P)
    rng = np.random.default_rng(41)
    Yc = Yc * (1 + 1e-3 * rng.standard_normal(Yc.shape))
    Yf = 0.3 * g.dz_f * rng.standard_normal(Yf.shape)
    o.dss_state(Yc, Yf)  # a continuous state, as T_exp sees it
    Yf[..., 0] = 0
    Yf[..., -1] = 0
    Yc, Yf = np.ascontiguousarray(Yc), np.ascontiguousarray(Yf)
    nh, nv = Yc.shape[0], g.nv
    s_c = (g.radius + g.z_c) / g.radius if deep else np.ones(nv)
    s_f = (g.radius + g.z_f) / g.radius if deep else np.ones(nv + 1)
    pad = lambda a: np.concatenate([np.asarray(a, dtype=np.float64), np.zeros(64 - len(a))])
    phic = P.grav * g.z_c
    dphif = np.zeros(nv + 1)
    dphif[1:-1] = phic[1:] - phic[:-1]
    z0 = np.zeros(nv + 1)
    vl = np.stack([pad(1 / s_c**2), pad(1 / s_f**2), pad(s_f), pad(g.dz_c), pad(g.dz_f), pad(s_c**2 * g.dz_c), pad(1 / (s_c**2 * g.dz_c)),
                   pad(1 / g.dz_f**2), pad(phic), pad(dphif),
                   pad(o.beta_rayleigh(g.z_f, P.alpha_rayleigh_w) if sponge else z0), pad(o.beta_rayleigh(g.z_c, P.alpha_rayleigh_uh) if sponge else z0[:-1]),
                   pad(o.beta_viscous(g.z_c) if sponge else z0[:-1]), pad(o.beta_viscous(g.z_f) if sponge else z0)])
    hgeo = _full_hgeo(g, P, deep)
    A = g.dxdxi.reshape(nh, 16, 2, 2)
    dA = A[..., 0, 0] * A[..., 1, 1] - A[..., 0, 1] * A[..., 1, 0]
    hgeo[:, HG_DSSW] = o.dss_w.reshape(nh, 16)
    hgeo[:, HG_A00], hgeo[:, HG_A00 + 1], hgeo[:, HG_A00 + 2], hgeo[:, HG_A00 + 3] = A[..., 0, 0], A[..., 0, 1], A[..., 1, 0], A[..., 1, 1]
    hgeo[:, HG_AI00], hgeo[:, HG_AI00 + 1] = A[..., 1, 1] / dA, -A[..., 0, 1] / dA
    hgeo[:, HG_AI00 + 2], hgeo[:, HG_AI00 + 3] = -A[..., 1, 0] / dA, A[..., 0, 0] / dA
    sc = np.array([P.R_d, P.cp_d, P.cv_d, P.T_0, P.p_ref_theta, P.T_surf_ref, P.T_min_ref, P.T_min_sgs, N.dt, float(o.nu4_vort), float(o.nu4_scalar),
                   N.divergence_damping_factor, 1, float(sponge), float(sponge), 3, 4, 3])
    Dm, wq = np.ascontiguousarray(g.D, dtype=np.float64), np.ascontiguousarray(g.wq, dtype=np.float64)
    offs, mem = G.dss_node_csr(g.topology, 4)
    off = np.ascontiguousarray(offs, dtype=np.int32)
    m32 = np.ascontiguousarray(mem[:, 0] * 16 + mem[:, 2] * 4 + mem[:, 1], dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    Ytc, Ytf, H = np.zeros_like(Yc), np.zeros_like(Yf), np.zeros_like(Yc)
    assert emux.emu_exp5(0, nh, nv, p(sc), p(vl), p(Dm), p(wq), p(hgeo), p(Yc), p(Yf), p(Ytc), p(Ytf), p(H), None) == 0
    assert emud.emu_dss_h(nh, nv, 4, len(off) - 1, p(off), p(m32), p(hgeo), p(H)) == 0
    assert emux.emu_exp5(1, nh, nv, p(sc), p(vl), p(Dm), p(wq), p(hgeo), p(Yc), p(Yf), p(Ytc), p(Ytf), p(H), None) == 0
    pc = o.set_implicit_precomputed_quantities(Yc.copy(), Yf.copy())
    tc, tf = o.remaining_tendency(Yc, Yf, pc)
    for k in range(4):
        assert rel(Ytc[:, k], tc[:, k]) < 1e-10, (k, rel(Ytc[:, k], tc[:, k]))
    assert rel(Ytf, tf) < 1e-9


@pytest.mark.parametrize("vdiff", [None, "explicit", "implicit"])
def test_emulated_fused_step_matches_oracle_step(emux, emud, emu5, emu, vdiff):
    """One ARS343 step assembled from the emulated PRODUCT kernels in the data flow of the fused stepper (capi.cu: impl_step, fused path):
    stage-solution form of the increments (U_i = u + Σ α_ij (N_j − u) + dt Σ β_ij T_exp[j], k_axpy_dss with dmask), k5_imp_stage → k_dss2,
    T_exp = k5_exp_a → k_dss2(∇²) → k7_exp_c, and the stiffly-accurate final increment from N₄ — against the oracle's LITERAL step
    (u + dt Σ bⱼ (T_exp[j] + T_imp[j]) with T_imp formed explicitly).  The host orchestration below restates impl_step's coefficient
    recursion; the arithmetic on the fields is all done by the kernels' own source.  With vertical diffusion: explicit → k_vdiff_tend2
    after k7_exp_c; implicit → k_imp_stage_diff in place of k5_imp_stage (the data flow of the fused stepper)."""
    HG_DSSW, HG_A00, HG_AI00 = 13, 14, 18
    P = prm.DycoreParams(zd_rayleigh=12000.0, zd_viscous=12000.0, D_0_diffusion=40.0, H_diffusion=6000.0)
    g = G.make_sphere_grid(FT=np.float64, h_elem=2, z_elem=8, z_max=30000.0, dz_bottom=500.0, radius=P.planet_radius)
    N = prm.DycoreNumerics(dt=300.0, rayleigh_sponge=True, viscous_sponge=True, vert_diff="VerticalDiffusion" if vdiff else None,
                           implicit_diffusion=(vdiff == "implicit"), approximate_linear_solve_iters=2)
    o = Oracle(g, P, N, np.float64)
    Yc, Yf = setups.dry_baroclinic_wave(g, P)
    Yc, Yf = np.ascontiguousarray(Yc), np.ascontiguousarray(Yf)
    nh, nv, ncf = Yc.shape[0], g.nv, 4
    dt = N.dt
    s_c, s_f = (g.radius + g.z_c) / g.radius, (g.radius + g.z_f) / g.radius
    pad = lambda a: np.concatenate([np.asarray(a, dtype=np.float64), np.zeros(64 - len(a))])
    phic = P.grav * g.z_c
    dphif = np.zeros(nv + 1)
    dphif[1:-1] = phic[1:] - phic[:-1]
    vl = np.stack([pad(1 / s_c**2), pad(1 / s_f**2), pad(s_f), pad(g.dz_c), pad(g.dz_f), pad(s_c**2 * g.dz_c), pad(1 / (s_c**2 * g.dz_c)),
                   pad(1 / g.dz_f**2), pad(phic), pad(dphif), pad(o.beta_rayleigh(g.z_f, P.alpha_rayleigh_w)),
                   pad(o.beta_rayleigh(g.z_c, P.alpha_rayleigh_uh)), pad(o.beta_viscous(g.z_c)), pad(o.beta_viscous(g.z_f))])
    hgeo = _full_hgeo(g, P, True)
    A = g.dxdxi.reshape(nh, 16, 2, 2)
    dA = A[..., 0, 0] * A[..., 1, 1] - A[..., 0, 1] * A[..., 1, 0]
    hgeo[:, HG_DSSW] = o.dss_w.reshape(nh, 16)
    hgeo[:, HG_A00], hgeo[:, HG_A00 + 1], hgeo[:, HG_A00 + 2], hgeo[:, HG_A00 + 3] = A[..., 0, 0], A[..., 0, 1], A[..., 1, 0], A[..., 1, 1]
    hgeo[:, HG_AI00], hgeo[:, HG_AI00 + 1] = A[..., 1, 1] / dA, -A[..., 0, 1] / dA
    hgeo[:, HG_AI00 + 2], hgeo[:, HG_AI00 + 3] = -A[..., 1, 0] / dA, A[..., 0, 0] / dA
    sc_exp = np.array([P.R_d, P.cp_d, P.cv_d, P.T_0, P.p_ref_theta, P.T_surf_ref, P.T_min_ref, P.T_min_sgs, dt, float(o.nu4_vort), float(o.nu4_scalar),
                       N.divergence_damping_factor, 1, 1.0, 1.0, 3, 4, 3])
    Dm, wq = np.ascontiguousarray(g.D, dtype=np.float64), np.ascontiguousarray(g.wq, dtype=np.float64)
    offs, mem = G.dss_node_csr(g.topology, 4)
    off = np.ascontiguousarray(offs, dtype=np.int32)
    m32 = np.ascontiguousarray(mem[:, 0] * 16 + mem[:, 2] * 4 + mem[:, 1], dtype=np.int32)
    nn = len(off) - 1
    p = lambda a: a.ctypes.data_as(C.c_void_p)

    def t_exp(Uc, Uf):
        Tc, Tf, H = np.zeros_like(Uc), np.zeros_like(Uf), np.zeros_like(Uc)
        assert emux.emu_exp5(0, nh, nv, p(sc_exp), p(vl), p(Dm), p(wq), p(hgeo), p(Uc), p(Uf), p(Tc), p(Tf), p(H), None) == 0
        assert emud.emu_dss_h(nh, nv, 4, nn, p(off), p(m32), p(hgeo), p(H)) == 0
        assert emux.emu_exp5(1, nh, nv, p(sc_exp), p(vl), p(Dm), p(wq), p(hgeo), p(Uc), p(Uf), p(Tc), p(Tf), p(H), None) == 0
        if vdiff == "explicit":  # k_vdiff_tend2 accumulates into Yₜ.c (the solver part of emu_vdiff is not used here)
            z = lambda a: np.zeros_like(a)
            jac, jacd = np.zeros((nh, 15, 16, nv + 1)), np.zeros((nh, 2, 16, nv + 1))
            assert emu.emu_vdiff(nh, nv, ncf, p(sc_vd), p(vl[:11]), p(hgeo), p(kdec), p(Uc), p(Uf), p(z(Uc)), p(z(Uf)), p(Tc), p(jac), p(jacd),
                                 p(z(Uc)), p(z(Uf))) == 0
        return Tc, Tf

    kdec = pad(P.D_0_diffusion * np.exp(-(g.z_c - g.z_f[0]) / P.H_diffusion))
    sc_vd = np.array([P.R_d, P.cp_d, P.cv_d, P.T_0, P.p_ref_theta, P.T_surf_ref, P.T_min_ref, P.T_min_sgs, dt, 1.0, 1, 1, 2, P.C_E * g.dz_c[0] / 2,
                      1.0, 2, 3, 1])

    def axpy_dss(base, terms, coefs, dmask):
        n = len(terms)
        Tc = (C.c_void_p * n)(*[t[0].ctypes.data for t in terms])
        Tf = (C.c_void_p * n)(*[t[1].ctypes.data for t in terms])
        cf = np.array(coefs, dtype=np.float64)
        oc, of = np.empty_like(base[0]), np.empty_like(base[1])
        assert emud.emu_axpy_dss_n(n, nh, nv, ncf, nn, p(off), p(m32), p(hgeo), p(base[0]), p(base[1]), Tc, Tf, p(cf), C.c_uint(dmask), p(oc), p(of)) == 0
        return oc, of

    def imp_stage(U, dtg):
        sc = np.array([P.R_d, P.cp_d, P.cv_d, P.T_0, P.p_ref_theta, P.T_surf_ref, P.T_min_ref, P.T_min_sgs, dt, 1.0, dtg, 3, ncf])
        Nc, Nf = np.zeros_like(U[0]), np.zeros_like(U[1])
        if vdiff == "implicit":
            scd = sc_vd.copy()
            scd[14] = dtg
            assert emu.emu_stage_diff(nh, nv, ncf, p(scd), p(vl[:11]), p(hgeo), p(kdec), p(U[0]), p(U[1]), p(Nc), p(Nf)) == 0
        else:
            assert emu5.emu_imp5(nh, nv, p(sc), p(vl[:11]), p(hgeo), p(U[0]), p(U[1]), p(Nc), p(Nf)) == 0
        assert emud.emu_dss_state(nh, nv, nn, p(off), p(m32), p(hgeo), p(Nc), p(Nf)) == 0
        return Nc, Nf

    ae, ai, be, bi, gam = prm.ars343()
    # stage-solution coefficients (capi.cu: impl_step, zform)
    al, bt = np.zeros((4, 4)), np.zeros((4, 4))
    for i in range(1, 4):
        for j in range(i):
            bt[i, j] = ae[i][j]
        for j in range(1, i):
            if ai[i][j] == 0:
                continue
            w = ai[i][j] / ai[j][j]
            al[i, j] += w
            for k in range(j):
                al[i, k] -= w * al[j, k]
                bt[i, k] -= w * bt[j, k]
    u = (Yc, Yf)
    Ns, Te = [None] * 4, [None] * 4
    Te[0] = t_exp(*u)
    for i in range(1, 4):
        terms, coefs, dmask = [], [], 0
        for j in range(1, i):
            if al[i, j] != 0:
                dmask |= 1 << len(terms)
                terms.append(Ns[j])
                coefs.append(al[i, j])
        for j in range(i):
            if bt[i, j] != 0:
                terms.append(Te[j])
                coefs.append(dt * bt[i, j])
        U = axpy_dss(u, terms, coefs, dmask)
        Ns[i] = imp_stage(U, dt * ai[i][i])
        Te[i] = t_exp(*Ns[i])
    terms = [Te[j] for j in range(4) if be[j] - ae[3][j] != 0]
    coefs = [dt * (be[j] - ae[3][j]) for j in range(4) if be[j] - ae[3][j] != 0]
    new_c, new_f = axpy_dss(Ns[3], terms, coefs, 0)
    oc, of = o.step(Yc.copy(), Yf.copy())
    for k in range(4):
        assert rel(new_c[:, k], oc[:, k]) < 1e-11, (k, rel(new_c[:, k], oc[:, k]))
    assert rel(new_f, of) < 1e-9, rel(new_f, of)
    assert rel(new_c - Yc, oc - Yc) < 1e-7  # the step increment itself, not only the state


def test_emulated_sem_quasimonotone_limiter_matches_oracle(emu):
    """k_lim_bounds / k_lim_apply (lim!, apply_sem_quasimonotone_limiter) on the CPU emulator against the oracle's restatement of
    Limiters.QuasiMonotoneLimiter: bounds from the reference state widened over the vertex neighbours, clip-and-redistribute per slab."""
    HG_WJ = 22
    P = prm.DycoreParams()
    g = G.make_sphere_grid(FT=np.float64, h_elem=3, z_elem=7, z_max=30000.0, dz_bottom=500.0, radius=P.planet_radius)
    N = prm.DycoreNumerics(dt=100.0, apply_sem_quasimonotone_limiter=True)
    o = Oracle(g, P, N, np.float64)
    Yc, _ = setups.dry_baroclinic_wave(g, P)
    rng = np.random.default_rng(43)
    rho = Yc[:, 0]
    ref = np.ascontiguousarray(np.concatenate([Yc, (rho * (0.5 + 0.3 * rng.random(rho.shape)))[:, None], (rho * 1e-2 * rng.random(rho.shape))[:, None]], axis=1))
    Y = ref.copy()
    Y[:, 4] = rho * (0.5 + 0.6 * (rng.random(rho.shape) - 0.3))  # over- and undershoots relative to the reference bounds
    Y[:, 5] = rho * 1e-2 * (rng.random(rho.shape) * 1.5 - 0.2)
    Y = np.ascontiguousarray(Y)
    nh, nv, ncf = Y.shape[0], g.nv, Y.shape[1]
    nbrs = o.neighboring_elements()
    nbr_off = np.zeros(nh + 1, dtype=np.int32)
    nbr_off[1:] = np.cumsum([len(x) for x in nbrs])
    nbr = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.int32) for x in nbrs]))
    hgeo = np.zeros((nh, HG_N, 16))
    hgeo[:, HG_WJ] = (g.W * g.J2).reshape(nh, 16)
    got = Y.copy()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert emu.emu_sem_limiter(nh, nv, ncf, p(nbr_off), p(nbr), p(hgeo), p(ref), p(got)) == 0
    want = Y.copy()
    o.limiters_func(want, ref)
    assert np.abs(want[:, 4:] - Y[:, 4:]).max() > 0
    assert np.array_equal(got[:, :4], Y[:, :4])
    for q in (4, 5):
        assert rel(got[:, q], want[:, q]) < 1e-13, (q, rel(got[:, q], want[:, q]))


@pytest.mark.parametrize("ntr", [3, 4])
def test_emulated_generic_dss_matches_oracle(emud, ntr):
    """The generic k_dss — impl_dss falls back to it for states with more than two tracers (n_tracers = 3, 4: 7 and 8 DSS items) — on the CPU
    emulator against the oracle's dss! of the state and a scalar weighted DSS of each tracer."""
    HG_DSSW, HG_A00, HG_AI00 = 13, 14, 18
    P = prm.DycoreParams()
    g = G.make_sphere_grid(FT=np.float64, h_elem=2, z_elem=6, z_max=30000.0, dz_bottom=500.0, radius=P.planet_radius)
    o = Oracle(g, P, prm.DycoreNumerics(), np.float64)
    Yc, Yf = setups.dry_baroclinic_wave(g, P)
    rng = np.random.default_rng(47)
    Yc = np.concatenate([Yc] + [Yc[:, :1] * 1e-2 * (k + 1) * rng.random(Yc[:, :1].shape) for k in range(ntr)], axis=1)
    Yc = np.ascontiguousarray(Yc * (1 + 1e-2 * rng.standard_normal(Yc.shape)))
    Yf = np.ascontiguousarray(0.3 * g.dz_f * rng.standard_normal(Yf.shape))
    nh, nv, ncf = Yc.shape[0], g.nv, Yc.shape[1]
    offs, mem = G.dss_node_csr(g.topology, 4)
    off = np.ascontiguousarray(offs, dtype=np.int32)
    m32 = np.ascontiguousarray(mem[:, 0] * 16 + mem[:, 2] * 4 + mem[:, 1], dtype=np.int32)
    A = g.dxdxi.reshape(nh, 16, 2, 2)
    dA = A[..., 0, 0] * A[..., 1, 1] - A[..., 0, 1] * A[..., 1, 0]
    hgeo = np.zeros((nh, HG_N, 16))
    hgeo[:, HG_DSSW] = o.dss_w.reshape(nh, 16)
    hgeo[:, HG_A00], hgeo[:, HG_A00 + 1], hgeo[:, HG_A00 + 2], hgeo[:, HG_A00 + 3] = A[..., 0, 0], A[..., 0, 1], A[..., 1, 0], A[..., 1, 1]
    hgeo[:, HG_AI00], hgeo[:, HG_AI00 + 1] = A[..., 1, 1] / dA, -A[..., 0, 1] / dA
    hgeo[:, HG_AI00 + 2], hgeo[:, HG_AI00 + 3] = -A[..., 1, 0] / dA, A[..., 0, 0] / dA
    gc, gf = Yc.copy(), Yf.copy()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert emud.emu_dss_generic(nh, nv, ncf, len(off) - 1, p(off), p(m32), p(hgeo), p(gc), p(gf)) == 0
    oc, of = Yc.copy(), Yf.copy()
    o.dss_state(oc, of)
    for k in range(ncf):
        assert rel(gc[:, k], oc[:, k]) < 1e-14, (k, rel(gc[:, k], oc[:, k]))
    assert rel(gf, of) < 1e-14
