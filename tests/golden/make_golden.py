"""Generates tests/golden/oracle_step_he2_ze8_f64.npz: one ARS343 step of the NumPy oracle from the
dry baroclinic-wave initial state (he2, ze8, z_max 30 km, dt 600 s, Float64, sponges on).
The reference itself cannot run here (no Julia; ClimaCore et al. un-vendored), so this fixture pins
the ORACLE, not the reference: "parity unpinned" (DESIGN.md).  Run: python -m tests.golden.make_golden"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def run_case():
    from climaatmos_jl_b200 import grid as G, params as prm, setups
    from oracle.dycore_oracle import Oracle

    P = prm.DycoreParams(zd_rayleigh=20000.0, zd_viscous=20000.0)
    g = G.make_sphere_grid(FT=np.float64, h_elem=2, z_elem=8, z_max=30000.0, dz_bottom=500.0, radius=P.planet_radius)
    N = prm.DycoreNumerics(dt=600.0, rayleigh_sponge=True, viscous_sponge=True)
    o = Oracle(g, P, N, np.float64)
    Yc, Yf = setups.dry_baroclinic_wave(g, P)
    return o.step(Yc, Yf)


def run_case_vdiff():
    """Second fixture (oracle_step_vdiff_he2_ze8_f64.npz): the same grid with one tracer (negative cells), implicit VerticalDiffusion with
    two iterations of the approximate arrowhead solve and the vertical mass-borrowing limiter — pins the oracle's SURVEY §8f n1/n2 additions."""
    from climaatmos_jl_b200 import grid as G, params as prm, setups
    from oracle.dycore_oracle import Oracle

    P = prm.DycoreParams(zd_rayleigh=20000.0, zd_viscous=20000.0, C_E=0.0044)
    g = G.make_sphere_grid(FT=np.float64, h_elem=2, z_elem=8, z_max=30000.0, dz_bottom=500.0, radius=P.planet_radius)
    N = prm.DycoreNumerics(dt=300.0, rayleigh_sponge=True, viscous_sponge=True, vert_diff="VerticalDiffusion", implicit_diffusion=True,
                           approximate_linear_solve_iters=2, tracer_nonnegativity_method="vertical_water_borrowing")
    o = Oracle(g, P, N, np.float64)
    Yc, Yf = setups.dry_baroclinic_wave(g, P)
    chi = 4e-4 + 1e-3 * np.cos(o.c.z / 700.0) * np.cos(3 * np.radians(g.lat))[..., None]
    Yc = np.concatenate([Yc, (Yc[:, 0] * chi)[:, None]], axis=1)
    return o.step(Yc, Yf)


if __name__ == "__main__":
    Yc, Yf = run_case_vdiff()
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_step_vdiff_he2_ze8_f64.npz")
    np.savez_compressed(out, Yc=Yc, Yf=Yf)
    print("wrote", out, Yc.shape, Yf.shape)
    Yc, Yf = run_case()
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_step_he2_ze8_f64.npz")
    np.savez_compressed(out, Yc=Yc, Yf=Yf)
    print("wrote", out, Yc.shape, Yf.shape)
