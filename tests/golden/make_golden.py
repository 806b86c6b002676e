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


if __name__ == "__main__":
    Yc, Yf = run_case()
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_step_he2_ze8_f64.npz")
    np.savez_compressed(out, Yc=Yc, Yf=Yf)
    print("wrote", out, Yc.shape, Yf.shape)
