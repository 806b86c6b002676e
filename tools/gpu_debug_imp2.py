import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from climaatmos_jl_b200 import dycore, params as prm


def run(env, FT, he, ze, zmax, dzb, dt, upw):
    for k in ("B200_IMP_KERNEL", "B200_IMP_THOMAS", "B200_GENERIC_NV", "B200_IMP_MINB"):
        os.environ.pop(k, None)
    os.environ.update(env)
    sim = dycore.AtmosSimulation(FT=FT, h_elem=he, z_elem=ze, z_max=zmax, dz_bottom=dzb, dt=dt, energy_q_tot_upwinding=upw)
    rng = np.random.default_rng(0)
    Yc, Yf = sim.Y.cpu()
    Yc = Yc * (1 + 1e-3 * rng.standard_normal(Yc.shape)).astype(FT)
    Yf = (Yf + 0.1 * rng.standard_normal(Yf.shape)).astype(FT)
    U = sim.to_device(Yc, Yf)
    N = U.zeros_like()
    sim.implicit_stage(N, U, 0.4358665215 * dt)
    torch.cuda.synchronize()
    out = N.cpu()
    sim.close()
    return out

np.set_printoptions(linewidth=250, precision=1)
FT = np.float64
for upw in ("none", "vanleer_limiter"):
    case = (3, 63, 60000.0, 30.0, 90.0)
    ref = run({"B200_IMP_KERNEL": "2"}, FT, *case, upw)
    got = run({"B200_IMP_THOMAS": "1"}, FT, *case, upw)
    for name, a, b in (("rho", got[0][:, 0], ref[0][:, 0]), ("rhoe", got[0][:, 3], ref[0][:, 3]), ("u3", got[1][:, 0], ref[1][:, 0])):
        d = np.abs(a - b) / np.abs(b).max()
        print(upw, name, "per j:", d.max(axis=(0, 2, 3)), "per i:", d.max(axis=(0, 1, 3)))
        print("   per v:", d.max(axis=(0, 1, 2)))
