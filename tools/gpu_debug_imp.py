"""Developer tool (GPU box): compare the fused implicit-stage kernel variants on the same input."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from climaatmos_jl_b200 import dycore, params as prm


def run(env, FT, he, ze, zmax, dzb, dt):
    for k in ("B200_IMP_KERNEL", "B200_IMP_THOMAS", "B200_GENERIC_NV", "B200_IMP_MINB"):
        os.environ.pop(k, None)
    os.environ.update(env)
    sim = dycore.AtmosSimulation(FT=FT, h_elem=he, z_elem=ze, z_max=zmax, dz_bottom=dzb, dt=dt)
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


for FT in (np.float64, np.float32):
    for case in ((4, 10, 30000.0, 500.0, 400.0), (3, 63, 60000.0, 30.0, 90.0)):
        ref = run({"B200_IMP_KERNEL": "2"}, FT, *case)
        for env in ({}, {"B200_IMP_THOMAS": "1"}, {"B200_GENERIC_NV": "1"}, {"B200_IMP_MINB": "3"}):
            got = run(env, FT, *case)
            msg = []
            for name, a, b in (("rho", got[0][:, 0], ref[0][:, 0]), ("u1", got[0][:, 1], ref[0][:, 1]), ("rhoe", got[0][:, 3], ref[0][:, 3]),
                               ("u3", got[1][:, 0], ref[1][:, 0])):
                d = np.abs(a.astype(np.float64) - b.astype(np.float64))
                bad = ~np.isfinite(a)
                k = np.unravel_index(np.argmax(np.where(bad, np.inf, d)), d.shape)
                msg.append(f"{name}: rel {np.nanmax(d) / np.abs(b).max():.2e} nan {int(bad.sum())} at (h,j,i,v)={tuple(int(x) for x in k)}")
            print(FT.__name__, case[:2], env, " | ".join(msg), flush=True)
