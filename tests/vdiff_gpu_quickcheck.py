"""One-shot GPU check of the vertical-diffusion kernels against the oracle, most important comparisons first, one flushed line
each (so that a call cut short by the GPU budget still leaves evidence).  Writes gpurun_out/vdiff_quickcheck.log."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t0 = time.time()
os.makedirs("gpurun_out", exist_ok=True)
LOG = open("gpurun_out/vdiff_quickcheck.log", "a")


def say(*a):
    msg = f"[{time.time() - t0:6.1f}s] " + " ".join(str(x) for x in a)
    print(msg, flush=True)
    LOG.write(msg + "\n")
    LOG.flush()
    os.fsync(LOG.fileno())


say("start")
import numpy as np
import torch
say("torch imported", torch.cuda.get_device_name(0))
from climaatmos_jl_b200 import dycore, params as prm
from oracle.dycore_oracle import Oracle


def rel(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    n = np.linalg.norm(b.ravel())
    return np.linalg.norm((a - b).ravel()) / n if n > 0 else np.linalg.norm(a.ravel())


def make(FT, vd, implicit, iters=2):
    P = prm.DycoreParams(D_0_diffusion=60.0, H_diffusion=5000.0)
    tr = [lambda lat, lon, z: 1e-2 * (1 + 0.5 * np.cos(z / 900.0) * np.cos(np.radians(lat))) + 0 * lon]
    sim = dycore.AtmosSimulation(FT=FT, h_elem=3, z_elem=10, z_max=30000.0, dz_bottom=500.0, dt=200.0, params=P, tracers=tr,
                                 vert_diff=vd, implicit_diffusion=implicit, approximate_linear_solve_iters=iters)
    return sim, Oracle(sim.grid, P, sim.numerics, FT)


for FT in (np.float64, np.float32):
    for vd in ("DecayWithHeightDiffusion", "VerticalDiffusion"):
        sim, o = make(FT, vd, True)
        Yc0, Yf0 = sim.Y.cpu()
        rng = np.random.default_rng(1234)
        Yc = (Yc0.astype(np.float64) * (1 + 1e-3 * rng.standard_normal(Yc0.shape))).astype(FT)
        Yf = (0.5 * sim.grid.dz_f * rng.standard_normal(Yf0.shape)).astype(FT)
        Yf[..., 0] = 0
        Yf[..., -1] = 0
        Y = sim.to_device(Yc, Yf)
        pc = o.set_implicit_precomputed_quantities(Yc.copy(), Yf.copy())
        Yt = Y.zeros_like()
        sim.implicit_tendency(Yt, Y)
        torch.cuda.synchronize()
        tc, tf = o.implicit_tendency(Yc, Yf, pc)
        gc, gf = Yt.cpu()
        say(FT.__name__, vd, "t_imp rel-L2", [f"{rel(gc[:, k], tc[:, k]):.2e}" for k in range(5)], f"u3 {rel(gf, tf):.2e}")
        dtg = sim.dt * 0.4358665215
        sim.update_jacobian(Y, dtg)
        Jm = o.update_jacobian(Yc, Yf, pc, dtg)
        Rc = (rng.standard_normal(Yc.shape) * np.abs(Yc).mean(axis=(0, 2, 3, 4), keepdims=True) * 1e-3).astype(FT)
        Rf = rng.standard_normal(Yf.shape).astype(FT)
        R = sim.to_device(Rc, Rf)
        dY = R.zeros_like()
        sim.ldiv(dY, R)
        torch.cuda.synchronize()
        dc, df = o.ldiv(Jm, Rc, Rf)
        gc, gf = dY.cpu()
        say(FT.__name__, vd, "ldiv  rel-L2", [f"{rel(gc[:, k], dc[:, k]):.2e}" for k in range(5)], f"u3 {rel(gf, df):.2e}")
        sim.close()
for FT in (np.float64, np.float32):
    for implicit in (True, False):
        sim, o = make(FT, "DecayWithHeightDiffusion", implicit)
        Yc0, Yf0 = sim.Y.cpu()
        sim.step(fused=True)
        torch.cuda.synchronize()
        gc, gf = sim.Y.cpu()
        o64 = Oracle(sim.grid, sim.params, sim.numerics, np.float64)
        oc, of = o64.step(Yc0.astype(np.float64), Yf0.astype(np.float64))
        say(FT.__name__, "step implicit_diffusion =", implicit, [f"{rel(gc[:, k], oc[:, k]):.2e}" for k in range(5)], f"u3 {rel(gf, of):.2e}",
            "launches", sim.launch_count())
        sim.close()
say("done")
