"""Developer tool (GPU box): print hook-by-hook and full-step errors of the CUDA path vs the NumPy oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from climaatmos_jl_b200 import dycore, params as prm
from climaatmos_jl_b200.dycore import FieldVector
from oracle.dycore_oracle import Oracle


def rel(a, b):
    d = np.linalg.norm((a.astype(np.float64) - b.astype(np.float64)).ravel())
    n = np.linalg.norm(b.astype(np.float64).ravel())
    return d / n if n > 0 else d


def report(tag, gc, gf, oc, of):
    names = ["rho", "u1", "u2", "rhoe"]
    s = " ".join(f"{names[k]}={rel(gc[:, k], oc[:, k]):.2e}" for k in range(4))
    print(f"{tag:28s} {s} u3={rel(gf[:, 0], of[:, 0]):.2e}", flush=True)


def run(FT, h_elem, z_elem, z_max, dzb, dt, sponge, oFT=None):
    P = prm.DycoreParams(zd_rayleigh=0.66 * z_max, zd_viscous=0.66 * z_max)
    sim = dycore.AtmosSimulation(FT=FT, h_elem=h_elem, z_elem=z_elem, z_max=z_max, dz_bottom=dzb, dt=dt,
                                 rayleigh_sponge=sponge, viscous_sponge=sponge, params=P)
    oFT = oFT or FT
    o = Oracle(sim.grid, P, sim.numerics, oFT)
    print(f"=== FT={np.dtype(FT).name} oracle={np.dtype(oFT).name} he={h_elem} ze={z_elem} sponge={sponge}")
    Yc0, Yf0 = sim.Y.cpu()
    # perturb the state so every term is exercised (w != 0, non-DSSed)
    rng = np.random.default_rng(1234)
    Yc = Yc0.astype(np.float64) * (1 + 1e-3 * rng.standard_normal(Yc0.shape))
    Yf = 0.5 * sim.grid.dz_f[None, None, None, None, :] * rng.standard_normal(Yf0.shape)
    Yc, Yf = Yc.astype(FT), Yf.astype(FT)
    Y = sim.to_device(Yc, Yf)
    oc, of = Yc.astype(oFT), Yf.astype(oFT)
    # dss
    sim.dss(Y); o.dss_state(oc, of)
    gc, gf = Y.cpu(); report("dss", gc, gf, oc, of)
    # cache_imp
    pre = {k: torch.zeros_like(Y.c[:, 0:1]) for k in ("K_c", "T_c", "p_c", "h_tot_c")}
    pre["u3_f"] = torch.zeros_like(Y.f); pre["u_c"] = torch.zeros_like(Y.c[:, 0:3])
    sim.set_implicit_precomputed_quantities(Y, precomputed=pre)
    pc = o.set_implicit_precomputed_quantities(oc, of)
    gc, gf = Y.cpu()
    print("cache_imp", " ".join(f"{k}={rel(pre[k].cpu().numpy()[:, 0], pc[kk]):.2e}" for k, kk in
                               (("K_c", "K"), ("T_c", "T"), ("p_c", "p"), ("h_tot_c", "h_tot"), ("u3_f", "fu3"))),
          f"u3c={rel(pre['u_c'].cpu().numpy()[:, 2], pc['u3c']):.2e}", f"Yf={rel(gf, of):.2e}")
    # t_imp
    Yt = Y.zeros_like(); sim.implicit_tendency(Yt, Y)
    tc, tf = o.implicit_tendency(oc, of, pc)
    report("t_imp", *Yt.cpu(), tc, tf)
    # wfact + ldiv
    dtg = dt * 0.4358665215
    sim.update_jacobian(Y, dtg)
    Jm = o.update_jacobian(oc, of, pc, dtg)
    Rc = (rng.standard_normal(Yc.shape) * np.abs(Yc).mean(axis=(0, 2, 3, 4), keepdims=True) * 1e-3).astype(FT)
    Rf = (rng.standard_normal(Yf.shape) * 1.0).astype(FT)
    R = sim.to_device(Rc, Rf); dY = R.zeros_like()
    sim.ldiv(dY, R)
    dc, df = o.ldiv(Jm, Rc.astype(oFT), Rf.astype(oFT))
    report("ldiv", *dY.cpu(), dc, df)
    # t_post_imp
    sim.correct_implicit_advection_tendency(Yt, Y)
    tc, tf = o.correct_implicit_advection_tendency(oc, of, pc)
    g = Yt.cpu(); print(f"{'t_post_imp':28s} rhoe={rel(g[0][:, 3], tc[:, 3]):.2e} others={np.abs(g[0][:, :3]).max():.1e},{np.abs(g[1]).max():.1e}")
    # t_exp
    sim.remaining_tendency(Yt, None, Y)
    tc, tf = o.remaining_tendency(oc, of, pc)
    report("t_exp", *Yt.cpu(), tc, tf)
    # full step from the IC
    for fused in (0, 1):
        sim.Y = sim.to_device(Yc0, Yf0); sim.t = 0
        sim.step(fused=bool(fused)); torch.cuda.synchronize()
        if fused == 0:
            t0 = time.time(); s1c, s1f = o.step(Yc0.astype(oFT), Yf0.astype(oFT)); t_or = time.time() - t0
        report(f"step fused={fused}", *sim.Y.cpu(), s1c, s1f)
    print(f"oracle step took {t_or:.2f}s; relchange rho {rel(s1c[:,0], Yc0[:,0]):.2e}")
    # 10 steps drift
    sim.Y = sim.to_device(Yc0, Yf0)
    oc, of = Yc0.astype(oFT), Yf0.astype(oFT)
    for k in range(10):
        sim.step(fused=True); oc, of = o.step(oc, of)
    report("10 steps fused", *sim.Y.cpu(), oc, of)
    print("launches", sim.launch_count())
    sim.close()


if __name__ == "__main__":
    run(np.float64, 4, 10, 30000.0, 500.0, 400.0, False)
    run(np.float64, 3, 63, 60000.0, 30.0, 120.0, True)
    run(np.float32, 4, 10, 30000.0, 500.0, 400.0, False)
    run(np.float32, 3, 63, 60000.0, 30.0, 120.0, True)
    run(np.float32, 3, 63, 60000.0, 30.0, 120.0, True, oFT=np.float64)
