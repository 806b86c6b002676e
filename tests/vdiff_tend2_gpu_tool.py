"""Last-GPU-seconds check of k_vdiff_tend2 (B200_VDIFF_KERNEL=2): T_exp parity with the oracle (explicit diffusion), bitwise identity
with k_vdiff_tend, then ms/step at he30/ze63 Float32 with explicit diffusion.  Lines are flushed as they are produced → gpurun_out/."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t0 = time.time()
os.makedirs("gpurun_out", exist_ok=True)
LOG = open("gpurun_out/vdiff_tend2.log", "a")


def say(*a):
    msg = f"[{time.time() - t0:5.1f}s] " + " ".join(str(x) for x in a)
    print(msg, flush=True)
    LOG.write(msg + "\n"); LOG.flush(); os.fsync(LOG.fileno())


import numpy as np, torch
from climaatmos_jl_b200 import dycore, params as prm
from climaatmos_jl_b200.grid import make_sphere_grid
from oracle.dycore_oracle import Oracle

rel = lambda a, b: float(np.linalg.norm((a.astype(np.float64) - b).ravel()) / np.linalg.norm(b.astype(np.float64).ravel()))
P = prm.DycoreParams(D_0_diffusion=60.0, H_diffusion=5000.0)
tr = [lambda lat, lon, z: 1e-2 * (1 + 0.5 * np.cos(z / 900.0) * np.cos(np.radians(lat))) + 0 * lon]
for FT in (np.float64, np.float32):
    res = {}
    for kern in ("2", "1"):
        os.environ["B200_VDIFF_KERNEL"] = kern
        sim = dycore.AtmosSimulation(FT=FT, h_elem=3, z_elem=10, z_max=30000.0, dz_bottom=500.0, dt=200.0, params=P, tracers=tr,
                                     vert_diff="VerticalDiffusion")
        Yt, Yl = sim.Y.zeros_like(), sim.Y.zeros_like()
        sim.remaining_tendency(Yt, Yl, sim.Y)
        torch.cuda.synchronize()
        res[kern] = (Yt.cpu()[0], Yl.cpu()[0])
        if kern == "2":
            o = Oracle(sim.grid, P, sim.numerics, np.float64)
            Yc, Yf = [a.astype(np.float64) for a in sim.Y.cpu()]
            pc = o.set_implicit_precomputed_quantities(Yc, Yf)
            tc, tf, lc = o.remaining_tendency(Yc, Yf, pc, with_lim=True)
            g = res["2"]
            say(FT.__name__, "T_exp (explicit diffusion, kernel 2) vs F64 oracle:", [f"{rel(g[0][:, k], tc[:, k]):.2e}" for k in range(4)],
                f"tracer {rel(g[0][:, 4] + g[1][:, 4], tc[:, 4] + lc[:, 4]):.2e}")
        sim.close()
    say(FT.__name__, "kernel 2 == kernel 1 bitwise:", bool(np.array_equal(res["1"][0], res["2"][0])))
os.environ["B200_VDIFF_KERNEL"] = "2"
P = prm.DycoreParams(zd_rayleigh=40000.0, zd_viscous=40000.0, D_0_diffusion=5.0, H_diffusion=800.0)
sim = dycore.AtmosSimulation(FT=np.float32, h_elem=30, z_elem=63, z_max=60000.0, dz_bottom=30.0, dt=90.0, rayleigh_sponge=True,
                             viscous_sponge=True, params=P, vert_diff="DecayWithHeightDiffusion")
for _ in range(3):
    sim.step(True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    sim.step(True)
e1.record()
torch.cuda.synchronize()
line = dict(config="dry_baroclinic_wave he30 ze63 Float32, vertical diffusion: explicit, k_vdiff_tend2", ms_per_step=e0.elapsed_time(e1) / 10,
            steps=10, warmup=3, finite=bool(torch.isfinite(sim.Y.c).all().item()))
say(json.dumps(line))
open("gpurun_out/vdiff_timing.jsonl", "a").write(json.dumps(line) + "\n")
