"""Developer tool (GPU box): parity at the BASELINE.json north-star size — Float32 CUDA step(s) against the Float64 NumPy oracle on
the dry baroclinic wave he30/ze63 (the oracle, element-chunked over the host threads, needs a few seconds per step) — and a
300-step soak of the fused, graph-replayed stepper (finite state, mass drift).  Writes gpurun_out/fullsize_parity.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
import numpy as np, torch
from concurrent.futures import ThreadPoolExecutor
from climaatmos_jl_b200 import dycore, params as prm
from oracle.dycore_oracle import Oracle

P = prm.DycoreParams(zd_rayleigh=40000.0, zd_viscous=40000.0)
sim = dycore.AtmosSimulation(FT=np.float32, h_elem=30, z_elem=63, z_max=60000.0, dz_bottom=30.0, dt=90.0,
                             rayleigh_sponge=True, viscous_sponge=True, params=P)
Yc0, Yf0 = sim.Y.cpu()
o = Oracle(sim.grid, P, sim.numerics, np.float64)
oc, of = Yc0.astype(np.float64), Yf0.astype(np.float64)
rel = lambda a, b: float(np.linalg.norm((a.astype(np.float64) - b).ravel()) / np.linalg.norm(b.ravel()))
out = {"config": "dry_baroclinic_wave he30 ze63 dt=90s, Float32 CUDA (fused, graph) vs Float64 oracle", "steps": []}
cores = min(16, os.cpu_count() or 1)
with ThreadPoolExecutor(cores) as pool:
    for k in range(2):
        t0 = time.time()
        oc, of = o.step(oc, of, pool=pool, nchunks=cores)
        t_cpu = time.time() - t0
        sim.step(True)
        torch.cuda.synchronize()
        gc, gf = sim.Y.cpu()
        row = {"step": k + 1, "oracle_seconds": round(t_cpu, 2), "rho": rel(gc[:, 0], oc[:, 0]), "u1": rel(gc[:, 1], oc[:, 1]),
               "u2": rel(gc[:, 2], oc[:, 2]), "rhoe": rel(gc[:, 3], oc[:, 3]), "u3": rel(gf[:, 0], of[:, 0])}
        out["steps"].append(row)
        print(row, flush=True)
g = sim.grid
vol = (g.W * g.J2)[..., None] * (((g.radius + g.z_c) / g.radius) ** 2 * g.dz_c)
m0 = float((vol * gc[:, 0].astype(np.float64)).sum())
for _ in range(300):
    sim.step(True)
torch.cuda.synchronize()
gc, gf = sim.Y.cpu()
out["soak"] = {"steps": 300, "finite": bool(np.isfinite(gc).all() and np.isfinite(gf).all()),
               "mass_drift": abs(float((vol * gc[:, 0].astype(np.float64)).sum()) - m0) / m0,
               "max_abs_w": float(np.abs(gf[:, 0] / g.dz_f).max())}
print(out["soak"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/fullsize_parity.json", "w"), indent=1)
sim.close()
