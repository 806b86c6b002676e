"""ms/step of the dry baroclinic wave he30/ze63 Float32 with vertical diffusion (DecayWithHeightDiffusion): explicit (fused, graph-replayed
step + one k_vdiff_tend per T_exp) and implicit (implicit stages through the hook sequence with k_ldiv_diff, 2 iterations), next to the
no-diffusion step.  CUDA events, 3 warm-up steps, 10 timed steps; one JSON line per configuration → gpurun_out/vdiff_timing.jsonl."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from climaatmos_jl_b200 import dycore, params as prm
from climaatmos_jl_b200.grid import make_sphere_grid

os.makedirs("gpurun_out", exist_ok=True)
out = open("gpurun_out/vdiff_timing.jsonl", "a")
HE = int(os.environ.get("HE", "30"))
P = prm.DycoreParams(zd_rayleigh=40000.0, zd_viscous=40000.0, D_0_diffusion=5.0, H_diffusion=800.0)  # toml/rcemipii_box.toml:61-65
t0 = time.time()
grid = make_sphere_grid(FT=np.float32, h_elem=HE, z_elem=63, z_max=60000.0, dz_bottom=30.0, radius=P.planet_radius, deep_atmosphere=True)
print(f"grid {time.time() - t0:.1f}s", flush=True)
for name, kw in (("none", {}), ("explicit", dict(vert_diff="DecayWithHeightDiffusion")),
                 ("implicit", dict(vert_diff="DecayWithHeightDiffusion", implicit_diffusion=True, approximate_linear_solve_iters=2))):
    sim = dycore.AtmosSimulation(FT=np.float32, h_elem=HE, z_elem=63, z_max=60000.0, dz_bottom=30.0, dt=90.0 * 30 / HE, rayleigh_sponge=True,
                                 viscous_sponge=True, params=P, grid=grid, **kw)
    for _ in range(3):
        sim.step(True)
    torch.cuda.synchronize()
    l0 = sim.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 10
    e0.record()
    for _ in range(K):
        sim.step(True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    line = dict(config=f"dry_baroclinic_wave he{HE} ze63 Float32, vertical diffusion: {name}", ms_per_step=ms, steps=K, warmup=3,
                launches_per_step=(sim.launch_count() - l0) / K, finite=bool(torch.isfinite(sim.Y.c).all().item()))
    print(json.dumps(line), flush=True)
    out.write(json.dumps(line) + "\n")
    out.flush()
    sim.close()
    del sim
    torch.cuda.empty_cache()
