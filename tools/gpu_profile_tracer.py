"""Developer tool (GPU box): a short tracer-carrying run (configs[2] shape) for ncu launch lists."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from climaatmos_jl_b200 import dycore, params as prm

P = prm.DycoreParams(zd_rayleigh=40000.0, zd_viscous=40000.0)
tr = [lambda lat, lon, z: 0.5 * (1 + np.sin(np.radians(lat)) * np.cos(np.radians(lon))) * np.exp(-z / 8000.0)]
sim = dycore.AtmosSimulation(FT=np.float32, h_elem=30, z_elem=63, z_max=60000.0, dz_bottom=30.0, dt=90.0,
                             rayleigh_sponge=True, viscous_sponge=True, params=P, tracers=None if os.environ.get("MOIST") else tr,
                             apply_sem_quasimonotone_limiter=bool(os.environ.get("LIMITER")),
                             **(dict(microphysics_model="0M", initial_condition="MoistBaroclinicWave") if os.environ.get("MOIST") else {}))
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    sim.step(True)
torch.cuda.synchronize()
