"""Developer tool (GPU box): a short run for `ncu -k regex:…` captures of individual kernels at he30/ze63.
Environment: MOIST=1 (0M-moist configuration), VDIFF=explicit|implicit (vertical diffusion)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from climaatmos_jl_b200 import dycore, params as prm

P = prm.DycoreParams(zd_rayleigh=40000.0, zd_viscous=40000.0, D_0_diffusion=5.0, H_diffusion=800.0)
kw = {}
if os.environ.get("MOIST"):
    kw.update(microphysics_model="0M", initial_condition="MoistBaroclinicWave")
if os.environ.get("VDIFF"):
    kw.update(vert_diff="DecayWithHeightDiffusion", implicit_diffusion=os.environ["VDIFF"] == "implicit", approximate_linear_solve_iters=2)
sim = dycore.AtmosSimulation(FT=np.float32, h_elem=30, z_elem=63, z_max=60000.0, dz_bottom=30.0, dt=90.0,
                             rayleigh_sponge=True, viscous_sponge=True, params=P, **kw)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    sim.step(True)
torch.cuda.synchronize()
