"""Developer tool (GPU box): CUDA-event timing of each hook / phase at the north-star size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from climaatmos_jl_b200 import dycore, params as prm

he = int(sys.argv[1]) if len(sys.argv) > 1 else 30
P = prm.DycoreParams(zd_rayleigh=40000.0, zd_viscous=40000.0)
MOIST = bool(os.environ.get("MOIST"))  # MOIST=1: the 0M-moist configuration (BASELINE.json configs[2])
sim = dycore.AtmosSimulation(FT=np.float32, h_elem=he, z_elem=63, z_max=60000.0, dz_bottom=30.0, dt=90.0 * 30 / he,
                             rayleigh_sponge=True, viscous_sponge=True, params=P,
                             **(dict(microphysics_model="0M", initial_condition="MoistBaroclinicWave", q_0=float(os.environ.get("Q0", "0.018"))) if MOIST else {}))
for _ in range(3):
    sim.step(True)
Y = sim.Y
Yt = Y.zeros_like(); R = Y.zeros_like(); dY = Y.zeros_like()


def timeit(name, fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:28s} {e0.elapsed_time(e1) / reps * 1e3:9.1f} us", flush=True)


timeit("t_exp phase0 (pre-DSS)", lambda: sim.remaining_tendency_phase(0, Yt, Y))
timeit("t_exp phase1 (DSS H)", lambda: sim.remaining_tendency_phase(1, Yt, Y))
timeit("t_exp phase2 (hyper apply)", lambda: sim.remaining_tendency_phase(2, Yt, Y))
timeit("t_exp total", lambda: sim.remaining_tendency(Yt, None, Y))
N = Y.zeros_like()
timeit("implicit stage (fused)", lambda: sim.implicit_stage(N, Y, 39.0))
timeit("dss state", lambda: sim.dss(Y))
QUICK = bool(os.environ.get("QUICK"))
if QUICK:
    timeit("step fused", lambda: sim.step(True), reps=10)
    print("finite:", bool(torch.isfinite(sim.Y.c).all()))
    sys.exit(0)
timeit("cache_imp", lambda: sim.set_implicit_precomputed_quantities(Y))
timeit("t_imp", lambda: sim.implicit_tendency(Yt, Y))
timeit("wfact", lambda: sim.update_jacobian(Y, 39.0))
timeit("ldiv", lambda: sim.ldiv(dY, R))
timeit("t_post_imp", lambda: sim.correct_implicit_advection_tendency(Yt, Y))
timeit("step fused", lambda: sim.step(True), reps=10)
timeit("step hooks", lambda: sim.step(False), reps=5)
print("finite:", bool(torch.isfinite(sim.Y.c).all()))
