"""Developer tool (torchrun, N GPUs): time the DSS calls (with NCCL halo) and a step at the weak-scaling size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from climaatmos_jl_b200 import dycore, params as prm
from climaatmos_jl_b200.parallel import DistributedComms
import bench

comms = DistributedComms()
w = bench.workload(comms.nranks)
P = prm.DycoreParams(zd_rayleigh=40000.0, zd_viscous=40000.0)
sim = dycore.AtmosSimulation(FT=np.float32, h_elem=w["h_elem"], z_elem=63, z_max=60000.0, dz_bottom=30.0, dt=w["dt"],
                             rayleigh_sponge=True, viscous_sponge=True, params=P, comms=comms if comms.nranks > 1 else None)
for _ in range(3):
    sim.step(True)
Y = sim.Y.clone(); Yt = Y.zeros_like()

def timeit(name, fn, reps=30):
    for _ in range(3):
        fn()
    comms.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = comms.max_over_ranks(e0.elapsed_time(e1) / reps * 1e3)
    if comms.rank == 0:
        print(f"{name:28s} {t:9.1f} us", flush=True)

if comms.rank == 0 and sim.part is not None:
    p = sim.part
    print("nh", p.nh, "ghost", p.nh_ghost, "send", len(p.send_elems), "neighbors", list(p.neighbor_ranks))
timeit("dss state", lambda: sim.dss(Y))
timeit("dss H (phase1)", lambda: sim.remaining_tendency_phase(1, Yt, Y))
timeit("t_exp total", lambda: sim.remaining_tendency(Yt, None, Y))
timeit("step", lambda: sim.step(True), reps=10)
sim.close(); comms.finalize()
