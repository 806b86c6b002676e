"""Multi-GPU soak: K fused steps of a mid-size problem on all ranks through the peer-memory halo (CUDA-graph replay),
against the SAME problem stepped on each rank's GPU alone; every rank's owned elements must agree BITWISE at several checkpoints.
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/gpu_soak_multi.py
[--steps 500] [--he 16] [--ze 31] [--moist].  Rank 0 prints one JSON line."""
import argparse
import json
import os
import sys
import time
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from climaatmos_jl_b200 import dycore, params as prm
from climaatmos_jl_b200.parallel import DistributedComms

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=500)
ap.add_argument("--he", type=int, default=16)
ap.add_argument("--ze", type=int, default=31)
ap.add_argument("--every", type=int, default=100)
ap.add_argument("--moist", action="store_true")
args = ap.parse_args()

comms = DistributedComms()
torch.cuda.set_device(comms.local_rank)
P = prm.DycoreParams(zd_rayleigh=30000.0, zd_viscous=30000.0)
kw = dict(FT=np.float32, h_elem=args.he, z_elem=args.ze, z_max=45000.0, dz_bottom=300.0, dt=90.0 * 30 / args.he / 2, rayleigh_sponge=True,
          viscous_sponge=True, params=P)
if args.moist:
    kw.update(microphysics_model="0M", initial_condition="MoistBaroclinicWave")
t0 = time.time()
sim = dycore.AtmosSimulation(comms=comms, **kw)
ref = dycore.AtmosSimulation(**kw)
own = sim.part.elems_ext[: sim.part.nh]
checks, ok_all = [], True
for k in range(1, args.steps + 1):
    sim.step(True)
    ref.step(True)
    if k % args.every == 0 or k == args.steps:
        torch.cuda.synchronize()
        gc, gf = sim.Y.cpu()
        rc, rf = ref.Y.cpu()
        ok = bool(np.array_equal(gc, rc[own]) and np.array_equal(gf, rf[own]) and np.isfinite(gc).all() and np.isfinite(gf).all())
        ok = comms.all_true(ok)
        ok_all = ok_all and ok
        checks.append({"step": k, "bitwise_equal_all_ranks": ok, "rank0_crc32": int(zlib.crc32(gc.tobytes()) ^ zlib.crc32(gf.tobytes()))})
halo = "nvlink-peer-memory" if getattr(sim, "peer_halo", False) else "nccl-send-recv"
if comms.rank == 0:
    print(json.dumps({"soak": f"{'moist 0M' if args.moist else 'dry'} baroclinic wave he{args.he} ze{args.ze} Float32, {args.steps} fused steps (graph replay)",
                      "n_gpus": comms.nranks, "halo": halo, "elements_per_rank": int(sim.part.nh), "ghost_elements_rank0": int(sim.part.nh_ghost),
                      "bitwise_equal_to_single_gpu_at_every_checkpoint": ok_all, "checkpoints": checks, "wall_s": round(time.time() - t0, 1)}))
sim.close()
ref.close()
