"""Worker for the multi-process tests (launched by torchrun / mp.spawn).

mode "gloo-halo": CPU, world_size ≥ 2, gloo — emulates the library's DSS halo protocol (send whole
element slabs of `send_elems`, receive into ghost slots, sum collocated nodes in ascending global element
order) with NumPy and checks the owned elements against the single-process oracle DSS.
mode "nccl-step": GPU, one rank per device — runs ARS343 steps through the C-ABI with the NCCL halo and
checks every rank's owned elements against a single-GPU run of the same global problem.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gloo_halo(rank, world, port):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from climaatmos_jl_b200 import grid as G, params as prm, partition
    from oracle.dycore_oracle import Oracle

    P = prm.DycoreParams()
    g = G.make_sphere_grid(FT=np.float64, h_elem=3, z_elem=5, z_max=30000.0, dz_bottom=1000.0, radius=P.planet_radius)
    rng = np.random.default_rng(7)
    Yc = rng.standard_normal((g.nelems, 4, 4, 4, g.nv))
    Yf = rng.standard_normal((g.nelems, 1, 4, 4, g.nv + 1))
    part = partition.partition_grid(g, rank, world)
    own = part.elems_ext[: part.nh]
    loc_c, loc_f = Yc[own].copy(), Yf[own].copy()
    ghost_c = np.zeros((part.nh_ghost,) + Yc.shape[1:])
    ghost_f = np.zeros((part.nh_ghost,) + Yf.shape[1:])
    reqs = []
    for k, q in enumerate(part.neighbor_ranks):
        s = part.send_elems[part.send_offset[k]:part.send_offset[k + 1]]
        sc, sf = torch.from_numpy(loc_c[s].copy()), torch.from_numpy(loc_f[s].copy())
        reqs += [dist.isend(sc, int(q), tag=0), dist.isend(sf, int(q), tag=1)]
        n = part.recv_offset[k + 1] - part.recv_offset[k]
        rc, rf = torch.empty((n,) + Yc.shape[1:], dtype=torch.float64), torch.empty((n,) + Yf.shape[1:], dtype=torch.float64)
        dist.recv(rc, int(q), tag=0)
        dist.recv(rf, int(q), tag=1)
        ghost_c[part.recv_offset[k]:part.recv_offset[k + 1]] = rc.numpy()
        ghost_f[part.recv_offset[k]:part.recv_offset[k + 1]] = rf.numpy()
    for r in reqs:
        r.wait()
    # local DSS over (local + ghost) with the partition's tables, members ordered by global element id
    ext_c, ext_f = np.concatenate([loc_c, ghost_c]), np.concatenate([loc_f, ghost_f])
    topo = G.Topology2D(part.nh + part.nh_ghost, [], part.interior_faces, part.local_vertices, part.local_vertex_offset)
    off, mem = G.dss_node_csr(topo, 4)
    A = g.dxdxi[part.elems_ext]
    Ainv = np.linalg.inv(A)
    WJ = (g.W * g.J2)[part.elems_ext]
    out_c, out_f = ext_c.copy(), ext_f.copy()
    for n in range(len(off) - 1):
        m = sorted(mem[off[n]:off[n + 1]].tolist(), key=lambda x: part.elems_ext[x[0]])
        if not any(e < part.nh for e, _, _ in m):
            continue
        wsum = sum(WJ[e, j, i] for e, i, j in m)
        for comp, arr, out in ((0, ext_c, out_c), (3, ext_c, out_c), (0, ext_f, out_f)):
            s = sum(WJ[e, j, i] / wsum * arr[e, comp, j, i] for e, i, j in m)
            for e, i, j in m:
                out[e, comp, j, i] = s
        su = sum(WJ[e, j, i] / wsum * (Ainv[e, j, i, 0, 0] * ext_c[e, 1, j, i] + Ainv[e, j, i, 1, 0] * ext_c[e, 2, j, i]) for e, i, j in m)
        sv = sum(WJ[e, j, i] / wsum * (Ainv[e, j, i, 0, 1] * ext_c[e, 1, j, i] + Ainv[e, j, i, 1, 1] * ext_c[e, 2, j, i]) for e, i, j in m)
        for e, i, j in m:
            out_c[e, 1, j, i] = A[e, j, i, 0, 0] * su + A[e, j, i, 1, 0] * sv
            out_c[e, 2, j, i] = A[e, j, i, 0, 1] * su + A[e, j, i, 1, 1] * sv
    o = Oracle(g, P, prm.DycoreNumerics(), np.float64)
    rc, rf = Yc.copy(), Yf.copy()
    o.dss_state(rc, rf)
    ec = np.abs(out_c[: part.nh] - rc[own]).max()
    ef = np.abs(out_f[: part.nh] - rf[own]).max()
    t = torch.tensor([ec, ef])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    dist.destroy_process_group()
    assert float(t.max()) < 1e-12, f"halo DSS mismatch {t}"


def nccl_step(tracers=False):
    import torch
    from climaatmos_jl_b200 import dycore, params as prm
    from climaatmos_jl_b200.parallel import DistributedComms

    comms = DistributedComms()
    P = prm.DycoreParams(zd_rayleigh=20000.0, zd_viscous=20000.0)
    kw = dict(FT=np.float32, h_elem=4, z_elem=15, z_max=30000.0, dz_bottom=300.0, dt=300.0, rayleigh_sponge=True, viscous_sponge=True, params=P)
    if tracers:  # a step-function tracer with the SEM quasi-monotone limiter: neighbour bounds travel through the peer-memory halo
        kw.update(tracers=[lambda lat, lon, z: (np.abs(lat) < 30.0) * (z < 12000.0) * 1.0 + 0 * lon], apply_sem_quasimonotone_limiter=True)
    sim = dycore.AtmosSimulation(comms=comms, **kw)
    for _ in range(3):
        sim.step(True)
    torch.cuda.synchronize()
    gc, gf = sim.Y.cpu()
    ref = dycore.AtmosSimulation(**kw)  # the same global problem on this rank's GPU alone
    for _ in range(3):
        ref.step(True)
    rc, rf = ref.Y.cpu()
    own = sim.part.elems_ext[: sim.part.nh]
    ok = np.array_equal(gc, rc[own]) and np.array_equal(gf, rf[own])
    err = max(np.abs(gc - rc[own]).max(), np.abs(gf - rf[own]).max())
    print(f"rank {comms.rank}/{comms.nranks}: owned {len(own)} elements, bitwise_equal={ok}, max abs diff {err:.3e}", flush=True)
    t = torch.tensor([0.0 if ok else 1.0], device="cuda")
    torch.distributed.all_reduce(t)
    sim.close(); ref.close()
    comms.finalize()
    if float(t.item()) != 0:
        sys.exit(3)


def halo_timeout():
    """Two ranks; after two common steps rank 1 stops stepping while rank 0 steps on.  Rank 0's halo wait must time out (the spin limit is
    shortened through B200_P2P_SPIN_LIMIT), the process and its CUDA context must survive, and the NEXT call on rank 0 must return an error
    code with a message naming the neighbour — not a trap, not a hang."""
    import torch
    from climaatmos_jl_b200 import dycore, params as prm
    from climaatmos_jl_b200.parallel import DistributedComms

    os.environ["B200_P2P_SPIN_LIMIT"] = "400000000"  # ≈ 0.2 s of SM clocks
    comms = DistributedComms()
    P = prm.DycoreParams(zd_rayleigh=20000.0, zd_viscous=20000.0)
    sim = dycore.AtmosSimulation(comms=comms, FT=np.float32, h_elem=4, z_elem=15, z_max=30000.0, dz_bottom=300.0, dt=300.0, params=P)
    assert getattr(sim, "peer_halo", False), "needs the peer-memory halo"
    for _ in range(2):
        sim.step(True)
    torch.cuda.synchronize()
    comms.barrier()
    msg = ""
    if comms.rank == 0:
        sim.step(True)  # the neighbour never arrives: every halo wait of this step times out, the kernels run on
        torch.cuda.synchronize()  # the context is alive (a trap would raise here)
        try:
            sim.step(True)
        except RuntimeError as e:
            msg = str(e)
        ok = "halo timed out" in msg and "neighbour rank 1" in msg
        x = torch.ones(4, device="cuda").sum().item()  # CUDA still works in this process
        print(f"rank 0: timeout_reported={ok and x == 4.0} message={msg[:160]!r}", flush=True)
    comms.barrier()
    comms.finalize()
    os._exit(0)  # the contexts hold peer mappings of a rank that is out of step: skip the orderly teardown


if __name__ == "__main__":
    if sys.argv[1] == "halo-timeout":
        halo_timeout()
    elif sys.argv[1] == "nccl-step":
        nccl_step()
    elif sys.argv[1] == "nccl-step-tracer-limiter":
        nccl_step(tracers=True)
