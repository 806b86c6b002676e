#!/usr/bin/env python
"""bench.py — SYPD and ms/step of the dry baroclinic-wave dycore step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N = 1; N > 1 under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one ARS343 IMEX step of the whole dycore (4 × T_exp_T_lim!, 3 × implicit stage, 7 state DSS
+ 4 hyperdiffusion DSS, stage increments).  N = 1 runs the configuration the metric is quoted on
(dry baroclinic wave, h_elem = 30, z_elem = 63, Float32, dt = 90 s); N > 1 is the weak-scaling series
"he30 per GPU": h_elem = 30/42/60/85 for N = 1/2/4/8 with dt ∝ 1/h_elem (90/64/45/32 s).

`value` (both arms) is the **he30-equivalent SYPD**: the simulated-years-per-day the job would deliver if every GPU's share of
elements were a he30 sphere stepped with dt = 90 s,
    value = SYPD_raw · (dt_1 / dt_N) · (elements_N / elements_1) = (90 s / year) / (ms_per_step) · elements_N / 5400,
so that value_N / (N · value_1) IS the weak-scaling efficiency (ms_1/ms_N)·(elements_N/(N·elements_1)) of SURVEY.md §8d.4, and
value_1 is the plain SYPD of the north-star configuration.  The raw SYPD of the h_elem actually run is reported beside it
(`sypd_raw`), as is ms/step.  With `--config strong` (he60 on every N) value is the raw SYPD and `scaling` is "strong".

The K timed steps are repeated as blocks until ≥ 0.6 s have been timed; `ms_per_step` is the median block (all blocks listed).
Prints ONE JSON line on rank 0.  `--impl reference` times the CPU restatement of the reference (oracle/, NumPy "port": the Julia
reference cannot run here or on the GPU box — no julia, ClimaCore un-vendored; profiles/r2_gpu_box_probe.txt) on the SAME grid and
time step with all host cores, for as many of the requested steps as fit its time budget (real counts are printed).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

WEAK_H = {1: 30, 2: 42, 4: 60, 8: 85}
ZD = 40000.0  # zd_rayleigh = zd_viscous as in toml/longrun_held_suarez.toml


def workload(n_gpus, config="weak"):
    """BASELINE.json configs.  "weak" (default) is the north-star series; the others are extra measurements (--config):
    "strong": dry BW he60/ze63 on every N (configs[4]); "he16": dry BW he16/ze63 (configs[1]); "moist": he30/ze63 0M-moist baroclinic
    wave with the thermodynamically active ρq_tot (configs[2], dycore + tracer advection only); "tracer": the same grid dry with one
    passive tracer; "hs": Held–Suarez he6/ze10 (configs[0])."""
    if config == "strong":
        return dict(h_elem=60, z_elem=63, z_max=60000.0, dz_bottom=30.0, dt=45.0, scaling="strong", name="dry_baroclinic_wave")
    if config == "he16":
        return dict(h_elem=16, z_elem=63, z_max=60000.0, dz_bottom=30.0, dt=120.0, scaling="weak", name="dry_baroclinic_wave")
    if config == "hs":
        return dict(h_elem=6, z_elem=10, z_max=55000.0, dz_bottom=500.0, dt=400.0, scaling="weak", name="held_suarez", sponge=False, rad="held_suarez", ic="DecayingProfile")
    h = WEAK_H.get(n_gpus, int(round(30 * np.sqrt(n_gpus))))
    dt = float(round(90.0 * 30 / h))
    w = dict(h_elem=h, z_elem=63, z_max=60000.0, dz_bottom=30.0, dt=dt, scaling="weak", name="dry_baroclinic_wave", series="he30_per_gpu")
    if config == "tracer":
        w.update(tracers=1, name="dry_baroclinic_wave + 1 passive tracer")
    if config == "moist":  # configs[2] as the reference runs it: EquilibriumMicrophysics0M, ρq_tot thermodynamically active
        w.update(moist=True, ic="MoistBaroclinicWave", name="moist_baroclinic_wave 0M (active rho*q_tot: dycore + tracer advection only)")
    if config in ("vdiff", "vdiff_implicit"):
        w.update(vert_diff="DecayWithHeightDiffusion", implicit_diffusion=(config == "vdiff_implicit"),
                 name="dry_baroclinic_wave + vertical diffusion (" + ("implicit, 2 solver iterations" if config == "vdiff_implicit" else "explicit") + ")")
    return w


def model_bytes_per_step(ncols, nv, s=4, k=0):
    """SURVEY.md §8d byte model: 54.5 S + 14 H."""
    c = ncols * nv * s
    f = ncols * (nv + 1) * s
    S = (4 + k) * c + f
    H = (4 + k) * c
    return 54.5 * S + 14 * H


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 7:
                continue
            try:
                mx = float(p[1])
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(p[0]))
                    for n, v in zip(names, p[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(n)
            except ValueError:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_threads(args):
    """Host threads for the CPU arm: every core this process may run on (--cpu-threads caps it; default: no cap)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return max(1, min(n, args.cpu_threads) if args.cpu_threads > 0 else n)


def equiv_factor(w, nelems):
    """he30-equivalent SYPD = SYPD_raw · (dt_1/dt_N) · (elements_N/elements_1) for the weak series; 1 otherwise."""
    if w["scaling"] != "weak" or w.get("series") != "he30_per_gpu":
        return 1.0
    return (90.0 / w["dt"]) * (nelems / 5400.0)


def oracle_steps(w, P, cores, steps, warm, budget_s, FT=np.float32):
    """Time `steps` oracle steps (after `warm`) on the ACTUAL grid of workload `w`; stops early when `budget_s` is used up.
    Returns (seconds per step, steps done, warm-ups done, columns)."""
    from concurrent.futures import ThreadPoolExecutor

    from climaatmos_jl_b200 import grid as G, params as prm, setups
    from oracle.dycore_oracle import Oracle

    g = G.make_sphere_grid(FT=FT, h_elem=w["h_elem"], z_elem=w["z_elem"], z_max=w["z_max"], dz_bottom=w["dz_bottom"], radius=P.planet_radius)
    sponge = w.get("sponge", True)
    hs = w.get("rad") == "held_suarez"
    N = prm.DycoreNumerics(dt=w["dt"], rayleigh_sponge=sponge, viscous_sponge=sponge, held_suarez=hs, disable_momentum_vertical_diffusion=hs,
                           vert_diff=w.get("vert_diff"), implicit_diffusion=bool(w.get("implicit_diffusion", False)), approximate_linear_solve_iters=2,
                           microphysics_model="0M" if w.get("moist") else None)
    o = Oracle(g, P, N, FT)
    Yc, Yf = (setups.decaying_profile(g, P) if w.get("ic") == "DecayingProfile" else
              setups.moist_baroclinic_wave(g, P) if w.get("moist") else setups.dry_baroclinic_wave(g, P))
    for _ in range(w.get("tracers", 0)):
        chi = 0.5 * (1 + np.sin(np.radians(g.lat[..., None])) * np.cos(np.radians(g.lon[..., None]))) * np.exp(-np.broadcast_to(g.z_c, Yc[:, 0].shape) / 8000.0)
        Yc = np.concatenate([Yc, (Yc[:, 0].astype(np.float64) * chi).astype(FT)[:, None]], axis=1)
    t_start = time.time()
    done_w = done = 0
    with ThreadPoolExecutor(cores) as pool:
        for _ in range(warm):
            Yc, Yf = o.step(Yc, Yf, pool=pool, nchunks=cores)
            done_w += 1
        t0 = time.time()
        for _ in range(steps):
            Yc, Yf = o.step(Yc, Yf, pool=pool, nchunks=cores)
            done += 1
            if time.time() - t_start > budget_s:
                break
        t_step = (time.time() - t0) / done
    return t_step, done, done_w, g.ncols, g.nelems


def run_reference(args):
    """CPU arm: the NumPy oracle (kind 'port') on this box's host cores, on the SAME grid, time step and options as the b200 arm
    at this N (no column scaling); rank 0 only.  Honours --steps/--warmup up to --ref-budget seconds and prints the real counts."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from climaatmos_jl_b200 import params as prm

    w = workload(args.gpus, args.config)
    P = prm.DycoreParams(zd_rayleigh=ZD, zd_viscous=ZD, D_0_diffusion=5.0, H_diffusion=800.0)
    cores = cpu_threads(args)
    t_step, steps, warm, ncols, nelems = oracle_steps(w, P, cores, max(1, args.steps), min(max(0, args.warmup), 1), args.ref_budget)
    ms = t_step * 1e3
    raw = (w["dt"] / (365 * 86400.0)) / (ms * 1e-3 / 86400.0)
    val = raw * equiv_factor(w, nelems)
    sample = (f"{steps} timed + {warm} warm-up NumPy-oracle step(s) on the full he{w['h_elem']}/ze{w['z_elem']} Float32 grid ({ncols} columns, dt {w['dt']:.0f} s), "
              f"element-chunked over {cores} host threads; no scaling")
    line = {
        "impl": "reference", "metric": "sypd", "value": val, "unit": "SYPD", "sypd_raw": raw, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "requested": {"steps": args.steps, "warmup": args.warmup, "budget_s": args.ref_budget},
        "ms_per_step": ms, "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(w), "h_elem": w["h_elem"], "z_elem": w["z_elem"], "dt_s": w["dt"], "elements_total": nelems,
                   "columns_total": ncols, "value_definition": VALUE_DEF},
        "cpu_baseline": {"value": val, "unit": "SYPD", "ms_per_step": ms, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "SYPD", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


VALUE_DEF = ("he30-equivalent SYPD = SYPD_raw·(dt_1/dt_N)·(elements_N/elements_1); equals the plain SYPD at N=1 and makes "
             "value_N/(N·value_1) the weak-scaling efficiency normalised by elements per GPU (SURVEY.md §8d.4)")


def workload_name(w):
    sp = w.get("sponge", True)
    return (f"{w['name']} he{w['h_elem']} ze{w['z_elem']} Float32 dt={w['dt']:.0f}s (ARS343, hyperdiffusion" + (", Rayleigh+viscous sponge)" if sp else ")"))


def pin_to_gpu_numa_node(local_rank):
    """Run this rank (and first-touch its pinned staging buffers) on the NUMA node its GPU hangs off, so that the e2e copies of the
    ranks of one box do not all stage through one socket's memory.  Best effort: any failure leaves the affinity untouched.
    Returns (node or None, one-line diagnosis) — the diagnosis goes into the e2e object so that a box that hides its topology
    (a VM reporting numa_node = -1 or a single node) is visible in the bench line."""
    try:
        import glob
        import torch

        nodes = len(glob.glob("/sys/devices/system/node/node[0-9]*"))
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{getattr(pr, 'pci_domain_id', 0):04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        path = f"/sys/bus/pci/devices/{bdf}/numa_node"
        if not os.path.exists(path):
            return None, f"{nodes} NUMA node(s) visible; {path} missing"
        node = int(open(path).read())
        if node < 0:
            return None, f"{nodes} NUMA node(s) visible; GPU {bdf} reports numa_node = {node} (topology hidden)"
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
            return node, f"{nodes} NUMA node(s) visible; rank pinned to node {node} ({len(ids)} cpus)"
        return None, f"{nodes} NUMA node(s) visible; node {node} has no cpu in this process's affinity mask"
    except Exception as ex:  # noqa: BLE001
        return None, f"probe failed: {type(ex).__name__}: {ex}"


def multi_gpu_check(comms):
    """Correctness evidence for N > 1 inside the bench line: 3 fused steps of a small global problem (he8/ze15 Float32, sponges) on
    all ranks through the halo, against the SAME problem stepped on this rank's GPU alone — every rank's owned elements must agree
    BITWISE (the DSS sums collocated nodes in ascending global element order on every rank)."""
    import torch
    from climaatmos_jl_b200 import dycore, params as prm

    P = prm.DycoreParams(zd_rayleigh=20000.0, zd_viscous=20000.0)
    kw = dict(FT=np.float32, h_elem=8, z_elem=15, z_max=30000.0, dz_bottom=300.0, dt=150.0, rayleigh_sponge=True, viscous_sponge=True, params=P)
    sim = dycore.AtmosSimulation(comms=comms, **kw)
    for _ in range(3):
        sim.step(True)
    torch.cuda.synchronize()
    gc, gf = sim.Y.cpu()
    own = sim.part.elems_ext[: sim.part.nh]
    halo = "nvlink-peer-memory" if getattr(sim, "peer_halo", False) else "nccl-send-recv"
    sim.close()
    comms.barrier()
    ref = dycore.AtmosSimulation(**kw)
    for _ in range(3):
        ref.step(True)
    torch.cuda.synchronize()
    rc, rf = ref.Y.cpu()
    ref.close()
    ok = bool(np.array_equal(gc, rc[own]) and np.array_equal(gf, rf[own]))
    import zlib

    crc = zlib.crc32(gc.tobytes()) ^ zlib.crc32(gf.tobytes())
    return {"grid": "dry_baroclinic_wave he8 ze15 Float32, 3 fused steps", "bitwise_equal_to_single_gpu_all_ranks": comms.all_true(ok),
            "halo": halo, "rank0_state_crc32": int(crc)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-budget", type=float, default=240.0, help="--impl reference: stop starting new oracle steps after this many seconds")
    ap.add_argument("--cpu-threads", type=int, default=0, help="upper bound on host threads used by the CPU arm (0 = all cores)")
    ap.add_argument("--min-timed-s", type=float, default=0.6, help="repeat the K-step block until this much device time has been measured")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-multi-gpu-check", action="store_true", help="skip the N > 1 bitwise check against a single-GPU run")
    ap.add_argument("--config", default="weak", choices=["weak", "strong", "he16", "tracer", "moist", "hs", "vdiff", "vdiff_implicit"], help="BASELINE.json config (default: the north-star weak series)")
    ap.add_argument("--unfused", action="store_true", help="hook-by-hook implicit stage instead of the fused kernel")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from climaatmos_jl_b200 import dycore, params as prm
    from climaatmos_jl_b200.parallel import DistributedComms

    comms = DistributedComms()
    rank, nranks = comms.rank, comms.nranks
    if nranks != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={nranks}; using WORLD_SIZE", file=sys.stderr)
    W = max(args.warmup, 3)
    K = args.steps
    w = workload(nranks, args.config)
    P = prm.DycoreParams(zd_rayleigh=ZD, zd_viscous=ZD, D_0_diffusion=5.0, H_diffusion=800.0)
    sponge = w.get("sponge", True)
    ntr = w.get("tracers", 0)
    tracers = [lambda lat, lon, z: 0.5 * (1 + np.sin(np.radians(lat)) * np.cos(np.radians(lon))) * np.exp(-z / 8000.0)] * ntr or None
    ntr += 1 if w.get("moist") else 0  # tracer components of Y.c (ρq_tot counts)
    sim = dycore.AtmosSimulation(FT=np.float32, h_elem=w["h_elem"], z_elem=w["z_elem"], z_max=w["z_max"], dz_bottom=w["dz_bottom"],
                                 dt=w["dt"], rayleigh_sponge=sponge, viscous_sponge=sponge, params=P, rad=w.get("rad"), tracers=tracers,
                                 initial_condition=w.get("ic", "DryBaroclinicWave"), microphysics_model="0M" if w.get("moist") else None,
                                 vert_diff=w.get("vert_diff"), implicit_diffusion=w.get("implicit_diffusion", False), approximate_linear_solve_iters=2,
                                 comms=comms if nranks > 1 else None)
    fused = not args.unfused
    nh_local = sim.Y.c.shape[0]
    ncols_total = sim.grid.ncols

    def timed(nsteps, body, tail=None):
        comms.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(nsteps):
            body()
        if tail is not None:
            tail()  # make the timing stream wait for work queued on side streams
        e1.record()
        torch.cuda.synchronize()
        t1 = time.time()
        comms.barrier()
        return comms.max_over_ranks(e0.elapsed_time(e1)) / nsteps, t0, t1

    # ---- device-resident throughput (`value`): W warm-up steps, then blocks of EXACTLY K steps (barrier + synchronize on both sides,
    # CUDA events, max over ranks) repeated until ≥ --min-timed-s of device time is covered; the median block is reported
    for _ in range(W):
        sim.step(fused)
    sampler = ClockSampler(comms.local_rank) if rank == 0 else None
    l0 = sim.launch_count()
    blocks = []
    t0 = None
    while True:
        m, ta, tb = timed(K, lambda: sim.step(fused))
        t0 = ta if t0 is None else t0
        t1 = tb
        blocks.append(m)
        # every rank sees the same max-over-ranks block times, so the loop count agrees across ranks
        if sum(blocks) * K * 1e-3 >= args.min_timed_s or len(blocks) >= 200:
            break
    launches = (sim.launch_count() - l0) // len(blocks)
    ms = float(np.median(blocks))
    clocks = sampler.stop(t0, t1) if sampler else None
    finite = bool(torch.isfinite(sim.Y.c).all().item())

    # ---- end to end through the public API with HOST buffers (`e2e`): EVERY step copies its input state from
    # pinned host memory to the device, steps it through the C-ABI, and copies the stepped state back to pinned
    # host memory.  The three legs run on three CUDA streams over NB = 3 device/host state buffers, so the copy-in of step
    # k+1 and the copy-out of step k-1 overlap the compute of step k (PCIe is full duplex).  With two buffers each buffer's
    # cycle copy-out → copy-in → compute is serial and bounds the step at (D2H + H2D + compute)/2 = 3.03 ms; the third
    # buffer lets the three legs run concurrently (round 2).
    numa_node, numa_diag = pin_to_gpu_numa_node(comms.local_rank) if nranks > 1 else (None, "single rank: not pinned")
    FieldVector = dycore.FieldVector
    NB = 3
    dev = [sim.Y] + [sim.Y.clone() for _ in range(NB - 1)]
    h_in = [(torch.empty(sim.Y.c.shape, dtype=sim.Y.c.dtype, pin_memory=True), torch.empty(sim.Y.f.shape, dtype=sim.Y.f.dtype, pin_memory=True))
            for _ in range(NB)]
    h_out = [(torch.empty_like(h_in[0][0]).pin_memory(), torch.empty_like(h_in[0][1]).pin_memory()) for _ in range(NB)]
    for hc_, hf_ in h_in:
        hc_.copy_(sim.Y.c)
        hf_.copy_(sim.Y.f)
    torch.cuda.synchronize()
    s_in, s_cmp, s_out = torch.cuda.Stream(), torch.cuda.current_stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(NB)]
    ev_cmp = [torch.cuda.Event() for _ in range(NB)]
    ev_out = [torch.cuda.Event() for _ in range(NB)]
    state = {"k": 0, "compute": True}

    def upload(k):
        b = k % NB
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_out[b])  # buffer b was last read by the copy-out of step k-NB
            dev[b].c.copy_(h_in[b][0], non_blocking=True)
            dev[b].f.copy_(h_in[b][1], non_blocking=True)
            ev_in[b].record(s_in)

    def e2e_step():
        k = state["k"]
        b = k % NB
        if k == 0:
            upload(0)
        upload(k + 1)  # prefetch the next step's input while this step computes
        s_cmp.wait_event(ev_in[b])
        sim.Y = dev[b]
        if state["compute"]:
            sim.step(fused)
        ev_cmp[b].record(s_cmp)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_cmp[b])
            h_out[b][0].copy_(dev[b].c, non_blocking=True)
            h_out[b][1].copy_(dev[b].f, non_blocking=True)
            ev_out[b].record(s_out)
        state["k"] = k + 1

    for _ in range(NB):
        e2e_step()
    torch.cuda.synchronize()
    Ke = max(4, K)  # blocks of K steps like the device-resident measurement (the pipeline fills and drains once per block)
    e2e_blocks = []
    while True:
        m, _, _ = timed(Ke, e2e_step, tail=lambda: [s_cmp.wait_event(e) for e in ev_out])
        e2e_blocks.append(m)
        if sum(e2e_blocks) * Ke * 1e-3 >= args.min_timed_s or len(e2e_blocks) >= 50:
            break
    ms_e2e = float(np.median(e2e_blocks))
    torch.cuda.synchronize()
    e2e_ok = bool(torch.isfinite(h_out[0][0]).all().item())
    # the same pipeline with the step left out (all ranks at once): what the box's host <-> device path alone allows
    state["compute"] = False
    copy_blocks = [timed(Ke, e2e_step, tail=lambda: [s_cmp.wait_event(e) for e in ev_out])[0] for _ in range(3)]
    state["compute"] = True
    ms_copies = float(np.median(copy_blocks))
    state_bytes = int(h_in[0][0].numel() * 4 + h_in[0][1].numel() * 4)
    sim.Y = dev[0]

    # ---- the three compute kernels timed alone with CUDA events on their stream (the dominant one is `roofline`)
    nv = w["z_elem"]
    c_b = nh_local * 16 * nv * 4
    f_b = nh_local * 16 * (nv + 1) * 4
    S_b, H_b = (4 + ntr) * c_b + f_b, (4 + ntr) * c_b
    Yt = sim.Y.zeros_like()
    N2 = sim.Y.zeros_like()

    def time_kernel(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    kern = []
    if ntr == 0 and not w.get("vert_diff"):
        dtg = w["dt"] * 0.4358665215084590
        sim.remaining_tendency_phase_a(Yt, sim.Y)  # fills H for phase C
        sim.remaining_tendency_phase(1, Yt, sim.Y)  # DSS of the ∇² fields
        for name, fn, nbytes, per_step, what in (
                ("k5_exp_a<float, 63>", lambda: sim.remaining_tendency_phase_a(Yt, sim.Y), 2 * S_b + H_b, 4, "T_exp_T_lim! pre-DSS kernel: read S, write S + H"),
                ("k8_imp_stage<float, 63>", lambda: sim.implicit_stage(N2, sim.Y, dtg), 2 * S_b, 3, "fused implicit stage (warp per column pair): read S, write S"),
                ("k7_exp_c<float, 63>", lambda: sim.remaining_tendency_phase_c(Yt, sim.Y), c_b + H_b + 2 * (3 * c_b + f_b), 4,
                 "hyperdiffusion apply: read ρ + H, read-modify-write uₕ, ρe_tot, u₃ of Yₜ")):
            k_ms = time_kernel(fn)
            kern.append({"kernel": name, "what": what, "ms_per_launch": k_ms, "launches_per_step": per_step, "share_of_step": per_step * k_ms / ms,
                         "algorithmic_bytes_per_launch": nbytes, "achieved": nbytes / (k_ms * 1e-3) / 1e9, "frac": nbytes / (k_ms * 1e-3) / 1e9 / peak})
    else:
        k_ms = time_kernel(lambda: sim.remaining_tendency_phase_a(Yt, sim.Y))
        nbytes = 2 * (4 * c_b + f_b) + 4 * c_b
        kern.append({"kernel": "k5_exp_a<float, 0, MOIST>" if w.get("moist") else "k5_exp_a<float, 63>", "what": "T_exp_T_lim! pre-DSS kernel (dry components)", "ms_per_launch": k_ms, "launches_per_step": 4,
                     "share_of_step": 4 * k_ms / ms, "algorithmic_bytes_per_launch": nbytes, "achieved": nbytes / (k_ms * 1e-3) / 1e9,
                     "frac": nbytes / (k_ms * 1e-3) / 1e9 / peak})
    dom = max(kern, key=lambda r: r["share_of_step"])
    mg = multi_gpu_check(comms) if nranks > 1 and not args.no_multi_gpu_check else None

    if rank == 0:
        sy = lambda m: (w["dt"] / (365 * 86400.0)) / (m * 1e-3 / 86400.0)
        eq = equiv_factor(w, sim.grid.nelems)
        step_bytes = model_bytes_per_step(ncols_total, nv, k=ntr)
        # ncu --set full traffic per launch at he30/ze63 (dram__bytes_read.sum + dram__bytes_write.sum), from the capture summarised
        # in profiles/ (TRAFFIC_SOURCE); scaled by elements per GPU
        traffic = TRAFFIC.get(dom["kernel"])
        line = {
            "metric": "sypd", "value": sy(ms) * eq, "unit": "SYPD", "sypd_raw": sy(ms), "n_gpus": nranks, "steps": K, "warmup": W, "ms_per_step": ms,
            "timed_blocks": len(blocks), "ms_per_step_blocks": [round(b, 5) for b in blocks[:40]],
            "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": workload_name(w),
                "h_elem": w["h_elem"], "z_elem": nv, "dt_s": w["dt"], "elements_total": sim.grid.nelems, "elements_per_gpu": nh_local,
                "columns_total": ncols_total, "parallelism": f"sfc-domain-decomposition x{nranks}",
                "halo": ("nvlink-peer-memory" if getattr(sim, "peer_halo", False) else "nccl-send-recv") if nranks > 1 else "none", "implicit_stage": "fused" if fused else "hooks",
                "launch": "CUDA graph replay of the step + programmatic dependent launch" if fused else "eager hook-by-hook",
                "l2_policy": "working set (≈1.2 GB of stage vectors per step) larger than the 126 MB L2; no explicit flush",
                "value_definition": VALUE_DEF,
            },
            "finite_state": finite,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "e2e": {"value": sy(ms_e2e) * eq, "unit": "SYPD", "sypd_raw": sy(ms_e2e), "ms_per_step": ms_e2e, "h2d_bytes_per_step": state_bytes, "d2h_bytes_per_step": state_bytes,
                    "pipeline": "3 streams (copy-in / step / copy-out) over 3 state buffers; pinned staging buffers first-touched on the GPU's NUMA node",
                    "copies_only_ms_per_step": ms_copies, "numa_node": numa_node, "numa_diag": numa_diag, "finite": e2e_ok, "timed_blocks": len(e2e_blocks)},
            "roofline": {"bound": "hbm", "kernel": f"{dom['kernel']} ({dom['what']}; the largest share of the step)",
                         "achieved": dom["achieved"], "peak": peak, "unit": "GB/s", "frac": dom["frac"],
                         "traffic": traffic * nh_local / 5400.0 if traffic else None, "traffic_source": TRAFFIC_SOURCE if traffic else None,
                         "ms_per_launch": dom["ms_per_launch"], "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"],
                         "share_of_step": dom["share_of_step"],
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 GB/s"},
            "roofline_kernels": kern,
            "roofline_step": {"model_bytes_per_step": step_bytes, "achieved_gbs": step_bytes / (ms * 1e-3) / 1e9 / nranks,
                              "frac": step_bytes / (ms * 1e-3) / 1e9 / nranks / peak, "model": "54.5 S + 14 H (SURVEY.md §8d)"},
        }
        if mg is not None:
            line["multi_gpu_check"] = mg
        if nranks == 1 and not args.no_cpu_baseline:
            # the oracle (kind "port") on the host cores of this box: a bounded sample of the SAME workload — 1 warm-up + 2 timed steps on
            # the full grid (≈ 15-25 s of CPU work), no scaling
            cores = cpu_threads(args)
            t_step, nst, nwu, ncols_cpu, _ = oracle_steps(w, P, cores, 2, 1, 60.0)
            line["cpu_baseline"] = {"value": sy(t_step * 1e3) * eq, "unit": "SYPD", "ms_per_step": t_step * 1e3, "cores": cores, "kind": "port",
                                    "sample": f"{nst} timed + {nwu} warm-up NumPy-oracle step(s) on the full he{w['h_elem']}/ze{nv} Float32 grid ({ncols_cpu} columns), "
                                              f"element-chunked over {cores} host threads; no scaling"}
        print(json.dumps(line), flush=True)
    sim.close()
    comms.finalize()


# ncu --set full DRAM traffic per launch at he30/ze63 Float32 (bytes): see profiles/ (updated per round with the capture)
TRAFFIC_SOURCE = "profiles/r2_ncu_full_final_step_kernels.txt"
TRAFFIC = {"k5_exp_a<float, 63>": 255.7e6, "k8_imp_stage<float, 63>": 174.4e6, "k7_exp_c<float, 63>": 264.1e6}


if __name__ == "__main__":
    main()
