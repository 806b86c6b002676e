#!/usr/bin/env python
"""bench.py — SYPD and ms/step of the dry baroclinic-wave dycore step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N = 1; N > 1 under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one ARS343 IMEX step of the whole dycore (4 × T_exp_T_lim!, 3 × implicit stage, 7 state DSS
+ 4 hyperdiffusion DSS, stage increments).  N = 1 runs the configuration the metric is quoted on
(dry baroclinic wave, h_elem = 30, z_elem = 63, Float32, dt = 90 s); N > 1 is the weak-scaling series
"he30 per GPU": h_elem = 30/42/60/85 for N = 1/2/4/8 with dt ∝ 1/h_elem (90/64/45/32 s).
Prints ONE JSON line on rank 0.  `--impl reference` times the CPU restatement of the reference
(oracle/, NumPy "port": the Julia reference cannot run here — no julia, ClimaCore un-vendored) on a
bounded sample of the same workload.
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
    "strong": dry BW he60/ze63 on every N (configs[4]); "he16": dry BW he16/ze63 (configs[1]); "tracer": he30/ze63 with one passive
    tracer (configs[2], dycore + tracer advection only); "hs": Held–Suarez he6/ze10 (configs[0])."""
    if config == "strong":
        return dict(h_elem=60, z_elem=63, z_max=60000.0, dz_bottom=30.0, dt=45.0, scaling="strong", name="dry_baroclinic_wave")
    if config == "he16":
        return dict(h_elem=16, z_elem=63, z_max=60000.0, dz_bottom=30.0, dt=120.0, scaling="weak", name="dry_baroclinic_wave")
    if config == "hs":
        return dict(h_elem=6, z_elem=10, z_max=55000.0, dz_bottom=500.0, dt=400.0, scaling="weak", name="held_suarez", sponge=False, rad="held_suarez")
    h = WEAK_H.get(n_gpus, int(round(30 * np.sqrt(n_gpus))))
    dt = float(round(90.0 * 30 / h))
    w = dict(h_elem=h, z_elem=63, z_max=60000.0, dz_bottom=30.0, dt=dt, scaling="weak", name="dry_baroclinic_wave")
    if config == "tracer":
        w.update(tracers=1, name="dry_baroclinic_wave + 1 passive tracer")
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
    """Host threads for the CPU arm: all cores the box offers, capped where NumPy-under-GIL stops scaling."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return max(1, min(n, args.cpu_threads))


def run_reference(args):
    """CPU arm: the NumPy oracle (kind 'port') on this box's host cores; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from climaatmos_jl_b200 import grid as G, params as prm, setups
    from oracle.dycore_oracle import Oracle

    w = workload(args.gpus)
    P = prm.DycoreParams(zd_rayleigh=ZD, zd_viscous=ZD, D_0_diffusion=5.0, H_diffusion=800.0)
    # bounded sample: a full-depth (ze63) sphere at reduced horizontal resolution, cost ∝ columns
    h_s = min(w["h_elem"], args.ref_h_elem)
    g = G.make_sphere_grid(FT=np.float32, h_elem=h_s, z_elem=w["z_elem"], z_max=w["z_max"], dz_bottom=w["dz_bottom"],
                           radius=P.planet_radius)
    N = prm.DycoreNumerics(dt=w["dt"], rayleigh_sponge=True, viscous_sponge=True)
    o = Oracle(g, P, N, np.float32)
    Yc, Yf = setups.dry_baroclinic_wave(g, P)
    steps, warm = max(1, min(args.steps, 3)), min(args.warmup, 1)
    from concurrent.futures import ThreadPoolExecutor

    cores = cpu_threads(args)
    with ThreadPoolExecutor(cores) as pool:
        for _ in range(warm):
            Yc, Yf = o.step(Yc, Yf, pool=pool, nchunks=cores)
        t0 = time.time()
        for _ in range(steps):
            Yc, Yf = o.step(Yc, Yf, pool=pool, nchunks=cores)
        t_step = (time.time() - t0) / steps
    scale = (w["h_elem"] / h_s) ** 2  # columns of the full workload / columns of the sample
    ms = t_step * scale * 1e3
    sypd = (w["dt"] / (365 * 86400.0)) / (ms * 1e-3 / 86400.0)
    sample = (f"{steps} oracle step(s) on he{h_s}/ze63 Float32 ({g.ncols} of {96 * w['h_elem'] ** 2} columns), element-chunked over "
              f"{cores} host threads, time scaled by columns")
    line = {
        "impl": "reference", "metric": "sypd", "value": sypd, "unit": "SYPD", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"dry_baroclinic_wave he{w['h_elem']} ze63 Float32 dt={w['dt']:.0f}s (ARS343, hyperdiffusion, Rayleigh+viscous sponge)"},
        "cpu_baseline": {"value": sypd, "unit": "SYPD", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": sypd, "unit": "SYPD", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-h-elem", type=int, default=16, help="horizontal resolution of the bounded CPU sample")
    ap.add_argument("--cpu-threads", type=int, default=16, help="upper bound on host threads used by the CPU arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="weak", choices=["weak", "strong", "he16", "tracer", "hs", "vdiff", "vdiff_implicit"], help="BASELINE.json config (default: the north-star weak series)")
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
    sim = dycore.AtmosSimulation(FT=np.float32, h_elem=w["h_elem"], z_elem=w["z_elem"], z_max=w["z_max"], dz_bottom=w["dz_bottom"],
                                 dt=w["dt"], rayleigh_sponge=sponge, viscous_sponge=sponge, params=P, rad=w.get("rad"), tracers=tracers,
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

    # ---- device-resident throughput (`value`)
    for _ in range(W):
        sim.step(fused)
    sampler = ClockSampler(comms.local_rank) if rank == 0 else None
    l0 = sim.launch_count()
    ms, t0, t1 = timed(K, lambda: sim.step(fused))
    launches = sim.launch_count() - l0
    clocks = sampler.stop(t0, t1) if sampler else None
    finite = bool(torch.isfinite(sim.Y.c).all().item())

    # ---- end to end through the public API with HOST buffers (`e2e`): EVERY step copies its input state from
    # pinned host memory to the device, steps it through the C-ABI, and copies the stepped state back to pinned
    # host memory.  The three legs run on three CUDA streams with double-buffered device/host states, so the
    # copy-in of step k+1 and the copy-out of step k-1 overlap the compute of step k (PCIe is full duplex).
    FieldVector = dycore.FieldVector
    dev = [sim.Y, sim.Y.clone()]
    h_in = [(torch.empty(sim.Y.c.shape, dtype=sim.Y.c.dtype, pin_memory=True), torch.empty(sim.Y.f.shape, dtype=sim.Y.f.dtype, pin_memory=True))
            for _ in range(2)]
    h_out = [(torch.empty_like(h_in[0][0]).pin_memory(), torch.empty_like(h_in[0][1]).pin_memory()) for _ in range(2)]
    for hc_, hf_ in h_in:
        hc_.copy_(sim.Y.c)
        hf_.copy_(sim.Y.f)
    torch.cuda.synchronize()
    s_in, s_cmp, s_out = torch.cuda.Stream(), torch.cuda.current_stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_cmp = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    state = {"k": 0}

    def upload(k):
        b = k & 1
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_out[b])  # buffer b was last read by the copy-out of step k-2
            dev[b].c.copy_(h_in[b][0], non_blocking=True)
            dev[b].f.copy_(h_in[b][1], non_blocking=True)
            ev_in[b].record(s_in)

    def e2e_step():
        k = state["k"]
        b = k & 1
        if k == 0:
            upload(0)
        upload(k + 1)  # prefetch the next step's input while this step computes
        s_cmp.wait_event(ev_in[b])
        sim.Y = dev[b]
        sim.step(fused)
        ev_cmp[b].record(s_cmp)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_cmp[b])
            h_out[b][0].copy_(dev[b].c, non_blocking=True)
            h_out[b][1].copy_(dev[b].f, non_blocking=True)
            ev_out[b].record(s_out)
        state["k"] = k + 1

    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    Ke = max(4, min(K, 10))
    ms_e2e, _, _ = timed(Ke, e2e_step, tail=lambda: (s_cmp.wait_event(ev_out[0]), s_cmp.wait_event(ev_out[1])))
    torch.cuda.synchronize()
    e2e_ok = bool(torch.isfinite(h_out[0][0]).all().item())
    state_bytes = int(h_in[0][0].numel() * 4 + h_in[0][1].numel() * 4)
    sim.Y = dev[0]

    # ---- dominant kernel (explicit-tendency phase A) timed alone with CUDA events on its stream
    Yt = sim.Y.zeros_like()
    for _ in range(3):
        sim.remaining_tendency_phase_a(Yt, sim.Y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        sim.remaining_tendency_phase_a(Yt, sim.Y)
    e1.record()
    torch.cuda.synchronize()
    k_ms = e0.elapsed_time(e1) / reps
    nv = w["z_elem"]
    c_b = nh_local * 16 * nv * 4
    f_b = nh_local * 16 * (nv + 1) * 4
    k_bytes = (4 * c_b + f_b) * 2 + 4 * c_b  # read Y (dry components), write Yₜ, write H = (∇²u, ∇²s_d)   
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = k_bytes / (k_ms * 1e-3) / 1e9

    if rank == 0:
        sy = lambda m: (w["dt"] / (365 * 86400.0)) / (m * 1e-3 / 86400.0)
        step_bytes = model_bytes_per_step(ncols_total, nv, k=ntr)
        line = {
            "metric": "sypd", "value": sy(ms), "unit": "SYPD", "n_gpus": nranks, "steps": K, "warmup": W, "ms_per_step": ms,
            "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"{w['name']} he{w['h_elem']} ze{w['z_elem']} Float32 dt={w['dt']:.0f}s (ARS343, hyperdiffusion" + (", Rayleigh+viscous sponge)" if sponge else ")"),
                "h_elem": w["h_elem"], "z_elem": nv, "dt_s": w["dt"], "elements_total": sim.grid.nelems, "elements_per_gpu": nh_local,
                "columns_total": ncols_total, "parallelism": f"sfc-domain-decomposition x{nranks}",
                "halo": ("nvlink-peer-memory" if getattr(sim, "peer_halo", False) else "nccl-send-recv") if nranks > 1 else "none", "implicit_stage": "fused" if fused else "hooks",
                "launch": "CUDA graph replay of the step + programmatic dependent launch" if fused else "eager hook-by-hook",
                "l2_policy": "working set (≈1.2 GB of stage vectors per step) larger than the 126 MB L2; no explicit flush",
                "weak_scaling_note": "dt ∝ 1/h_elem; efficiency = (ms_1/ms_N)·(elements_N/(N·elements_1))",
            },
            "finite_state": finite,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "e2e": {"value": sy(ms_e2e), "unit": "SYPD", "ms_per_step": ms_e2e, "h2d_bytes_per_step": state_bytes, "d2h_bytes_per_step": state_bytes,
                    "pipeline": "3 streams (copy-in / step / copy-out), double-buffered", "finite": e2e_ok},
            "roofline": {"bound": "hbm", "kernel": "k5_exp_a<float, 63> (T_exp_T_lim! pre-DSS kernel, the largest share of the step)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch at he30/ze63 (114.1 + 141.9 MB), from the
                         # ncu --set full capture summarised in profiles/r1_ncu_full_session2_kernels.txt; scaled by elements
                         "traffic": 256.0e6 * nh_local / 5400.0, "ms_per_launch": k_ms, "algorithmic_bytes_per_launch": k_bytes,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 GB/s"},
            "roofline_step": {"model_bytes_per_step": step_bytes, "achieved_gbs": step_bytes / (ms * 1e-3) / 1e9 / nranks,
                              "frac": step_bytes / (ms * 1e-3) / 1e9 / nranks / peak, "model": "54.5 S + 14 H (SURVEY.md §8d)"},
        }
        if nranks == 1 and not args.no_cpu_baseline:
            from climaatmos_jl_b200 import grid as G, setups
            from oracle.dycore_oracle import Oracle

            h_s = 16
            g = G.make_sphere_grid(FT=np.float32, h_elem=h_s, z_elem=nv, z_max=w["z_max"], dz_bottom=w["dz_bottom"], radius=P.planet_radius)
            o = Oracle(g, P, sim.numerics, np.float32)
            Yc, Yf = setups.dry_baroclinic_wave(g, P)
            from concurrent.futures import ThreadPoolExecutor

            cores = cpu_threads(args)
            with ThreadPoolExecutor(cores) as pool:
                tc = time.time()
                o.step(Yc, Yf, pool=pool, nchunks=cores)
                t_cpu = (time.time() - tc) * (w["h_elem"] / h_s) ** 2
            line["cpu_baseline"] = {"value": sy(t_cpu * 1e3), "unit": "SYPD", "ms_per_step": t_cpu * 1e3, "cores": cores, "kind": "port",
                                    "sample": f"1 NumPy-oracle step on he{h_s}/ze63 Float32 ({g.ncols} of {ncols_total} columns), element-chunked over {cores} host threads, time scaled by columns"}
        print(json.dumps(line), flush=True)
    sim.close()
    comms.finalize()


if __name__ == "__main__":
    main()
