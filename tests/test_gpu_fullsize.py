"""GPU parity at the BASELINE.json sizes (driver-visible, `-m gpu`): Float32 CUDA through the C-ABI against the Float64 oracle on the
dry baroclinic wave he16/ze63 (configs[1]) and he30/ze63 (the north star), one fused, graph-replayable ARS343 step, every prognostic
field held to the north-star bar rel-L2 ≤ 1e-5; and a 100-step drift test with the ze63 numerics that includes u₃.

The oracle runs element-chunked over the host threads (≈ 8 s per he30 step on 16 cores).  Measured errors are appended to
gpurun_out/parity_report.jsonl (copied to profiles/ per round)."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from climaatmos_jl_b200 import dycore, params as prm
from oracle.dycore_oracle import Oracle
from tests.conftest import record_parity

torch = pytest.importorskip("torch")

NAMES = ("rho", "u1", "u2", "rhoe")
ZD = 40000.0  # zd_rayleigh = zd_viscous as in bench.py (toml/longrun_held_suarez.toml)


def rel(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


def errors(sim, oc, of):
    gc, gf = sim.Y.cpu()
    e = {n: rel(gc[:, k], oc[:, k]) for k, n in enumerate(NAMES)}
    e["u3"] = rel(gf[:, 0], of[:, 0])
    return e


@pytest.mark.parametrize("h_elem,dt", [(16, 120.0), (30, 90.0)])
def test_baseline_configs_one_step_float32_within_1e5(h_elem, dt):
    """BASELINE.json configs[1] (he16/ze63, dt 120 s) and the north-star he30/ze63 (dt 90 s): rel-L2 per prognostic field after one step
    ≤ 1e-5 — ALL five fields, u₃ included (the Float32 kernels evaluate the vertical pressure-gradient differences in difference form,
    common.cuh pgf_diff; the reference's literal Float32 formulation itself sits at ≈1.5e-5, test_oracle_identities.py)."""
    P = prm.DycoreParams(zd_rayleigh=ZD, zd_viscous=ZD)
    sim = dycore.AtmosSimulation(FT=np.float32, h_elem=h_elem, z_elem=63, z_max=60000.0, dz_bottom=30.0, dt=dt, rayleigh_sponge=True,
                                 viscous_sponge=True, params=P)
    Yc0, Yf0 = sim.Y.cpu()
    o = Oracle(sim.grid, P, sim.numerics, np.float64)
    cores = max(1, min(32, len(os.sched_getaffinity(0))))
    with ThreadPoolExecutor(cores) as pool:
        oc, of = o.step(Yc0.astype(np.float64), Yf0.astype(np.float64), pool=pool, nchunks=cores)
    sim.step(fused=True)
    torch.cuda.synchronize()
    e = errors(sim, oc, of)
    record_parity(f"one_step_he{h_elem}ze63_f32", e)
    for n, v in e.items():
        assert v <= 1e-5, f"he{h_elem}/ze63 Float32 after one step: {n} rel-L2 {v:.3e} > 1e-5 ({e})"
    sim.close()


def test_hundred_steps_bounded_drift_ze63_with_u3():
    """Bounded drift over 100 steps (north star) with the vertical grid, sponges and time step class of the ze63 configs (he6/ze63,
    dt 300 s ≈ the he16 Courant number): Float32 CUDA vs Float64 oracle, all five fields, recorded at steps 1, 25, 50, 100.
    Bounds = what Float32 arithmetic itself does to this flow, measured with the oracle (the reference's literal formulation) run
    in Float32 against itself in Float64 (profiles/r2_drift_float32.md): ρ 3.2e-7, uₕ 2.5e-5 / 4.4e-5, ρe_tot 1.2e-6, and u₃ — a
    near-zero field whose round-off excites small acoustic/gravity-wave adjustments — 3.5e-2 at step 25, settling to 2.3e-2 at
    step 100 (no growth).  The CUDA path measures 3.7e-7, 2.9e-5 / 5.1e-5, 1.4e-6, and 2.5e-2 → 1.7e-2 for u₃: the same drift."""
    P = prm.DycoreParams(zd_rayleigh=ZD, zd_viscous=ZD)
    sim = dycore.AtmosSimulation(FT=np.float32, h_elem=6, z_elem=63, z_max=60000.0, dz_bottom=30.0, dt=300.0, rayleigh_sponge=True,
                                 viscous_sponge=True, params=P)
    o = Oracle(sim.grid, P, sim.numerics, np.float64)
    oc, of = [a.astype(np.float64) for a in sim.Y.cpu()]
    cores = max(1, min(32, len(os.sched_getaffinity(0))))
    hist = {}
    with ThreadPoolExecutor(cores) as pool:
        for k in range(1, 101):
            sim.step(fused=True)
            oc, of = o.step(oc, of, pool=pool, nchunks=cores)
            if k in (1, 25, 50, 100):
                torch.cuda.synchronize()
                hist[k] = errors(sim, oc, of)
    record_parity("drift_100_steps_he6ze63_f32", {str(k): v for k, v in hist.items()})
    gc, gf = sim.Y.cpu()
    assert np.isfinite(gc).all() and np.isfinite(gf).all()
    e1, e100 = hist[1], hist[100]
    assert max(e1.values()) <= 1e-5, e1
    for k in (25, 50, 100):
        e = hist[k]
        assert e["rho"] <= 2e-6 and e["rhoe"] <= 5e-6, (k, e)
        assert e["u1"] <= 1e-4 and e["u2"] <= 1.5e-4, (k, e)
        assert e["u3"] <= 5e-2, (k, e)
    assert e100["u3"] <= 1.5 * max(hist[25]["u3"], hist[50]["u3"]), hist  # bounded, not growing
    sim.close()


def test_moist_config_one_step_float32_he30():
    """BASELINE.json configs[2] as the reference runs it — the 0M-moist baroclinic wave at he30/ze63 with the thermodynamically active
    ρq_tot (dt 90 s): Float32 CUDA after one fused step against the Float64 oracle.  ρ, uₕ, ρe_tot, ρq_tot within the 1e-5 north-star
    bar; u₃ within the moist Float32 bar of tests/test_gpu_moist.py (the floor of the reference's own formulation in Float32 there is
    2.4e-5 … 7.9e-5, tests/test_oracle_moist.py::test_float32_floor_of_the_moist_reference_formulation).  Water and mass conserved."""
    from tests.test_gpu_moist import U3_MOIST

    P = prm.DycoreParams(zd_rayleigh=ZD, zd_viscous=ZD)
    sim = dycore.AtmosSimulation(FT=np.float32, h_elem=30, z_elem=63, z_max=60000.0, dz_bottom=30.0, dt=90.0, rayleigh_sponge=True,
                                 viscous_sponge=True, params=P, microphysics_model="0M", initial_condition="MoistBaroclinicWave")
    Yc0, Yf0 = sim.Y.cpu()
    o = Oracle(sim.grid, P, sim.numerics, np.float64)
    cores = max(1, min(32, len(os.sched_getaffinity(0))))
    with ThreadPoolExecutor(cores) as pool:
        oc, of = o.step(Yc0.astype(np.float64), Yf0.astype(np.float64), pool=pool, nchunks=cores)
    sim.step(fused=True)
    torch.cuda.synchronize()
    gc, gf = sim.Y.cpu()
    e = {n: rel(gc[:, k], oc[:, k]) for k, n in enumerate(NAMES + ("rhoq",))}
    e["u3"] = rel(gf[:, 0], of[:, 0])
    record_parity("one_step_moist_he30ze63_f32", e)
    for n, v in e.items():
        assert v <= (U3_MOIST if n == "u3" else 1e-5), f"moist he30/ze63 Float32 after one step: {n} rel-L2 {v:.3e} ({e})"
    W = o.c.WJ
    for k in (0, 4):
        a, b = (W * Yc0[:, k].astype(np.float64)).sum(), (W * gc[:, k].astype(np.float64)).sum()
        assert abs(a - b) / abs(a) < 2e-6, (k, a, b)
    sim.close()
