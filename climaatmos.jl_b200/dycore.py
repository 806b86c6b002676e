"""Host-side mirror of the reference's simulation / hook surface for the dry dycore path.

Names follow the reference (without Julia's ``!``): ``AtmosSimulation``
(src/simulation/AtmosSimulations.jl:307-462), ``solve_atmos`` (src/simulation/solve.jl:119-158),
and the ClimaODEFunction hooks wired at src/simulation/integrator.jl:215-225:
``remaining_tendency`` (T_exp_T_lim!), ``implicit_tendency`` (T_imp!), ``update_jacobian`` (Wfact),
``ldiv`` (ldiv!), ``correct_implicit_advection_tendency`` (T_post_imp!), ``dss``,
``set_implicit_precomputed_quantities`` (cache_imp!).  Every hook mutates its first argument(s) and
forwards to the C-ABI in include/b200_dycore.h; PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import time

import numpy as np

from . import capi
from .grid import make_sphere_grid
from .params import DycoreNumerics, DycoreParams
from . import setups


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("climaatmos_jl_b200 needs a CUDA device: there is no CPU fallback")
    return torch


@dataclasses.dataclass
class FieldVector:
    """``Y``: ``c`` = Y.c parent array [nh, 4, 4, 4, nv] (ρ, uₕ₁, uₕ₂, ρe_tot), ``f`` = Y.f [nh, 1, 4, 4, nv+1]."""

    c: "object"
    f: "object"

    def clone(self):
        return FieldVector(self.c.clone(), self.f.clone())

    def zeros_like(self):
        t = _torch()
        return FieldVector(t.zeros_like(self.c), t.zeros_like(self.f))

    def cpu(self):
        return self.c.detach().cpu().numpy(), self.f.detach().cpu().numpy()


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class AtmosSimulation:
    """Dry cubed-sphere simulation on one rank (one process per GPU).

    Keyword surface follows ``AtmosSimulation{FT}(; …)`` / the YAML keys of
    config/default_configs/default_config.yml: ``h_elem, z_elem, z_max, dz_bottom, dt,
    rayleigh_sponge, viscous_sponge, hyperdiff, initial_condition, deep_atmosphere, vert_diff, implicit_diffusion,
    approximate_linear_solve_iters``.
    """

    def __init__(self, FT=np.float32, h_elem=6, z_elem=10, z_max=30000.0, dz_bottom=500.0, dt=400.0,
                 rayleigh_sponge=False, viscous_sponge=False, hyperdiff=True, deep_atmosphere=True,
                 initial_condition="DryBaroclinicWave", energy_q_tot_upwinding="vanleer_limiter", rad=None,
                 tracers=None, tracer_upwinding="vanleer_limiter", apply_sem_quasimonotone_limiter=False,
                 vert_diff=None, implicit_diffusion=False, approximate_linear_solve_iters=1, tracer_nonnegativity_method=None,
                 microphysics_model=None, q_0=0.018,
                 params: DycoreParams | None = None, device=None, comms=None, grid=None):
        torch = _torch()
        self.torch = torch
        self.FT = np.dtype(FT).type
        self.params = params or DycoreParams()
        self.numerics = DycoreNumerics(dt=float(dt), hyperdiff=hyperdiff, rayleigh_sponge=rayleigh_sponge,
                                       viscous_sponge=viscous_sponge, energy_upwinding=energy_q_tot_upwinding,
                                       held_suarez=(rad == "held_suarez"), tracer_upwinding=tracer_upwinding,
                                       apply_sem_quasimonotone_limiter=bool(apply_sem_quasimonotone_limiter),
                                       # vert_diff / implicit_diffusion / approximate_linear_solve_iters: default_config.yml:166-168,
                                       # 397-402; momentum diffusion is off for Held–Suarez runs (type_getters.jl:46)
                                       vert_diff=vert_diff, implicit_diffusion=bool(implicit_diffusion),
                                       approximate_linear_solve_iters=int(approximate_linear_solve_iters),
                                       disable_momentum_vertical_diffusion=(rad == "held_suarez"),
                                       tracer_nonnegativity_method=tracer_nonnegativity_method,
                                       # microphysics_model: None (dry) | "0M" (EquilibriumMicrophysics0M: ρq_tot is component 4 of Y.c)
                                       microphysics_model=microphysics_model)
        self.grid = grid or make_sphere_grid(FT=self.FT, h_elem=h_elem, z_elem=z_elem, z_max=z_max, dz_bottom=dz_bottom,
                                             radius=self.params.planet_radius, deep_atmosphere=deep_atmosphere)
        self.comms = comms  # parallel.DistributedComms or None
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        if (microphysics_model == "0M") != (initial_condition == "MoistBaroclinicWave"):
            raise ValueError('microphysics_model "0M" goes with initial_condition "MoistBaroclinicWave" (and only with it)')
        if initial_condition == "MoistBaroclinicWave":
            Yc, Yf = setups.moist_baroclinic_wave(self.grid, self.params, q_0=q_0)
        elif initial_condition == "DryBaroclinicWave":
            Yc, Yf = setups.dry_baroclinic_wave(self.grid, self.params)
        elif initial_condition == "DecayingProfile":
            Yc, Yf = setups.decaying_profile(self.grid, self.params)
        else:
            raise ValueError(f"unknown initial_condition {initial_condition}")
        # passive grid-scale tracers ρχ appended after ρe_tot (e.g. the chemistry tracer ρq_gas_A,
        # setups/common/prognostic_variables.jl:138-145): ``tracers`` = list of χ(lat°, lon°, z) callables or arrays
        n_passive = len(tracers) if tracers else 0
        self.n_tracers = n_passive + (1 if microphysics_model == "0M" else 0)  # tracer components of Y.c (ρq_tot first)
        if n_passive:
            zz = np.broadcast_to(self.grid.z_c, Yc[:, 0].shape)
            extra = []
            for tr in tracers:
                chi = tr(self.grid.lat[..., None], self.grid.lon[..., None], zz) if callable(tr) else np.asarray(tr)
                extra.append((Yc[:, 0].astype(np.float64) * chi).astype(self.FT)[:, None])
            Yc = np.concatenate([Yc] + extra, axis=1)
        self.part = None
        if comms is not None and comms.nranks > 1:
            from .partition import partition_grid

            self.part = partition_grid(self.grid, comms.rank, comms.nranks)
            Yc, Yf = Yc[self.part.elems_ext[: self.part.nh]], Yf[self.part.elems_ext[: self.part.nh]]
            self.ctx = capi.create_context(self.grid, self.params, self.numerics, part=self.part,
                                           nccl_id=comms.nccl_unique_id(), rank=comms.rank, nranks=comms.nranks,
                                           n_tracers=self.n_tracers)
            import os as _os

            self.peer_halo = False
            if not _os.environ.get("B200_HALO_NCCL"):
                self.peer_halo = capi.setup_peer_halo(self.ctx, self.part, comms)
        else:
            self.ctx = capi.create_context(self.grid, self.params, self.numerics, n_tracers=self.n_tracers)
        self.lib = capi.load()
        self.Y = FieldVector(torch.from_numpy(np.ascontiguousarray(Yc)).to(self.device),
                             torch.from_numpy(np.ascontiguousarray(Yf)).to(self.device))
        self.t = 0.0
        self.set_implicit_precomputed_quantities(self.Y)

    # ------------------------------------------------------------------ plumbing
    @property
    def dt(self):
        return self.numerics.dt

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def to_device(self, Yc, Yf):
        t = self.torch
        return FieldVector(t.from_numpy(np.ascontiguousarray(Yc, dtype=self.FT)).to(self.device),
                           t.from_numpy(np.ascontiguousarray(Yf, dtype=self.FT)).to(self.device))

    def launch_count(self):
        return int(self.lib.b200_launch_count(self.ctx))

    def close(self):
        if self.ctx is not None:
            self.lib.b200_destroy(self.ctx)
            self.ctx = None

    # ------------------------------------------------------------------ hooks (integrator.jl:215-225)
    def set_implicit_precomputed_quantities(self, Y, t=0.0, precomputed=None):
        """cache_imp! (precomputed_quantities.jl:698-831). ``precomputed``: optional dict of output tensors
        with keys u_c, u3_f, K_c, T_c, p_c, h_tot_c."""
        cp = capi.CachePtrs()
        if precomputed:
            for k, v in precomputed.items():
                setattr(cp, k, v.data_ptr())
        capi.check(self.lib.b200_cache_imp(self.ctx, _p(Y.c), _p(Y.f), C.byref(cp), self._stream()), "b200_cache_imp", self.ctx)

    def remaining_tendency(self, Yt, Yt_lim, Y, t=0.0):
        """T_exp_T_lim! (remaining_tendency.jl:48-58)."""
        capi.check(self.lib.b200_t_exp_lim(self.ctx, _p(Yt.c), _p(Yt.f), _p(Yt_lim.c) if Yt_lim else None,
                                           _p(Yt_lim.f) if Yt_lim else None, _p(Y.c), _p(Y.f), float(t), self._stream()),
                   "b200_t_exp_lim", self.ctx)
        return Yt

    def remaining_tendency_phase(self, phase, Yt, Y):
        """Profiling aid: one phase of T_exp_T_lim! (0 pre-DSS kernel, 1 DSS of ∇² fields, 2 hyperdiffusion apply)."""
        capi.check(self.lib.b200_t_exp_phase(self.ctx, int(phase), _p(Yt.c), _p(Yt.f), _p(Y.c), _p(Y.f), self._stream()),
                   "b200_t_exp_phase", self.ctx)

    def remaining_tendency_phase_a(self, Yt, Y):
        self.remaining_tendency_phase(0, Yt, Y)

    def remaining_tendency_phase_c(self, Yt, Y):
        """Hyperdiffusion apply kernel alone (uses the ∇² fields left by the last phase 0 / phase 1 of this context)."""
        self.remaining_tendency_phase(2, Yt, Y)

    def implicit_stage(self, N, U, dtgamma):
        """Fused implicit stage (one Newton iteration of the CTS stage solve, integrator.jl:63-120): N ← U − J⁻¹R(U)
        plus the T_post_imp! correction, out of place."""
        capi.check(self.lib.b200_implicit_stage(self.ctx, _p(N.c), _p(N.f), _p(U.c), _p(U.f), float(dtgamma), self._stream()),
                   "b200_implicit_stage", self.ctx)

    def implicit_tendency(self, Yt, Y, t=0.0):
        """T_imp! (implicit_tendency.jl:36-98)."""
        capi.check(self.lib.b200_t_imp(self.ctx, _p(Yt.c), _p(Yt.f), _p(Y.c), _p(Y.f), float(t), self._stream()), "b200_t_imp", self.ctx)

    def update_jacobian(self, Y, dtgamma, t=0.0):
        """Wfact (jacobian.jl:74-75)."""
        capi.check(self.lib.b200_wfact(self.ctx, _p(Y.c), _p(Y.f), float(dtgamma), float(t), self._stream()), "b200_wfact", self.ctx)

    def jacobian_planes(self):
        """Debug/test aid: the coefficient planes stored by the last Wfact as a tensor [nh, 15, 16, nv + 1] (b200_debug_jacobian)."""
        nh, nf = int(self.Y.c.shape[0]), self.grid.nv + 1
        out = self.torch.empty((nh, 15, 16, nf), dtype=self.Y.c.dtype, device=self.device)
        capi.check(self.lib.b200_debug_jacobian(self.ctx, C.c_void_p(out.data_ptr()), out.numel() * out.element_size(), self._stream()),
                   "b200_debug_jacobian", self.ctx)
        return out

    def ldiv(self, dY, R):
        """ldiv!(ΔY, jacobian, R) (jacobian.jl:78-82)."""
        capi.check(self.lib.b200_ldiv(self.ctx, _p(dY.c), _p(dY.f), _p(R.c), _p(R.f), self._stream()), "b200_ldiv", self.ctx)

    def correct_implicit_advection_tendency(self, Yt, Y, t=0.0):
        """T_post_imp! (implicit_tendency.jl:322-339)."""
        capi.check(self.lib.b200_t_post_imp(self.ctx, _p(Yt.c), _p(Yt.f), _p(Y.c), _p(Y.f), float(t), self._stream()),
                   "b200_t_post_imp", self.ctx)

    def dss(self, Y, t=0.0):
        """dss! (constrain_state.jl:59-64): weighted DSS of Y.c (uₕ as a Covariant12 vector) and Y.f."""
        self.weighted_dss([(Y.c, int(Y.c.shape[1]), 0, 2), (Y.f, 1, 1, 0)])

    def weighted_dss(self, fields):
        """Spaces.weighted_dss!(pairs...): ``fields`` = [(tensor, ncomp, is_face, kind)]."""
        n = len(fields)
        ptrs = (C.c_void_p * n)(*[f[0].data_ptr() for f in fields])
        nf = (C.c_int32 * n)(*[f[1] for f in fields])
        isf = (C.c_int32 * n)(*[f[2] for f in fields])
        kind = (C.c_int32 * n)(*[f[3] for f in fields])
        capi.check(self.lib.b200_dss(self.ctx, ptrs, nf, isf, kind, n, self._stream()), "b200_dss", self.ctx)

    def constrain_state(self, Y, t=0.0):
        """constrain_state! (constrain_state.jl:34-39): no-op for dry / non-EDMF configurations."""

    def limiters_func(self, Y, t, ref_Y):
        """lim!(Y, p, t, ref_Y) (limited_tendencies.jl:64-122): SEM quasi-monotone limiter of the tracers of Y with bounds from
        ref_Y; the reference's no-op when no limiter is configured (defaults)."""
        capi.check(self.lib.b200_lim(self.ctx, _p(Y.c), _p(Y.f), _p(ref_Y.c), _p(ref_Y.f), float(t), self._stream()), "b200_lim", self.ctx)

    def initialize_implicit_stage_problem(self, Y, dtgamma):
        """initialize_imp! (initialize_implicit_problem.jl:33-57): no-op unless PrognosticEDMFX."""

    # ------------------------------------------------------------------ stepping (CTS.step!)
    def step(self, fused=True):
        capi.check(self.lib.b200_step_ars343(self.ctx, _p(self.Y.c), _p(self.Y.f), float(self.t), int(fused), self._stream()),
                   "b200_step_ars343", self.ctx)
        self.t += self.dt


def sypd(sim_seconds: float, wall_seconds: float) -> float:
    """Simulated years per day, 365-day year (src/simulation/solve.jl:40-45)."""
    return (sim_seconds / (365 * 86400.0)) / (wall_seconds / 86400.0)


def solve_atmos(sim: AtmosSimulation, n_steps: int, fused=True):
    """``solve_atmos!`` (solve.jl:119-158): one untimed step, then a device-timed solve; returns the
    numbers the reference logs (``sypd``, ``wall_time_per_timestep``)."""
    torch = sim.torch
    sim.step(fused)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_steps):
        sim.step(fused)
    e1.record()
    torch.cuda.synchronize()
    wall = e0.elapsed_time(e1) * 1e-3
    return dict(walltime=wall, wall_time_per_timestep=wall / n_steps, sypd=sypd(n_steps * sim.dt, wall))
