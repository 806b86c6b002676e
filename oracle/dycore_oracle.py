"""CPU ORACLE — test infrastructure only (NOT part of the product path).

NumPy restatement of the reference's dry dynamical-core time step.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module; the product package never does.

PARITY UNPINNED.  The arithmetic of this path lives in un-vendored Julia packages
(ClimaCore 0.15.1, ClimaTimeSteppers 0.10.6, Thermodynamics 1.3.0 — pins from
``.buildkite/Manifest-v1.11.toml``), none of which is under /root/reference, and Julia is not
installed here, so the reference cannot be executed and the repo holds no golden vectors for this
path (its hot-path unit tests are ``@test_skip`` placeholders).  What pins this
oracle instead: the structural identities the reference *does* test (tests/test_oracle_*.py):
χ≡1 consistency (test/prognostic_equations/tracer_mass_consistency_tests.jl:52-84),
impenetrability (advection_tests.jl:46-58), corrected+central == upwind
(correct_implicit_advection_tests.jl:40-67), sponge profiles
(test/parameterized_tendencies/sponge.jl:44-80), operator identities
(docs/src/discretization.md:76-94), a finite-difference check of the analytic Jacobian,
conservation to round-off, and convergence of the total tendency to zero at the design orders on the
analytic steady state of src/setups/DryBaroclinicWave.jl (the PDE itself as the judge).

Every function cites the reference file:line it restates.  Operators are applied *literally*, in
the order the reference composes them and with full per-point metric arrays (no factorisation),
so that the optimised CUDA kernels are checked against an independent formulation.

Array conventions: centre fields ``[h, j, i, v]`` (v = 0..Nv-1), face fields ``[h, j, i, f]``
(f = 0..Nv; face f is the lower face of centre f).  State ``Yc[h, 4, j, i, v]`` =
(ρ, uₕ₁, uₕ₂, ρe_tot), ``Yf[h, 1, j, i, f]`` = u₃ (all covariant components).
"""
from __future__ import annotations

import numpy as np


class Geom:
    """Per-point LocalGeometry pieces (ClimaCore ``Geometry.LocalGeometry`` [UPSTREAM-RECALL]):
    J, WJ, gⁱʲ, gᵢⱼ at one staggering, built as the product of the horizontal 2-D geometry and
    the vertical 1-D geometry with the deep-atmosphere scale factor ((R+z)/R)."""

    def __init__(self, grid, z, dz, FT):
        R = grid.radius
        s = (R + z) / R if grid.deep else np.ones_like(z)
        A = grid.dxdxi  # [h,j,i,a,b]
        G = np.einsum("...ab,...ac->...bc", A, A)  # covariant 2-D metric
        Ginv = np.linalg.inv(G)
        s2 = (s * s)[None, None, None, :]
        c = lambda a: np.ascontiguousarray(a, dtype=FT)
        self.J = c(grid.J2[..., None] * s2 * dz[None, None, None, :])
        self.WJ = c(grid.W[..., None] * grid.J2[..., None] * s2 * dz[None, None, None, :])
        self.g11 = c(Ginv[..., 0, 0, None] / s2)
        self.g12 = c(Ginv[..., 0, 1, None] / s2)
        self.g22 = c(Ginv[..., 1, 1, None] / s2)
        self.g33 = c(np.broadcast_to(1.0 / (dz * dz), self.J.shape))
        self.c11 = c(G[..., 0, 0, None] * s2)
        self.c12 = c(G[..., 0, 1, None] * s2)
        self.c22 = c(G[..., 1, 1, None] * s2)
        self.c33 = c(np.broadcast_to(dz * dz, self.J.shape))
        self.z = c(np.broadcast_to(z, self.J.shape))
        # local (east, north) physical components ↔ contravariant: u^a = (A_k⁻¹)·(u, v)
        Ainv = np.linalg.inv(A)
        self.Ainv = c(Ainv[..., None, :, :] / s[None, None, None, :, None, None])  # [h,j,i,v,a,b]
        # covariant ← physical: u_b = Σ_a (u, v)_a ∂x_a/∂ξ_b, with ∂x/∂ξ scaled by (R+z)/R in the deep shell
        self.A = c(A[..., None, :, :] * s[None, None, None, :, None, None])

    def slice(self, sl):
        """View of the geometry restricted to the element range ``sl`` (for the multi-threaded CPU arm)."""
        import copy

        g = copy.copy(self)
        for k, v in self.__dict__.items():
            if isinstance(v, np.ndarray):
                setattr(g, k, v[sl])
        return g


class Oracle:
    def __init__(self, grid, params, numerics, FT=np.float64):
        self.grid, self.P, self.N, self.FT = grid, params, numerics, FT
        g = grid
        self.nv = g.nv
        self.D = np.asarray(g.D, dtype=FT)
        self.c = Geom(g, g.z_c, g.dz_c, FT)
        self.f = Geom(g, g.z_f, g.dz_f, FT)
        from climaatmos_jl_b200.grid import dss_node_csr  # index tables are data, not arithmetic

        self.dss_offs, self.dss_mem = dss_node_csr(g.topology, g.nq)
        # cache.jl:179-183: Φ = grav·z ; ᶠgradᵥ_ᶜΦ
        self.Phi = np.asarray(params.grav * self.c.z, dtype=FT)
        self.gradv_Phi = self.gradv_c2f(self.Phi)
        # cache.jl:318-343 compute_coriolis (deep: 2Ω ẑ projected; shallow: 2Ω sin φ ŵ)
        Om = params.Omega
        lat = np.radians(g.lat)[..., None]
        fu, fv, fw = 0.0 * lat, 2 * Om * np.cos(lat), 2 * Om * np.sin(lat)
        self.f3_c = np.asarray(fw / g.dz_c[None, None, None, :], dtype=FT)  # CT3 = w / (∂z/∂ξ³)
        if g.deep:
            Ai = self.f.Ainv
            self.f1_f = np.asarray(Ai[..., 0, 0] * fu + Ai[..., 0, 1] * fv, dtype=FT)
            self.f2_f = np.asarray(Ai[..., 1, 0] * fu + Ai[..., 1, 1] * fv, dtype=FT)
        else:
            self.f1_f = self.f2_f = None
        h = g.node_horizontal_length_scale()
        # hyperdiffusion.jl:21-28
        self.nu4_vort = FT(numerics.nu4_vorticity_coeff * h**3)
        self.nu4_scalar = FT(self.nu4_vort / FT(numerics.prandtl_number))
        # DSS weights: WJ / Σ_collocated WJ (horizontal; docs/src/discretization.md:139-154)
        WJ2 = g.W * g.J2
        tot = WJ2.copy()
        for n in range(len(self.dss_offs) - 1):
            m = self.dss_mem[self.dss_offs[n] : self.dss_offs[n + 1]]
            ssum = sum(WJ2[e, j, i] for (e, i, j) in m)
            for e, i, j in m:
                tot[e, j, i] = ssum
        self.dss_w = np.asarray(WJ2 / tot, dtype=FT)
        self.lat_rad = np.radians(g.lat)[..., None]
        if self.moist and (getattr(numerics, "vert_diff", None) or numerics.held_suarez):
            raise ValueError("oracle: microphysics_model 0M is restated without vertical diffusion / Held–Suarez forcing")

    def slice(self, sl):
        """Shallow copy whose element-local arrays are views of the element range ``sl``; every tendency /
        Jacobian / solve routine is element-local, so running them on slices gives bitwise the same numbers.
        Used by ``step(..., pool=...)`` to run the CPU baseline on all host cores."""
        import copy

        o = copy.copy(self)
        o.c, o.f = self.c.slice(sl), self.f.slice(sl)
        for k in ("Phi", "gradv_Phi", "f3_c", "f1_f", "f2_f", "lat_rad"):
            v = getattr(self, k)
            setattr(o, k, None if v is None else v[sl])
        return o

    # ------------------------------------------------------------------ horizontal SEM (A.1)
    def dx(self, a):  # Σ_k D[i,k] a[h,j,k,v]
        return np.einsum("ik,hjkv->hjiv", self.D, a)

    def dy(self, a):
        return np.einsum("jk,hkiv->hjiv", self.D, a)

    def dxT(self, a):  # Σ_k D[k,i] a[h,j,k,v]
        return np.einsum("ki,hjkv->hjiv", self.D, a)

    def dyT(self, a):
        return np.einsum("kj,hkiv->hjiv", self.D, a)

    def grad(self, a):
        return self.dx(a), self.dy(a)

    def wgrad(self, a, G):
        W = G.WJ / G.J
        return -self.dxT(W * a) / W, -self.dyT(W * a) / W

    def div(self, u1c, u2c, G):  # contravariant components in
        return (self.dx(G.J * u1c) + self.dy(G.J * u2c)) / G.J

    def wdiv(self, u1c, u2c, G):
        return -(self.dxT(G.WJ * u1c) + self.dyT(G.WJ * u2c)) / G.WJ

    def curl3(self, u1, u2, G):  # covariant (u1,u2) → contravariant 3
        return (self.dx(u2) - self.dy(u1)) / G.J

    def curl12(self, u3, G):  # covariant u3 → contravariant (1,2)
        return self.dy(u3) / G.J, -self.dx(u3) / G.J

    def wcurl3(self, u1, u2, G):
        W = G.WJ / G.J
        return (self.dyT(W * u1) - self.dxT(W * u2)) / G.WJ

    def wcurl12(self, u3, G):
        W = G.WJ / G.J
        return -self.dyT(W * u3) / G.WJ, self.dxT(W * u3) / G.WJ

    @staticmethod
    def ct12(u1, u2, G):  # CT12(C12): u^a = g^{ab} u_b (flat: no vertical coupling)
        return G.g11 * u1 + G.g12 * u2, G.g12 * u1 + G.g22 * u2

    @staticmethod
    def c12(v1, v2, G):
        return G.c11 * v1 + G.c12 * v2, G.c12 * v1 + G.c22 * v2

    def split_div(self, F1, F2, psi, G):
        """docs/src/discretization.md:245-253: ½ wdiv(Fψ) + ½[ψ wdiv(F) + F·grad ψ]."""
        gx, gy = self.grad(psi)
        half = self.FT(0.5)
        return half * self.wdiv(F1 * psi, F2 * psi, G) + half * (psi * self.wdiv(F1, F2, G) + F1 * gx + F2 * gy)

    # ------------------------------------------------------------------ vertical FD (A.2)
    def interp_f2c(self, a):
        return self.FT(0.5) * (a[..., :-1] + a[..., 1:])

    def interp_c2f(self, a):  # Extrapolate BCs (abbreviations.jl:173-176)
        out = np.empty(a.shape[:-1] + (a.shape[-1] + 1,), dtype=a.dtype)
        out[..., 1:-1] = self.FT(0.5) * (a[..., :-1] + a[..., 1:])
        out[..., 0] = a[..., 0]
        out[..., -1] = a[..., -1]
        return out

    def winterp_c2f(self, w, a):  # abbreviations.jl:184-187 ; discretization.md:171
        return self.interp_c2f(w * a) / self.interp_c2f(w)

    def gradv_c2f(self, a):  # SetGradient(0) BCs (abbreviations.jl:206-209)
        out = np.zeros(a.shape[:-1] + (a.shape[-1] + 1,), dtype=a.dtype)
        out[..., 1:-1] = a[..., 1:] - a[..., :-1]
        return out

    def advdiv_f2c(self, u3c):  # SetValue(0) BCs (abbreviations.jl:106-109); CT3 in
        Ju = self.f.J * u3c
        Ju = Ju.copy()
        Ju[..., 0] = 0
        Ju[..., -1] = 0
        return (Ju[..., 1:] - Ju[..., :-1]) / self.c.J

    def curlv_c2f(self, u1, u2):  # SetCurl(0) BCs (abbreviations.jl:216-219) → CT12
        o1 = np.zeros(u1.shape[:-1] + (u1.shape[-1] + 1,), dtype=u1.dtype)
        o2 = np.zeros_like(o1)
        o1[..., 1:-1] = -(u2[..., 1:] - u2[..., :-1]) / self.f.J[..., 1:-1]
        o2[..., 1:-1] = (u1[..., 1:] - u1[..., :-1]) / self.f.J[..., 1:-1]
        return o1, o2

    def upwind1(self, v, a):  # UpwindBiasedProductC2F; boundary faces one-sided (flux zeroed later)
        lo = np.concatenate([a[..., :1], a], -1)
        hi = np.concatenate([a, a[..., -1:]], -1)
        return v * np.where(v >= 0, lo, hi)

    def lin_vanleer(self, v, a, dt):
        """LinVanLeerC2F, MonotoneLocalExtrema constraint, FirstOrderOneSided BCs
        (abbreviations.jl:250-256; Lin 1994) [UPSTREAM-RECALL]: interior faces 2..Nv-2 use the
        limited slope of the upwind cell with the Courant correction (1 ∓ v³dt); the two faces
        next to each boundary fall back to first-order upwinding."""
        FT = self.FT
        out = self.upwind1(v, a)
        nv = a.shape[-1]
        if nv < 4:
            return out

        def slope(am, a0, ap):
            d = ((a0 - am) + (ap - a0)) / FT(2)
            mn = np.minimum(np.minimum(am, a0), ap)
            mx = np.maximum(np.maximum(am, a0), ap)
            lim = np.minimum(np.abs(d), np.minimum(FT(2) * (a0 - mn), FT(2) * (mx - a0)))
            return np.sign(d) * lim

        # face f (2..nv-2): a⁻⁻=a[f-2], a⁻=a[f-1], a⁺=a[f], a⁺⁺=a[f+1]
        amm, am, ap, app = a[..., :-3], a[..., 1:-2], a[..., 2:-1], a[..., 3:]
        vf = v[..., 2:-2]
        pos = am + slope(amm, am, ap) / FT(2) * (FT(1) - vf * FT(dt))
        neg = ap - slope(am, ap, app) / FT(2) * (FT(1) + vf * FT(dt))
        out[..., 2:-2] = vf * np.where(vf >= 0, pos, neg)
        return out

    def upwind3(self, v, a):
        """ᶠupwind3 = Upwind3rdOrderBiasedProductC2F(bottom = ThirdOrderOneSided(), top = ThirdOrderOneSided())
        (abbreviations.jl:229-240) [UPSTREAM-RECALL ClimaCore 0.15.1 finite_difference.jl]: interior faces (four-point stencil available)
            (v·(7(a⁺ + a⁻) − (a⁺⁺ + a⁻⁻)) − |v|·(3(a⁺ − a⁻) − (a⁺⁺ − a⁻⁻))) / 12,
        i.e. the upwind-biased cubic-accurate face value (−a⁻⁻ + 5a⁻ + 2a⁺)/6 for v > 0 and its mirror image for v < 0; the first
        interior face uses the one-sided right-biased reconstruction v·(4a⁻ + 10a⁺ − 2a⁺⁺)/12 and the last one the left-biased
        v·(−2a⁻⁻ + 10a⁻ + 4a⁺)/12 whatever the sign of v (boundary_width 2); the two boundary faces carry no flux (ᶜadvdivᵥ SetValue(0)).
        Columns with fewer than three levels fall back to first-order upwinding."""
        FT = self.FT
        out = self.upwind1(v, a)
        nv = a.shape[-1]
        if nv < 3:
            return out
        out[..., 1] = v[..., 1] * (FT(4) * a[..., 0] + FT(10) * a[..., 1] - FT(2) * a[..., 2]) / FT(12)
        out[..., nv - 1] = v[..., nv - 1] * (-FT(2) * a[..., nv - 3] + FT(10) * a[..., nv - 2] + FT(4) * a[..., nv - 1]) / FT(12)
        if nv >= 4:
            amm, am, ap, app = a[..., :-3], a[..., 1:-2], a[..., 2:-1], a[..., 3:]
            vf = v[..., 2:-2]
            out[..., 2:-2] = (vf * (FT(7) * (ap + am) - (app + amm)) - np.abs(vf) * (FT(3) * (ap - am) - (app - amm))) / FT(12)
        return out

    # ------------------------------------------------------------------ thermodynamics (A.4)
    def exner(self, p):
        return (p / self.FT(self.P.p_ref_theta)) ** self.FT(self.P.kappa_d)

    def T_ref(self, p):  # refstate_thermodynamics.jl air_temperature_reference
        P = self.P
        return self.FT(P.T_min_ref) + self.FT(P.T_surf_ref - P.T_min_ref) * self.exner(p) ** 7

    def theta_vr(self, p):
        return self.T_ref(p) / self.exner(p)

    def phi_r(self, p):
        P, FT = self.P, self.FT
        Pi = self.exner(p)
        return -FT(P.cp_d) * (FT(P.T_min_ref) * np.log(Pi) + FT(P.T_surf_ref - P.T_min_ref) / FT(7) * (Pi**7 - FT(1)))

    def sd_r(self, p):
        P, FT = self.P, self.FT
        return FT(P.cp_d) * (self.T_ref(p) - FT(P.T_0)) + self.phi_r(p)

    # ------------------------------------------------------------------ moist thermodynamics (EquilibriumMicrophysics0M)
    # Thermodynamics.jl 1.3.0 is not vendored; these restate its PUBLISHED formulation as documented in
    # docs/src/thermodynamics.md:60-150 (calorically perfect constituents, Romps 2008 energy references, Rankine–Kirchhoff
    # saturation vapour pressure, Pressel 2015 liquid-fraction-weighted latent heat, Kaul 2015 supercooled-liquid ramp)
    # and are anchored on the reference's call sites (precomputed_quantities.jl:735-815, manual_sparse_jacobian.jl:653-690).
    @property
    def moist(self):
        return getattr(self.N, "microphysics_model", None) == "0M"

    @property
    def q0(self):
        """Index of the first PASSIVE tracer in Y.c (prognostic_variables.jl:54-61: ρ, uₕ, ρe_tot, ρq_tot, chemistry…)."""
        return 5 if self.moist else 4

    def gas_constant_air(self, qt, ql, qi):  # R_m = R_d (1 − q_t) + R_v q_v   (thermodynamics.md:62-66)
        P, FT = self.P, self.FT
        return FT(P.R_d) * (FT(1) - qt) + FT(P.R_v) * (qt - ql - qi)

    def cv_m(self, qt, ql, qi):  # thermodynamics.md:78-81
        P, FT = self.P, self.FT
        return FT(P.cv_d) + FT(P.cv_v - P.cv_d) * qt + FT(P.cp_l - P.cv_v) * ql + FT(P.cp_i - P.cv_v) * qi

    def internal_energy(self, T, qt, ql, qi):  # thermodynamics.md:103-111
        P, FT = self.P, self.FT
        return (self.cv_m(qt, ql, qi) * (T - FT(P.T_0)) + (qt - ql - qi) * FT(P.e_int_v0) - qi * FT(P.e_int_i0)
                - (FT(1) - qt) * FT(P.R_d * P.T_0))

    def air_temperature(self, e_int, qt, ql, qi):  # the inversion of internal_energy
        P, FT = self.P, self.FT
        return FT(P.T_0) + (e_int - (qt - ql - qi) * FT(P.e_int_v0) + qi * FT(P.e_int_i0) + (FT(1) - qt) * FT(P.R_d * P.T_0)) / self.cv_m(qt, ql, qi)

    def liquid_fraction(self, T):
        """Supercooled-liquid ramp between T_icenuc and T_freeze (thermodynamics.md:158-164); returns (λ, dλ/dT)."""
        P, FT = self.P, self.FT
        w = FT(P.T_freeze - P.T_icenuc)
        x = np.clip((T - FT(P.T_icenuc)) / w, FT(0), FT(1))
        n = FT(P.pow_icenuc)
        inside = (T > FT(P.T_icenuc)) & (T < FT(P.T_freeze))
        lam = np.where(T >= FT(P.T_freeze), FT(1), np.where(T <= FT(P.T_icenuc), FT(0), x**n))
        dlam = np.where(inside, n * np.where(inside, x, FT(1)) ** (n - FT(1)) / w, FT(0))
        return lam.astype(FT), dlam.astype(FT)

    def _ln_pvs(self, T, lam):
        """ln of the Rankine–Kirchhoff saturation vapour pressure with the λ-weighted latent heat (thermodynamics.md:129-139):
        p_vs = p_tr (T/T_tr)^(Δcp/R_v) exp((L_0 − Δcp T_0)/R_v (1/T_tr − 1/T)); λ = 1 over liquid, 0 over ice."""
        P, FT = self.P, self.FT
        dcp = lam * FT(P.cp_v - P.cp_l) + (FT(1) - lam) * FT(P.cp_v - P.cp_i)
        L0 = lam * FT(P.LH_v0) + (FT(1) - lam) * FT(P.LH_s0)
        return (np.log(FT(P.press_triple)) + dcp / FT(P.R_v) * np.log(T / FT(P.T_triple))
                + (L0 - dcp * FT(P.T_0)) / FT(P.R_v) * (FT(1) / FT(P.T_triple) - FT(1) / T)), dcp, L0

    def q_vap_saturation(self, T, rho, lam):
        lnp, _, _ = self._ln_pvs(T, lam)
        return np.exp(lnp) / (rho * self.FT(self.P.R_v) * T)

    def saturation_adjustment(self, rho, e_int, qt, maxiter=40):
        """TD.saturation_adjustment(thermo_params, TD.ρe(), ρ, e_int, q_tot) → (T, q_liq, q_ice)
        (precomputed_quantities.jl:639-642): unsaturated air keeps T(e_int, q_tot); otherwise T solves
        e_int = I(T, q_tot, λ q_c, (1 − λ) q_c) with q_c = max(0, q_tot − q_vs(T, ρ)) (equilibrium partition), by Newton's method
        with the analytic derivative from the all-vapour temperature, iterated to round-off (the reference stops at a relative
        temperature tolerance of 1e-4 after a Newton step [UPSTREAM-RECALL], i.e. at an error of the order of the tolerance
        squared).  Safeguard for strongly supersaturated input only: an iterate beyond the dew point (q_c = 0, where the residual
        is linear and Newton would return to the starting point) is pulled back half-way to the last iterate with a negative
        residual; ordinary atmospheric states never take that branch."""
        P, FT = self.P, self.FT
        z = np.zeros_like(qt)
        T1 = self.air_temperature(e_int, qt, z, z)
        lam, dlam = self.liquid_fraction(T1)
        sat = qt > self.q_vap_saturation(T1, rho, lam)
        T = T1.copy()
        if np.any(sat):
            Ts, r, e, q = T1[sat], rho[sat], e_int[sat], qt[sat]
            Tlo = Ts.copy()
            tol = FT(8) * np.finfo(FT).eps
            for _ in range(maxiter):
                lam, dlam = self.liquid_fraction(Ts)
                lnp, dcp, L0 = self._ln_pvs(Ts, lam)
                qvs = np.exp(lnp) / (r * FT(P.R_v) * Ts)
                dlnp = (L0 + dcp * (Ts - FT(P.T_0))) / (FT(P.R_v) * Ts * Ts) + dlam * (
                    FT(P.cp_i - P.cp_l) / FT(P.R_v) * np.log(Ts / FT(P.T_triple))
                    + (FT(-P.LH_f0) - FT(P.cp_i - P.cp_l) * FT(P.T_0)) / FT(P.R_v) * (FT(1) / FT(P.T_triple) - FT(1) / Ts))
                on = q > qvs
                qc = np.where(on, q - qvs, FT(0))
                dqc = -qvs * (dlnp - FT(1) / Ts)
                ql, qi = lam * qc, (FT(1) - lam) * qc
                dql, dqi = dlam * qc + lam * dqc, -dlam * qc + (FT(1) - lam) * dqc
                f = self.internal_energy(Ts, q, ql, qi) - e
                df = (self.cv_m(q, ql, qi) + (Ts - FT(P.T_0)) * (FT(P.cp_l - P.cv_v) * dql + FT(P.cp_i - P.cv_v) * dqi)
                      - dqc * FT(P.e_int_v0) - dqi * FT(P.e_int_i0))
                Tlo = np.where(on & (f < 0), Ts, Tlo)
                Tn = np.where(on, Ts - f / df, FT(0.5) * (Tlo + Ts))
                done = on & (np.abs(Tn - Ts) <= tol * Ts)
                Ts = Tn
                if np.all(done):
                    break
            T[sat] = Ts
        lam, _ = self.liquid_fraction(T)
        qc = np.where(sat, np.maximum(FT(0), qt - self.q_vap_saturation(T, rho, lam)), FT(0))
        return T, (lam * qc).astype(FT), ((FT(1) - lam) * qc).astype(FT)

    def q_tot_r(self, p):
        """refstate_thermodynamics.jl:140-150: RH_ref · q_sat(T_r(p), ρ_r(p)) over liquid, zero above 250 hPa."""
        FT = self.FT
        T_r = self.T_ref(p)
        rho_r = p / (FT(self.P.R_d) * T_r)
        return np.where(p < FT(25000), FT(0), FT(0.5) * self.q_vap_saturation(T_r, rho_r, np.ones_like(p)))

    def h_eff_plus_Phi(self, T, qt, ql, qi):
        """ᶜh_eff_plus_Φ! (eddy_diffusion_closures.jl:970-983) with ᶜsuspended_water (:997-1007, equilibrium branch):
        (h_v q_v + h_l q_l + h_i q_i)/max(q_v + q_l + q_i, ε) + Φ; h_v = cp_v (T − T_0) + L_v0, h_l = cp_l (T − T_0),
        h_i = cp_i (T − T_0) − L_f0 (thermodynamics.md:117-121)."""
        P, FT = self.P, self.FT
        qv, ql, qi = np.maximum(FT(0), qt - ql - qi), np.maximum(FT(0), ql), np.maximum(FT(0), qi)
        dT = T - FT(P.T_0)
        num = (FT(P.cp_v) * dT + FT(P.LH_v0)) * qv + (FT(P.cp_l) * dT) * ql + (FT(P.cp_i) * dT - FT(P.LH_f0)) * qi
        return num / np.maximum(qv + ql + qi, np.finfo(FT).eps) + self.Phi

    # ------------------------------------------------------------------ sponges
    def beta_rayleigh(self, z, alpha):  # rayleigh_sponge.jl:7-28
        P = self.P
        zeta = np.sin(np.pi * ((z - P.zd_rayleigh) / (self.grid.z_max - P.zd_rayleigh) / 2)) ** 2
        return np.asarray(np.where(z > P.zd_rayleigh, alpha, 0.0) * zeta, dtype=self.FT)

    def beta_viscous(self, z):  # viscous_sponge.jl:9-26
        P = self.P
        zeta = np.sin(np.pi * ((z - P.zd_viscous) / (self.grid.z_max - P.zd_viscous) / 2)) ** 2
        return np.asarray(np.where(z > P.zd_viscous, P.kappa_2_sponge, 0.0) * zeta, dtype=self.FT)

    # ------------------------------------------------------------------ cache_imp!
    def set_implicit_precomputed_quantities(self, Yc, Yf):
        """precomputed_quantities.jl:698-831 (dry branch). Mutates Yf at the boundary faces
        (set_velocity_at_surface!/top!, :486-554) and returns the precomputed dict."""
        FT, P, c, f = self.FT, self.P, self.c, self.f
        rho, u1, u2, rhoe = Yc[:, 0], Yc[:, 1], Yc[:, 2], Yc[:, 3]
        u3 = Yf[:, 0]
        # :707 ᶠuₕ³ = ᶠwinterp(ρJ, CT3(uₕ)); flat grid: g³ʰ = 0 → CT3(uₕ) = 0
        uh3 = self.winterp_c2f(rho * c.J, np.zeros_like(rho))
        u3[..., 0] = -uh3[..., 0] / f.g33[..., 0]
        u3[..., -1] = -uh3[..., -1] / f.g33[..., -1]
        # :567-572 set_velocity_quantities!
        u3c = self.interp_f2c(u3)  # covariant, interpolated component-wise
        fu3 = uh3 + f.g33 * u3  # ᶠu³
        # utilities.jl:194-209 compute_kinetic
        c1, c2 = self.ct12(u1, u2, c)
        K = FT(0.5) * ((u1 * c1 + u2 * c2) + self.interp_f2c(u3 * (f.g33 * u3)) + FT(2) * (FT(0) * u3c))
        e_int = rhoe / rho - K - self.Phi
        if self.moist:  # :735-747 EquilibriumMicrophysics0M (no T floor in this branch), :794-815
            qt = np.maximum(FT(0), Yc[:, 4] / rho)
            T, ql, qi = self.saturation_adjustment(rho, e_int, qt)
            Rm = self.gas_constant_air(qt, ql, qi)
            h_tot = rhoe / rho + Rm * T  # TD.total_enthalpy = e_tot + R_m T
            p = rho * Rm * T
            return dict(u1=u1, u2=u2, u3c=u3c, fu3=fu3, K=K, T=T, p=p, h_tot=h_tot, qt=qt, ql=ql, qi=qi, Rm=Rm, cvm=self.cv_m(qt, ql, qi))
        # Thermodynamics.jl air_temperature with the dry-air reference internal energy −R_d·T_0
        # (docs/src/thermodynamics.md:103-111): e_int = cv_d (T − T_0) − R_d T_0
        T = np.maximum(FT(P.T_min_sgs), FT(P.T_0) + (e_int + FT(P.R_d) * FT(P.T_0)) / FT(P.cv_d))
        h_tot = rhoe / rho + FT(P.R_d) * T
        p = rho * FT(P.R_d) * T
        return dict(u1=u1, u2=u2, u3c=u3c, fu3=fu3, K=K, T=T, p=p, h_tot=h_tot)

    # ------------------------------------------------------------------ T_imp!
    def vertical_transport(self, rho, fu3, chi, dt, upwinding):
        """implicit_tendency.jl:120-143."""
        rJf = self.interp_c2f(rho * self.c.J) / self.f.J
        if upwinding == "none":
            return -self.advdiv_f2c(rJf * fu3 * self.interp_c2f(chi))
        if upwinding == "first_order":
            return -self.advdiv_f2c(rJf * self.upwind1(fu3, chi))
        if upwinding == "vanleer_limiter":
            return -self.advdiv_f2c(rJf * self.lin_vanleer(fu3, chi, dt))
        if upwinding == "third_order":
            return -self.advdiv_f2c(rJf * self.upwind3(fu3, chi))
        raise ValueError(upwinding)

    def theta_v(self, T, p, pc=None):
        """refstate_thermodynamics.jl:56-61: θ_v = T R_m / (Π R_d)."""
        if pc is not None and "Rm" in pc:
            return T * pc["Rm"] / (self.exner(p) * self.FT(self.P.R_d))
        return T / self.exner(p)

    def implicit_tendency(self, Yc, Yf, pc):
        """implicit_tendency.jl:36-98 → implicit_vertical_advection_tendency! :185-298 (dry)."""
        FT, P = self.FT, self.P
        Ytc, Ytf = np.zeros_like(Yc), np.zeros_like(Yf)
        rho = Yc[:, 0]
        Ytc[:, 0] -= self.advdiv_f2c(self.interp_c2f(rho * self.c.J) / self.f.J * pc["fu3"])
        Ytc[:, 3] += self.vertical_transport(rho, pc["fu3"], pc["h_tot"], self.N.dt, "none")
        if self.moist:  # :210-214 central transport of the active tracer ρq_tot (q_tot = specific(ρq_tot, ρ), not clipped)
            Ytc[:, 4] += self.vertical_transport(rho, pc["fu3"], Yc[:, 4] / rho, self.N.dt, "none")
        p, T = pc["p"], pc["T"]
        dth = self.theta_v(T, p, pc) - self.theta_vr(p)
        Ytf[:, 0] -= self.gradv_Phi - self.gradv_c2f(self.phi_r(p)) + FT(P.cp_d) * self.interp_c2f(dth) * self.gradv_c2f(self.exner(p))
        if self.N.rayleigh_sponge:
            Ytf[:, 0] += -self.beta_rayleigh(self.f.z, P.alpha_rayleigh_w) * Yf[:, 0]
        if self.vert_diff and self.implicit_diffusion:  # implicit_tendency.jl:69-78 (diff_mode == Implicit())
            self.vertical_diffusion_boundary_layer_tendency(Ytc, Yc, pc)
        return Ytc, Ytf

    # ------------------------------------------------------------------ vertical diffusion
    @property
    def vert_diff(self):
        return getattr(self.N, "vert_diff", None)

    @property
    def implicit_diffusion(self):
        return bool(getattr(self.N, "implicit_diffusion", False))

    def eddy_diffusivity(self, Yc, pc):
        """ᶜcompute_eddy_diffusivity_coefficient (src/cache/eddy_diffusivity_coefficient.jl:16-42) with
        eddy_diffusivity_coefficient_H / eddy_diffusivity_coefficient (src/cache/precomputed_quantities.jl:652-676).
        K_u = K_h (vertical_diffusion_boundary_layer.jl:18, manual_sparse_jacobian.jl:984-992)."""
        FT, P, c = self.FT, self.P, self.c
        if self.vert_diff == "DecayWithHeightDiffusion":
            z_sfc = self.f.z[..., :1]
            return np.asarray(FT(P.D_0_diffusion) * np.exp(-(c.z - z_sfc) / FT(P.H_diffusion)), dtype=FT)
        if self.vert_diff == "VerticalDiffusion":
            u1, u2 = Yc[:, 1, ..., :1], Yc[:, 2, ..., :1]  # Fields.level(ᶜuₕ, 1)
            g11, g12, g22 = c.g11[..., :1], c.g12[..., :1], c.g22[..., :1]
            norm = np.sqrt(u1 * (g11 * u1 + g12 * u2) + u2 * (g12 * u1 + g22 * u2))  # LinearAlgebra.norm of a Covariant12Vector
            z_a = FT(self.grid.dz_c[0]) / FT(2)  # Fields.Δz_field(level 1) / 2
            K_E = FT(P.C_E) * norm * z_a
            p = pc["p"]
            p_pbl, p_strato = FT(85000), FT(10000)
            return np.asarray(np.where(p > p_pbl, K_E, K_E * np.exp(-(((p_pbl - p) / p_strato) ** 2))), dtype=FT)
        raise ValueError(f"vert_diff = {self.vert_diff!r}")

    def _rhoK_face(self, Yc, pc):
        """ᶠρK = ᶠinterp(ρ) / ᶠinterp(1 / max(K_h, ε)) — harmonic-mean face diffusivity
        (vertical_diffusion_boundary_layer.jl:86-90; manual_sparse_jacobian.jl:1078-1082)."""
        eps = np.finfo(self.FT).eps
        K = self.eddy_diffusivity(Yc, pc)
        return self.interp_c2f(Yc[:, 0]) / self.interp_c2f(self.FT(1) / np.maximum(K, eps))

    def diffdiv_f2c(self, F3):
        """ᶜdiffdivᵥ (abbreviations.jl:124-135) of a covariant-3 face flux: SetValue(C3(0)) at both boundaries;
        the divergence acts on the contravariant component g³³ F₃."""
        return self.advdiv_f2c(self.f.g33 * F3)

    def vertical_diffusion_boundary_layer_tendency(self, Ytc, Yc, pc):
        """vertical_diffusion_boundary_layer_tendency! (src/prognostic_equations/vertical_diffusion_boundary_layer.jl:64-154),
        dry branch + passive tracers; ADDS into Ytc.  Called from additional_tendency! (remaining_tendency.jl:185-195) when
        diffusion is explicit, from implicit_tendency! otherwise."""
        FT, P, c, f = self.FT, self.P, self.c, self.f
        rho = Yc[:, 0]
        rhoK = self._rhoK_face(Yc, pc)
        if not getattr(self.N, "disable_momentum_vertical_diffusion", False):
            # :91-96  uₕₜ −= C12(ᶜdivᵥ(−2 ᶠρK ᶠstrain_rate) / ρ), strain rate of UVW(ᶜu) (utilities.jl:251-262):
            # ε = (G + Gᵀ)/2 with G[w, b] = ∂(u, v, w)_b/∂z on interior faces, zero on the boundary faces (SetGradient(0)).
            # The vertical divergence contracts the w index: (divᵥ τ)_b = (1/J) δ(J τ_wb / (∂z/∂ξ³)), and ε_wb = ½ ∂(u,v)_b/∂z for
            # the two horizontal components, which are all C12 keeps on a flat grid.
            A = c.A
            Ai = c.Ainv
            up = [Ai[..., 0, a] * Yc[:, 1] + Ai[..., 1, a] * Yc[:, 2] for a in range(2)]  # physical (u, v) = A⁻ᵀ uₕ
            dzf = np.asarray(np.broadcast_to(self.grid.dz_f, f.J.shape), dtype=FT)
            dv = []
            for a in range(2):
                strain = FT(0.5) * self.gradv_c2f(up[a]) / dzf  # ε_wa on faces (W component: covariant / (∂z/∂ξ³))
                tau = -FT(2) * rhoK * strain
                Jt = f.J * (tau / dzf)  # contravariant-3 part of the w index
                dv.append((Jt[..., 1:] - Jt[..., :-1]) / c.J / rho)  # ᶜdivᵥ has no BCs; the boundary strain is already 0
            for b in range(2):
                Ytc[:, 1 + b] -= A[..., 0, b] * dv[0] + A[..., 1, b] * dv[1]
        # :101-102  ρe_totₜ −= ᶜdiffdivᵥ(−(ᶠρK ᶠgradᵥ(dry_static_energy(T, Φ))))
        s_d = FT(P.cp_d) * (pc["T"] - FT(P.T_0)) + self.Phi
        Ytc[:, 3] -= self.diffdiv_f2c(-(rhoK * self.gradv_c2f(s_d)))
        # :150-153 passive grid-scale tracers
        for q in range(4, Yc.shape[1]):
            Ytc[:, q] -= self.diffdiv_f2c(-(rhoK * self.gradv_c2f(Yc[:, q] / rho)))

    def update_diffusion_jacobian(self, Jm, Yc, pc, dtg):
        """update_diffusion_jacobian! (manual_sparse_jacobian.jl:1031-1261), dry non-EDMF branch.  Tridiagonal centre→centre
        blocks stored as (lo, d, hi) with lo[..., 0] = hi[..., -1] = 0.  ᶜdiffusion_h_matrix = ᶜadvdivᵥ_matrix ⋅ Diag(ᶠρK) ⋅
        ᶠgradᵥ_matrix (:1078-1084); K_u = K_h so ᶜdiffusion_u_matrix is the same matrix (:1096-1100)."""
        FT, P, c, f = self.FT, self.P, self.c, self.f
        dtg = FT(dtg)
        rho = Yc[:, 0]
        w = f.J * f.g33 * self._rhoK_face(Yc, pc)
        w[..., 0] = 0
        w[..., -1] = 0
        lo, hi = w[..., :-1] / c.J, w[..., 1:] / c.J  # columns k-1 and k+1 of row k
        d = -(lo + hi)
        z = np.zeros_like(rho[..., :1])
        left = lambda a: np.concatenate([z, a[..., :-1]], -1)  # value at k-1
        right = lambda a: np.concatenate([a[..., 1:], z], -1)  # value at k+1

        def scaled_cols(sc):  # dtγ · D · Diag(sc) − I
            return (dtg * lo * left(sc), dtg * d * sc - FT(1), dtg * hi * right(sc))

        out = dict(rhoe_rhoe=scaled_cols(FT(P.cp_d) / (FT(P.cv_d) * rho)))  # :1129-1131 (cv_m = cv_d when dry)
        out["tracer"] = scaled_cols(FT(1) / rho) if Yc.shape[1] > 4 else None  # :1190-1195
        if not getattr(self.N, "disable_momentum_vertical_diffusion", False):  # :1252-1258  dtγ Diag(1/ρ) ⋅ D_u − I
            out["uh_uh"] = (dtg * lo / rho, dtg * d / rho - FT(1), dtg * hi / rho)
        else:
            out["uh_uh"] = None
        Jm["diff"] = out
        return Jm

    @staticmethod
    def _tri_solve(tri, rhs):
        """Thomas algorithm along the last axis (ClimaCore single_field_solver.jl tridiagonal case [UPSTREAM-RECALL])."""
        l, d, u = tri
        n = rhs.shape[-1]
        cp, dp = np.zeros_like(rhs), np.zeros_like(rhs)
        cp[..., 0] = u[..., 0] / d[..., 0]
        dp[..., 0] = rhs[..., 0] / d[..., 0]
        for i in range(1, n):
            den = d[..., i] - l[..., i] * cp[..., i - 1]
            cp[..., i] = u[..., i] / den
            dp[..., i] = (rhs[..., i] - l[..., i] * dp[..., i - 1]) / den
        x = np.zeros_like(rhs)
        x[..., -1] = dp[..., -1]
        for i in range(n - 2, -1, -1):
            x[..., i] = dp[..., i] - cp[..., i] * x[..., i + 1]
        return x

    @staticmethod
    def _tri_mul(tri, x):
        l, d, u = tri
        y = d * x
        y[..., 1:] += l[..., 1:] * x[..., :-1]
        y[..., :-1] += u[..., :-1] * x[..., 1:]
        return y

    def ldiv_iterative(self, Jm, Rc, Rf):
        """ldiv! with implicit diffusion: ApproximateBlockArrowheadIterativeSolve(ρ, ρe_tot; alg₁ = BlockLowerTriangularSolve(ρ),
        alg₂ = BlockLowerTriangularSolve(uₕ), P_alg₁ = MainDiagonalPreconditioner(), n_iters)
        (manual_sparse_jacobian.jl:538-578) [UPSTREAM-RECALL ClimaCore 0.15.1 MatrixFields/field_matrix_solver.jl]:
        a SchurComplementReductionSolve — x₂ solves (A₂₂ − A₂₁A₁₁⁻¹A₁₂) x₂ = b₂ − A₂₁A₁₁⁻¹b₁ by a StationaryIterativeSolve
        x ← x + P⁻¹(b − S x) started from x = P⁻¹ b, with P = A₂₂ − A₂₁ diag(A₁₁)⁻¹ A₁₂ solved block-lower-triangularly
        (uₕ first); then x₁ = A₁₁⁻¹(b₁ − A₁₂x₂).  names₁ = (ρ, ρe_tot): A_ρρ = −I, A_ρe,ρ = 0, A_ρe,ρe tridiagonal;
        names₂ = (passive tracers, uₕ, u₃): passive tracers and uₕ do not couple to names₁ on a flat grid, so their
        tridiagonal solves are exact at x[0]."""
        FT = self.FT
        D = Jm["diff"]
        n_iters = int(getattr(self.N, "approximate_linear_solve_iters", 1))
        dYc, dYf = np.zeros_like(Rc), np.zeros_like(Rf)
        Rrho, R1, R2, Rre, R3 = Rc[:, 0], Rc[:, 1], Rc[:, 2], Rc[:, 3], Rf[:, 0]
        cl = lambda a: np.concatenate([a[..., :1] * 0, a], -1)
        ch = lambda a: np.concatenate([a, a[..., -1:] * 0], -1)
        f2 = lambda a21, r: a21[0] * cl(r) + a21[1] * ch(r)  # face-row bidiagonal block · centre vector
        c2 = lambda a12, x: a12[0] * x[..., :-1] + a12[1] * x[..., 1:]  # centre-row bidiagonal block · face vector
        Aee = D["rhoe_rhoe"]
        # uₕ and passive tracers: (uₕ,uₕ) / (ρχ,ρχ) tridiagonal or the −I fallback
        if D["uh_uh"] is not None:
            x1, x2 = self._tri_solve(D["uh_uh"], R1), self._tri_solve(D["uh_uh"], R2)
        else:
            x1, x2 = -R1, -R2
        dYc[:, 1], dYc[:, 2] = x1, x2
        for q in range(4, Rc.shape[1]):
            dYc[:, q] = self._tri_solve(D["tracer"], Rc[:, q])
        # Schur right-hand side for u₃ and the lower-triangular coupling to uₕ
        b3 = R3 - f2(Jm["u3_rho"], -Rrho) - f2(Jm["u3_rhoe"], self._tri_solve(Aee, Rre))
        b3 = b3 - f2(Jm["u3_uh"][0], x1) - f2(Jm["u3_uh"][1], x2)
        A33 = Jm["u3_u3"]

        def schur_mul(x):  # (A₃₃ − A₃ρ A_ρρ⁻¹ A_ρ3 − A₃e A_ee⁻¹ A_e3) x
            return self._tri_mul(A33, x) - f2(Jm["u3_rho"], -c2(Jm["rho_u3"], x)) - f2(Jm["u3_rhoe"], self._tri_solve(Aee, c2(Jm["rhoe_u3"], x)))

        # preconditioner: A₁₁ replaced by its main diagonal (−1 for ρ, d_ee for ρe_tot) → tridiagonal
        l, d, u = [a.copy() for a in A33]
        for a21, a12, dinv in ((Jm["u3_rho"], Jm["rho_u3"], -np.ones_like(Rrho)), (Jm["u3_rhoe"], Jm["rhoe_u3"], FT(1) / Aee[1])):
            lo21, hi21 = a21
            lo12, hi12 = a12[0] * dinv, a12[1] * dinv
            l -= lo21 * cl(lo12)
            d -= lo21 * cl(hi12) + hi21 * ch(lo12)
            u -= hi21 * ch(hi12)
        P33 = (l, d, u)
        x3 = self._tri_solve(P33, b3)
        for _ in range(n_iters):
            x3 = x3 + self._tri_solve(P33, b3 - schur_mul(x3))
        dYf[:, 0] = x3
        dYc[:, 0] = -(Rrho - c2(Jm["rho_u3"], x3))
        dYc[:, 3] = self._tri_solve(Aee, Rre - c2(Jm["rhoe_u3"], x3))
        return dYc, dYf

    def jacobian_dense_column(self, Jm, h, j, i, ntr=0):
        """Test helper: the full Jacobian of one column as a dense matrix over (ρ, uₕ₁, uₕ₂, ρe_tot, tracers…, u₃)."""
        nv = self.nv
        nc = 4 + ntr
        N = nc * nv + nv + 1
        M = np.zeros((N, N))
        o = lambda q: q * nv
        o3 = nc * nv
        sel = lambda a: np.asarray(a[h, j, i], dtype=np.float64)
        D = Jm.get("diff")
        for q in range(nc):
            M[o(q) : o(q) + nv, o(q) : o(q) + nv] = -np.eye(nv)
        if D is not None:
            def put_tri(q, tri):
                l, d, u = [sel(a) for a in tri]
                B = np.diag(d) + np.diag(l[1:], -1) + np.diag(u[:-1], 1)
                M[o(q) : o(q) + nv, o(q) : o(q) + nv] = B
            put_tri(3, D["rhoe_rhoe"])
            if D["uh_uh"] is not None:
                put_tri(1, D["uh_uh"]); put_tri(2, D["uh_uh"])
            for q in range(4, nc):
                put_tri(q, D["tracer"])
        for q, key in ((0, "rho_u3"), (3, "rhoe_u3")) + (((4, "rhoq_u3"),) if "rhoq_u3" in Jm else ()):
            lo, hi = [sel(a) for a in Jm[key]]
            for k in range(nv):
                M[o(q) + k, o3 + k] += lo[k]
                M[o(q) + k, o3 + k + 1] += hi[k]
        for q, blk in ((0, Jm["u3_rho"]), (3, Jm["u3_rhoe"]), (1, Jm["u3_uh"][0]), (2, Jm["u3_uh"][1])) + (((4, Jm["u3_rhoq"]),) if "u3_rhoq" in Jm else ()):
            lo, hi = [sel(a) for a in blk]
            for fidx in range(nv + 1):
                if fidx > 0:
                    M[o3 + fidx, o(q) + fidx - 1] += lo[fidx]
                if fidx < nv:
                    M[o3 + fidx, o(q) + fidx] += hi[fidx]
        l, d, u = [sel(a) for a in Jm["u3_u3"]]
        M[o3:, o3:] = np.diag(d) + np.diag(l[1:], -1) + np.diag(u[:-1], 1)
        return M

    def correct_implicit_advection_tendency(self, Yc, Yf, pc):
        """implicit_tendency.jl:322-339 (T_post_imp!)."""
        Ytc, Ytf = np.zeros_like(Yc), np.zeros_like(Yf)
        rho = Yc[:, 0]
        up = self.vertical_transport(rho, pc["fu3"], pc["h_tot"], self.N.dt, self.N.energy_upwinding)
        ce = self.vertical_transport(rho, pc["fu3"], pc["h_tot"], self.N.dt, "none")
        Ytc[:, 3] = up - ce
        if self.moist:  # :333-338
            q = Yc[:, 4] / rho
            Ytc[:, 4] = (self.vertical_transport(rho, pc["fu3"], q, self.N.dt, self.N.energy_upwinding)
                         - self.vertical_transport(rho, pc["fu3"], q, self.N.dt, "none"))
        return Ytc, Ytf

    # ------------------------------------------------------------------ Wfact / ldiv!
    def update_jacobian(self, Yc, Yf, pc, dtg):
        """manual_sparse_jacobian.jl:713-870 (dry, flat): band blocks stored by diagonals.

        Returned dict: bidiagonal centre-row blocks as (lo, hi) = entries at faces (k, k+1);
        bidiagonal face-row blocks as (lo, hi) = entries at centres (f-1, f) (zero on boundary
        rows); tridiagonal (u₃,u₃) as (l, d, u)."""
        FT, P, c, f = self.FT, self.P, self.c, self.f
        dtg = FT(dtg)
        rho, u1, u2 = Yc[:, 0], Yc[:, 1], Yc[:, 2]
        u3 = Yf[:, 0]
        K, p, T, h_tot = pc["K"], pc["p"], pc["T"], pc["h_tot"]
        moist = self.moist
        kappa = pc["Rm"] / pc["cvm"] if moist else FT(P.R_d) / FT(P.cv_d)  # ᶜkappa_m_field! :653-662
        z = lambda a: np.zeros_like(a)
        # :746-754
        dK_duh = self.ct12(u1, u2, c)  # Diag(CT12(uₕ)ᵀ)
        g33u3 = f.g33 * u3
        dK_du3 = (FT(0.5) * g33u3[..., :-1], FT(0.5) * g33u3[..., 1:])  # ᶜinterp_matrix⋅Diag(CT3(u₃))
        # :756 ᶠp_grad_matrix = Diag(-1/ᶠinterp(ρ)) ⋅ ᶠgradᵥ_matrix  (boundary rows zero)
        rf = self.interp_c2f(rho)
        pg_lo, pg_hi = z(rf), z(rf)
        pg_lo[..., 1:-1] = FT(1) / rf[..., 1:-1]
        pg_hi[..., 1:-1] = -FT(1) / rf[..., 1:-1]
        # :758-759 ᶜadvection_matrix = -(ᶜadvdivᵥ_matrix) ⋅ Diag(ᶠinterp(ρJ)/ᶠJ)
        rJf = self.interp_c2f(rho * c.J) / f.J
        Jf = f.J.copy()
        Jf[..., 0] = 0
        Jf[..., -1] = 0  # SetValue(0) rows/cols of the divergence matrix
        adv_lo = Jf[..., :-1] / c.J * rJf[..., :-1]
        adv_hi = -Jf[..., 1:] / c.J * rJf[..., 1:]
        # :767-768, :783-785
        hf = self.interp_c2f(h_tot)
        A_rho_u3 = (dtg * adv_lo * f.g33[..., :-1], dtg * adv_hi * f.g33[..., 1:])
        A_rhoe_u3 = (dtg * adv_lo * (hf * f.g33)[..., :-1], dtg * adv_hi * (hf * f.g33)[..., 1:])
        # :816-825
        thv = self.theta_v(T, p, pc)
        Pi = self.exner(p)
        dp_drho = kappa * (FT(P.T_0) * FT(P.cp_d) - K - self.Phi) + (FT(P.R_d) - kappa * FT(P.cv_d)) * T
        buoy = FT(P.cp_d) * self.interp_c2f(thv) * self.gradv_c2f(Pi) / rf  # diag on faces
        im_lo, im_hi = z(rf), z(rf)  # ᶠinterp_matrix (Extrapolate rows)
        im_lo[..., 1:-1] = FT(0.5)
        im_hi[..., 1:-1] = FT(0.5)
        im_hi[..., 0] = FT(1)
        im_lo[..., -1] = FT(1)
        cl = lambda a: np.concatenate([a[..., :1], a], -1)  # centre value at (f-1), padded
        ch = lambda a: np.concatenate([a, a[..., -1:]], -1)  # centre value at f, padded
        A_u3_rho = (
            dtg * (pg_lo * cl(dp_drho) + buoy * im_lo),
            dtg * (pg_hi * ch(dp_drho) + buoy * im_hi),
        )
        A_u3_rhoe = (dtg * pg_lo * cl(kappa), dtg * pg_hi * ch(kappa)) if moist else (dtg * pg_lo * kappa, dtg * pg_hi * kappa)
        # :855-868
        mk = lambda a: -kappa * a
        A_u3_uh = tuple(
            (dtg * pg_lo * cl(mk(rho) * dK_duh[a]), dtg * pg_hi * ch(mk(rho) * dK_duh[a])) for a in range(2)
        )
        X_lo, X_hi = pg_lo * cl(mk(rho)), pg_hi * ch(mk(rho))  # rows f, cols centres f-1, f
        dKlo, dKhi = dK_du3  # rows k: faces k, k+1
        l = X_lo * cl(dKlo)  # via centre f-1 → face f-1
        d = X_lo * cl(dKhi) + X_hi * ch(dKlo)
        u = X_hi * ch(dKhi)  # via centre f → face f+1
        beta = self.beta_rayleigh(f.z, P.alpha_rayleigh_w) if self.N.rayleigh_sponge else z(rf)
        A_u3_u3 = (dtg * l, dtg * (d - beta) - FT(1), dtg * u)
        Jm = dict(rho_u3=A_rho_u3, rhoe_u3=A_rhoe_u3, u3_rho=A_u3_rho, u3_rhoe=A_u3_rhoe, u3_uh=A_u3_uh, u3_u3=A_u3_u3)
        if moist:  # (ρq_tot, u₃) :770-790 and (u₃, ρq_tot) :827-831 with ᶜ∂p∂ρq_tot_field! :670-690
            qf = self.interp_c2f(Yc[:, 4] / rho)
            Jm["rhoq_u3"] = (dtg * adv_lo * (qf * f.g33)[..., :-1], dtg * adv_hi * (qf * f.g33)[..., 1:])
            dp_dq = kappa * (-FT(P.e_int_v0) - FT(P.R_d * P.T_0) - FT(P.cv_v - P.cv_d) * (T - FT(P.T_0))) + FT(P.R_v - P.R_d) * T
            Jm["u3_rhoq"] = (dtg * pg_lo * cl(dp_dq), dtg * pg_hi * ch(dp_dq))
        if self.vert_diff and self.implicit_diffusion:  # diffusion_flag = DerivativeFlag(atmos.diff_mode) (:88)
            self.update_diffusion_jacobian(Jm, Yc, pc, dtg)
        return Jm

    def ldiv(self, Jm, Rc, Rf):
        """jacobian.jl:78-82 → BlockArrowheadSolve(ρ, ρe_tot; alg₂ = BlockLowerTriangularSolve(uₕ))
        (manual_sparse_jacobian.jl:579-584) [UPSTREAM-RECALL]: Schur complement onto u₃, Thomas
        solve, back-substitution.  Scalar diagonal blocks are -I, (uₕ,uₕ) = -I (:476-481)."""
        FT = self.FT
        if "diff" in Jm:  # use_derivative(diffusion_flag) → ApproximateBlockArrowheadIterativeSolve (:538-578)
            return self.ldiv_iterative(Jm, Rc, Rf)
        dYc, dYf = np.zeros_like(Rc), np.zeros_like(Rf)
        Rrho, R1, R2, Rre = Rc[:, 0], Rc[:, 1], Rc[:, 2], Rc[:, 3]
        R3 = Rf[:, 0]
        cl = lambda a: np.concatenate([a[..., :1] * 0, a], -1)  # centre (f-1) value, 0 outside
        ch = lambda a: np.concatenate([a, a[..., -1:] * 0], -1)
        fl = lambda a: np.concatenate([a[..., :1] * 0, a[..., :-1]], -1)  # face f-1
        fh = lambda a: np.concatenate([a[..., 1:], a[..., -1:] * 0], -1)  # face f+1
        l, d, u = [a.copy() for a in Jm["u3_u3"]]
        # Schur: A22 + A21·A12 (A11 = -I)
        # moist without diffusion/sedimentation: ApproximateBlockArrowheadIterativeSolve (:538-578) with A₁₁ = −I — its main-diagonal
        # preconditioner IS A₁₁, so every iterate equals the exact arrowhead solve with one more scalar (ρq_tot)
        scal = [(0, "rho"), (3, "rhoe")] + ([(4, "rhoq")] if "rhoq_u3" in Jm else [])
        for a21, a12 in [(Jm["u3_" + k], Jm[k + "_u3"]) for _, k in scal]:
            lo21, hi21 = a21  # row f: centres f-1, f
            lo12, hi12 = a12  # row k: faces k, k+1
            l += lo21 * cl(lo12)
            d += lo21 * cl(hi12) + hi21 * ch(lo12)
            u += hi21 * ch(hi12)
        # velocities first: Δuₕ = -R_uₕ ; passive tracers only have the fallback -I block (:476-481)
        dYc[:, 1], dYc[:, 2] = -R1, -R2
        dYc[:, self.q0:] = -Rc[:, self.q0:]
        rhs = R3.copy()
        for a21, r in [(Jm["u3_" + k], Rc[:, idx]) for idx, k in scal]:
            rhs += a21[0] * cl(r) + a21[1] * ch(r)
        for a in range(2):
            r = (R1, R2)[a]
            rhs += Jm["u3_uh"][a][0] * cl(r) + Jm["u3_uh"][a][1] * ch(r)
        # Thomas algorithm
        n = rhs.shape[-1]
        cp = np.zeros_like(rhs)
        dp = np.zeros_like(rhs)
        cp[..., 0] = u[..., 0] / d[..., 0]
        dp[..., 0] = rhs[..., 0] / d[..., 0]
        for i in range(1, n):
            den = d[..., i] - l[..., i] * cp[..., i - 1]
            cp[..., i] = u[..., i] / den
            dp[..., i] = (rhs[..., i] - l[..., i] * dp[..., i - 1]) / den
        x = np.zeros_like(rhs)
        x[..., -1] = dp[..., -1]
        for i in range(n - 2, -1, -1):
            x[..., i] = dp[..., i] - cp[..., i] * x[..., i + 1]
        dYf[:, 0] = x
        # back-substitute scalars: -Δρ + A12 Δu₃ = R  ⇒ Δρ = A12 Δu₃ - R
        for idx, k in scal:
            a12 = Jm[k + "_u3"]
            dYc[:, idx] = a12[0] * x[..., :-1] + a12[1] * x[..., 1:] - Rc[:, idx]
        return dYc, dYf

    # ------------------------------------------------------------------ DSS
    def weighted_dss(self, fields):
        """Spaces.weighted_dss! (constrain_state.jl:59-64; discretization.md:139-154).
        ``fields`` = list of (kind, arrays): kind 'scalar' → [a]; 'c12' → [u1, u2] summed in the
        local (east, north) basis via ∂x/∂ξ [UPSTREAM-RECALL dss_transform]. In place."""
        A = np.asarray(self.grid.dxdxi, dtype=self.FT)
        Ainv = np.asarray(np.linalg.inv(self.grid.dxdxi), dtype=self.FT)
        w = self.dss_w
        offs, mem = self.dss_offs, self.dss_mem
        E, I, Jn = mem[:, 0], mem[:, 1], mem[:, 2]
        seg = np.repeat(np.arange(len(offs) - 1), np.diff(offs))
        nn = len(offs) - 1
        for kind, arrs in fields:
            if kind == "scalar":
                comps = [w[E, Jn, I, None] * arrs[0][E, Jn, I, :]]
            else:
                # covariant → physical: (u, v) = (A⁻¹)ᵀ (u1, u2)
                a1, a2 = arrs[0][E, Jn, I, :], arrs[1][E, Jn, I, :]
                Ai = Ainv[E, Jn, I]
                uu = Ai[:, 0, 0, None] * a1 + Ai[:, 1, 0, None] * a2
                vv = Ai[:, 0, 1, None] * a1 + Ai[:, 1, 1, None] * a2
                comps = [w[E, Jn, I, None] * uu, w[E, Jn, I, None] * vv]
            sums = []
            for cmp in comps:
                s = np.zeros((nn, cmp.shape[1]), dtype=self.FT)
                # fixed summation order: members in table order
                maxm = int(np.max(np.diff(offs)))
                for q in range(maxm):
                    sel = offs[:-1] + q
                    ok = sel < offs[1:]
                    s[ok] += cmp[sel[ok]]
                sums.append(s[seg])
            if kind == "scalar":
                arrs[0][E, Jn, I, :] = sums[0]
            else:
                Am = A[E, Jn, I]
                arrs[0][E, Jn, I, :] = Am[:, 0, 0, None] * sums[0] + Am[:, 1, 0, None] * sums[1]
                arrs[1][E, Jn, I, :] = Am[:, 0, 1, None] * sums[0] + Am[:, 1, 1, None] * sums[1]

    def dss_state(self, Yc, Yf):
        self.weighted_dss(
            [("scalar", [Yc[:, 0]]), ("c12", [Yc[:, 1], Yc[:, 2]]), ("scalar", [Yc[:, 3]])]
            + [("scalar", [Yc[:, q]]) for q in range(4, Yc.shape[1])]
            + [("scalar", [Yf[:, 0]])]
        )

    # ------------------------------------------------------------------ lim!  (limited_tendencies.jl:64-122)
    def neighboring_elements(self):
        """``Topologies.local_neighboring_elements``: elements sharing a vertex (hence also a face) with each element
        [UPSTREAM-RECALL ClimaCore 0.15.1 src/Topologies/topology2d.jl], from the vertex tables."""
        if getattr(self, "_nbrs", None) is None:
            topo = self.grid.topology
            lv, lo = np.asarray(topo.local_vertices), np.asarray(topo.local_vertex_offset)
            nb = [set() for _ in range(self.grid.nelems)]
            for v in range(len(lo) - 1):
                es = [int(e) for e in lv[lo[v]:lo[v + 1], 0]]
                for e in es:
                    nb[e].update(es)
            self._nbrs = [sorted(s - {e}) for e, s in enumerate(nb)]
        return self._nbrs

    def limiter_bounds(self, ref_rhoq, ref_rho):
        """``Limiters.compute_bounds!`` [UPSTREAM-RECALL ClimaCore src/Limiters/quasimonotone.jl]: per element and level the
        min / max of q = ρq/ρ over the Nq² nodes (compute_element_bounds!), then min / max over the element and its
        vertex-neighbours (compute_neighbor_bounds_local!).  Returns (q_min, q_max), each [h, v]."""
        q = ref_rhoq / ref_rho
        lo, hi = q.min(axis=(1, 2)), q.max(axis=(1, 2))
        qmin, qmax = lo.copy(), hi.copy()
        for e, ns in enumerate(self.neighboring_elements()):
            for n in ns:
                qmin[e] = np.minimum(qmin[e], lo[n])
                qmax[e] = np.maximum(qmax[e], hi[n])
        return qmin, qmax

    def apply_limiter(self, rhoq, rho, qmin, qmax):
        """``Limiters.apply_limiter!`` → ``apply_limit_slab!`` [UPSTREAM-RECALL]: per (element, level) slab clip ρq to
        [ρ q_min, ρ q_max] (bounds relaxed to contain the slab mean) and redistribute the clipped tracer mass over the nodes that
        still have room, in proportion to ρ·WJ; at most Nq² iterations, stop when |Δmass| ≤ rtol·|mass| with rtol = eps(FT).
        All slabs are iterated together; sums run over the nodes in the order n = 4j + i.  ``rhoq`` is modified in place."""
        FT = self.FT
        nh, nq, _, nv = rhoq.shape
        r = rho.reshape(nh, nq * nq, nv)
        x = rhoq.reshape(nh, nq * nq, nv).copy()
        w = np.asarray(self.c.WJ, dtype=FT).reshape(nh, nq * nq, nv)
        nn = nq * nq
        rtol = np.finfo(FT).eps

        def ssum(a):  # sequential sum over the nodes (the kernel's order)
            t = np.zeros(a.shape[:1] + a.shape[2:], dtype=FT)
            for n in range(nn):
                t = t + a[:, n]
            return t

        total_mass = ssum(r * w)
        tracer_mass = ssum(x * w)
        q_avg = tracer_mass / total_mass
        lo, hi = np.minimum(qmin, q_avg)[:, None, :], np.maximum(qmax, q_avg)[:, None, :]
        active = np.ones(total_mass.shape, dtype=bool)
        for _ in range(nn):
            xmax, xmin = r * hi, r * lo
            over, under = x > xmax, (x < xmin) & ~(x > xmax)
            d = np.where(over, (x - xmax) * w, np.where(under, (x - xmin) * w, FT(0)))
            dm = ssum(d)
            a3 = active[:, None, :]
            x = np.where(a3 & over, xmax, np.where(a3 & under, xmin, x))
            active = active & ~(np.abs(dm) <= rtol * np.abs(tracer_mass))
            if not active.any():
                break
            add = dm > 0
            room = np.where(add[:, None, :], x < xmax, x > xmin)
            mass_at = ssum(np.where(room, r * w, FT(0)))
            with np.errstate(divide="ignore", invalid="ignore"):
                dq = dm / mass_at
            upd = active[:, None, :] & room
            x = np.where(upd, x + r * dq[:, None, :], x)
        rhoq[...] = x.reshape(rhoq.shape)

    def limiters_func(self, Yc, ref_Yc):
        """lim!(Y, p, t, ref_Y) (limited_tendencies.jl:64-122): the SEM quasi-monotone limiter (bounds from ref_Y, applied to every
        tracer of Y.c in place; ``apply_sem_quasimonotone_limiter``), then the vertical mass-borrowing limiter
        (``tracer_nonnegativity_method: vertical_water_borrowing``).  No-op when neither is configured."""
        if getattr(self.N, "apply_sem_quasimonotone_limiter", False):
            for q in range(4, Yc.shape[1]):
                qmin, qmax = self.limiter_bounds(ref_Yc[:, q], ref_Yc[:, 0])
                self.apply_limiter(Yc[:, q], Yc[:, 0], qmin, qmax)
        # :95-121 vertical water borrowing (tracer_nonnegativity_method: vertical_water_borrowing, cache.jl:216-219):
        # χ = ρχ/ρ in scratch, Limiters.apply_limiter!(χ, ρ, VerticalMassBorrowingLimiter((0,))), ρχ = χ·ρ; all tracers
        # (vertical_water_borrowing_species = nothing)
        if getattr(self.N, "tracer_nonnegativity_method", None) == "vertical_water_borrowing":
            rho = Yc[:, 0]
            for q in range(4, Yc.shape[1]):
                chi = Yc[:, q] / rho
                self.vertical_mass_borrowing(chi, rho)
                Yc[:, q] = chi * rho

    def vertical_mass_borrowing(self, q, rho, qmin=0.0):
        """ClimaCore ``Limiters.apply_limiter!(q, ρ, ::VerticalMassBorrowingLimiter)`` [UPSTREAM-RECALL ClimaCore 0.15.1
        src/Limiters/vertical_mass_borrowing_limiter.jl, after E3SM's ``massborrow``]: per column, sweep level 1 → Nv carrying the
        mass deficit ``bmass`` (weights ρ·Δz) — a level that would end below ``qmin`` is set to ``qmin`` and passes its deficit on —
        then sweep Nv → 1 while a deficit remains.  In place on ``q``."""
        FT = self.FT
        m = rho * np.asarray(self.grid.dz_c, dtype=FT)  # ρ · Fields.Δz_field
        qm = FT(qmin)
        bmass = np.zeros_like(q[..., 0])
        nv = q.shape[-1]
        for v in range(nv):
            nmass = q[..., v] + bmass / m[..., v]
            pos = nmass > qm
            q[..., v] = np.where(pos, nmass, qm)
            bmass = np.where(pos, FT(0), (nmass - qm) * m[..., v])
        for v in range(nv - 1, -1, -1):
            need = bmass < 0
            nmass = q[..., v] + bmass / m[..., v]
            pos = nmass > qm
            q[..., v] = np.where(need, np.where(pos, nmass, qm), q[..., v])
            bmass = np.where(need, np.where(pos, FT(0), (nmass - qm) * m[..., v]), bmass)
        return q

    # ------------------------------------------------------------------ T_exp_T_lim!
    def vector_laplacian(self, u1, u2, u3, G):
        """hyperdiffusion.jl:141 / :273-277: C123(wgradₕ(divₕ(u))) - C123(wcurlₕ(C123(curlₕ(u))))
        returning (graddiv_12, curlcurl_123) separately so callers can scale the grad-div part."""
        c1, c2 = self.ct12(u1, u2, G)
        gd1, gd2 = self.wgrad(self.div(c1, c2, G), G)
        w3 = self.curl3(u1, u2, G)  # CT3
        w1, w2 = self.curl12(u3, G)  # CT12
        k1, k2 = self.c12(w1, w2, G)
        k3 = G.c33 * w3
        e3c = self.wcurl3(k1, k2, G)
        e1c, e2c = self.wcurl12(k3, G)
        e1, e2 = self.c12(e1c, e2c, G)
        e3 = G.c33 * e3c
        return (gd1, gd2), (e1, e2, e3)

    def remaining_tendency(self, Yc, Yf, pc, with_lim=False):
        """remaining_tendency.jl:48-58: returns (Ytc, Ytf) and, with ``with_lim``, also Yₜ_lim.c — the limited
        tracer tendencies (horizontal advection + tracer hyperdiffusion; zero for the non-tracer components)."""
        Ytc, Ytf, L = self._rt_pre(Yc, Yf, pc)
        Ylc = np.zeros_like(Yc)
        self._tracer_pre(Ytc, Ylc, Yc, Yf, pc)
        if L is not None:
            Lq = self._tracer_laplacians(Yc, pc)
            self.weighted_dss([("c12", [L[0], L[1]]), ("scalar", [L[2]]), ("scalar", [L[3]])] + [("scalar", [a]) for a in Lq])  # :18-21
            self._rt_post(Ytc, Ytf, Yc, L, pc, Lq[0] if self.moist else None)
            self._tracer_post(Ylc, Yc, Lq)
        return (Ytc, Ytf, Ylc) if with_lim else (Ytc + Ylc, Ytf)

    # ---- passive grid-scale tracers ρχ (components 4.. of Y.c), e.g. the chemistry tracer ρq_gas_A
    def _tracer_pre(self, Ytc, Ylc, Yc, Yf, pc):
        """horizontal_tracer_advection_tendency! (advection.jl:113-143, into Yₜ_lim), explicit vertical transport
        with ``tracer_upwinding`` (advection.jl:249-255, into Yₜ) and the viscous-sponge tracer term
        (viscous_sponge.jl:226-231, into Yₜ).  Element-local."""
        FT, N, c = self.FT, self.N, self.c
        rho, u1, u2 = Yc[:, 0], Yc[:, 1], Yc[:, 2]
        c1, c2 = self.ct12(u1, u2, c)
        for q in range(4, Yc.shape[1]):
            chi = Yc[:, q] / rho
            Ylc[:, q] -= self.split_div(rho * c1, rho * c2, chi, c)  # every tracer variable, ρq_tot included (advection.jl:121-124)
            if q < self.q0:
                # ρq_tot: vertical transport is implicit (advection.jl:250); viscous sponge on the total water: the aggregate
                # tendency goes to ρq_tot AND ρ, the water enthalpy flux to ρe_tot (viscous_sponge.jl:158-199)
                if N.viscous_sponge:
                    g = self.grad(chi)
                    g = self.ct12(g[0], g[1], c)
                    d = self.beta_viscous(c.z) * self.wdiv(rho * g[0], rho * g[1], c)
                    Ytc[:, q] += d
                    Ytc[:, 0] += d
                    hw = rho * self.h_eff_plus_Phi(pc["T"], pc["qt"], pc["ql"], pc["qi"])
                    Ytc[:, 3] += self.beta_viscous(c.z) * self.wdiv(hw * g[0], hw * g[1], c)
                continue
            Ytc[:, q] += self.vertical_transport(rho, pc["fu3"], chi, N.dt, N.tracer_upwinding)
            if N.viscous_sponge:
                g = self.grad(chi)
                g = self.ct12(g[0], g[1], c)
                Ytc[:, q] += self.beta_viscous(c.z) * self.wdiv(rho * g[0], rho * g[1], c)

    def _tracer_laplacians(self, Yc, pc=None):
        """prep_tracer_hyperdiffusion_tendency! (hyperdiffusion.jl:420-432): ∇²χ = wdivₕ(gradₕ(ρχ/ρ)).  For ρq_tot the field that
        is used afterwards is ᶜ∇²q_tot_eff = wdivₕ(gradₕ(q_tot − q_tot_r(p))) (hyperdiffusion.jl:148-165); the plain ∇²q_tot the
        reference also computes (and DSSes) is never read, so the q_tot slot carries ∇²q_tot_eff."""
        out = []
        for q in range(4, Yc.shape[1]):
            chi = Yc[:, q] / Yc[:, 0]
            if q < self.q0:
                chi = chi - self.q_tot_r(pc["p"])
            g = self.grad(chi)
            out.append(self.wdiv(*self.ct12(g[0], g[1], self.c), self.c))
        return out

    def _tracer_post(self, Ylc, Yc, Lq):
        """apply_tracer_hyperdiffusion_tendency! (hyperdiffusion.jl:524-532): ρχₜ −= ν₄ₛ wdivₕ(ρ gradₕ(∇²χ))."""
        rho = Yc[:, 0]
        for k, q in enumerate(range(4, Yc.shape[1])):
            g = self.grad(Lq[k])
            g = self.ct12(g[0], g[1], self.c)
            d = self.nu4_scalar * self.wdiv(rho * g[0], rho * g[1], self.c)
            Ylc[:, q] -= d
            if q < self.q0:  # hyperdiffusion.jl:480-484: the water mass tendency also enters Yₜ.c.ρ
                Ylc[:, 0] -= d

    def _rt_post(self, Ytc, Ytf, Yc, L, pc=None, Lq_tot=None):
        """apply_hyperdiffusion_tendency! (hyperdiffusion.jl:247-316) on the DSSed ∇² fields (element-local)."""
        FT, N, c = self.FT, self.N, self.c
        rho = Yc[:, 0]
        L1, L2, L3, Ls = L
        (gd, cc) = self.vector_laplacian(L1, L2, L3, c)  # apply :273-277
        ddf = FT(N.divergence_damping_factor)
        Q1, Q2, Q3 = ddf * gd[0] - cc[0], ddf * gd[1] - cc[1], -cc[2]
        Ytc[:, 1] -= self.nu4_vort * Q1
        Ytc[:, 2] -= self.nu4_vort * Q2
        Ytf[:, 0] -= self.nu4_vort * self.winterp_c2f(c.J * rho, Q3)
        gL = self.grad(Ls)
        gL = self.ct12(gL[0], gL[1], c)
        Ytc[:, 3] -= self.nu4_scalar * self.wdiv(rho * gL[0], rho * gL[1], c)  # :291,307
        if self.moist:  # water enthalpy flux ν ρ (h_eff + Φ) gradₕ(∇²q_tot_eff)  (:293-306)
            gq = self.grad(Lq_tot)
            gq = self.ct12(gq[0], gq[1], c)
            hw = rho * self.h_eff_plus_Phi(pc["T"], pc["qt"], pc["ql"], pc["qi"])
            Ytc[:, 3] -= self.nu4_scalar * self.wdiv(hw * gq[0], hw * gq[1], c)

    def _rt_pre(self, Yc, Yf, pc):
        """Everything of remaining_tendency! that is element-local before the DSS of the ∇² fields.
        Returns (Ytc, Ytf, (∇²u₁, ∇²u₂, ∇²u₃, ∇²s_d) or None)."""
        FT, P, N, c, f = self.FT, self.P, self.N, self.c, self.f
        L = None
        Ytc, Ytf = np.zeros_like(Yc), np.zeros_like(Yf)
        rho, u1, u2, rhoe = Yc[:, 0], Yc[:, 1], Yc[:, 2], Yc[:, 3]
        u3 = Yf[:, 0]
        K, T, p, h_tot, u3c, fu3 = pc["K"], pc["T"], pc["p"], pc["h_tot"], pc["u3c"], pc["fu3"]
        cp_d = FT(P.cp_d)
        # ---- horizontal_dynamics_tendency! advection.jl:36-91
        c1, c2 = self.ct12(u1, u2, c)  # horizontal contravariant components of ᶜu
        one = np.ones_like(rho)
        Ytc[:, 0] -= self.split_div(rho * c1, rho * c2, one, c)
        Ytc[:, 3] -= self.split_div(rho * c1, rho * c2, h_tot, c)
        Pi = self.exner(p)
        dth = self.theta_v(T, p, pc) - self.theta_vr(p)
        g1 = self.grad(K + self.Phi - self.phi_r(p))
        gPi, gth, gthPi = self.grad(Pi), self.grad(dth), self.grad(dth * Pi)
        for a in range(2):
            Ytc[:, 1 + a] -= g1[a] + cp_d * (dth * gPi[a] + gthPi[a] - Pi * gth[a]) / FT(2)
        # ---- hyperdiffusion_tendency! remaining_tendency.jl:15-24
        if N.hyperdiff:
            (gd, cc) = self.vector_laplacian(u1, u2, u3c, c)  # prep :141
            L1, L2, L3 = gd[0] - cc[0], gd[1] - cc[1], -cc[2]
            s_d = cp_d * (T - FT(P.T_0)) + self.Phi - self.sd_r(p)  # :142-147
            gs = self.grad(s_d)
            Ls = self.wdiv(*self.ct12(gs[0], gs[1], c), c)
            L = (L1, L2, L3, Ls)
        # ---- explicit_vertical_advection_tendency! advection.jl:205-290
        w3 = self.wcurl3(u1, u2, c)  # ᶜω³ :228
        o1, o2 = self.curlv_c2f(u1, u2)  # ᶠω¹² :233
        a1, a2 = self.wcurl12(u3, f)  # :237
        o1, o2 = o1 + a1, o2 + a2
        if self.f1_f is not None:  # deep atmosphere :273-278
            t1, t2 = self.f1_f + o1, self.f2_f + o2
        else:
            t1, t2 = o1, o2
        V = self.interp_c2f(rho * c.J) * fu3  # CT3 component
        # (CT12 × CT3) → C12: J (ω²V, -ω¹V)
        x1, x2 = f.J * (t2 * V), -f.J * (t1 * V)
        tot3 = self.f3_c + w3
        Ytc[:, 1] -= self.interp_f2c(x1) / (rho * c.J) + (-c.J * tot3 * c2)  # CT3 × CT12 → C12
        Ytc[:, 2] -= self.interp_f2c(x2) / (rho * c.J) + (c.J * tot3 * c1)
        ub1, ub2 = self.interp_c2f(c1), self.interp_c2f(c2)
        Ytf[:, 0] -= f.J * (t1 * ub2 - t2 * ub1) + self.gradv_c2f(K)  # CT12 × CT12 → C3
        # ---- additional_tendency! remaining_tendency.jl:166-171
        if N.rayleigh_sponge:
            b = self.beta_rayleigh(c.z, P.alpha_rayleigh_uh)
            Ytc[:, 1] += -b * u1
            Ytc[:, 2] += -b * u2
        if N.held_suarez:  # held_suarez.jl:111-296 (flat surface: z_surface = 0 ⇒ p_surface = MSLP)
            lat = self.lat_rad
            s2, c2 = np.asarray(np.sin(lat) ** 2, dtype=FT), np.asarray(np.cos(lat) ** 2, dtype=FT)
            sigma = p / FT(P.MSLP)
            hf = np.maximum(FT(0), (sigma - FT(P.sigma_b)) / FT(1 - P.sigma_b))
            k_a, k_s, k_f = FT(1 / (40 * P.day)), FT(1 / (4 * P.day)), FT(1 / P.day)
            ppr = p / FT(P.p_ref_theta)
            Teq = np.maximum(FT(P.T_min_hs), (FT(P.T_equator_dry) - FT(P.dT_y_dry) * s2 - FT(P.dtheta_z) * np.log(ppr) * c2) * ppr ** FT(P.kappa_d))
            dRT = (k_a + (k_s - k_a) * hf * c2 * c2) * rho * (p / (rho * FT(P.R_d)) - Teq)
            Ytc[:, 1] += -(k_f * hf) * u1
            Ytc[:, 2] += -(k_f * hf) * u2
            Ytc[:, 3] += -dRT * FT(P.cv_d)
        if N.viscous_sponge:  # viscous_sponge.jl:138-175
            bc, bf = self.beta_viscous(c.z), self.beta_viscous(f.z)
            (gd, cc) = self.vector_laplacian(u1, u2, np.zeros_like(u1), c)
            Ytc[:, 1] += bc * (gd[0] - cc[0])
            Ytc[:, 2] += bc * (gd[1] - cc[1])
            g3 = self.grad(u3)
            Ytf[:, 0] += bf * self.wdiv(*self.ct12(g3[0], g3[1], f), f)
            gs = self.grad(cp_d * (T - FT(P.T_0)) + self.Phi)
            gs = self.ct12(gs[0], gs[1], c)
            Ytc[:, 3] += bc * self.wdiv(rho * gs[0], rho * gs[1], c)
        if self.vert_diff and not self.implicit_diffusion:  # remaining_tendency.jl:185-195 (diff_mode == Explicit())
            self.vertical_diffusion_boundary_layer_tendency(Ytc, Yc, pc)
        return Ytc, Ytf, L

    # ------------------------------------------------------------------ ARS343 step
    def _implicit_stage_local(self, Uc, Uf, dtg, log):
        """cache_imp! → Wfact → T_imp! → ldiv! → U −= ΔU → cache_imp! → T_post_imp!, in place (element-local)."""
        FT = self.FT
        pc = self.set_implicit_precomputed_quantities(Uc, Uf)
        log("cache_imp")
        tc, tf = Uc.copy(), Uf.copy()
        Jm = self.update_jacobian(Uc, Uf, pc, dtg)
        log("wfact")
        Rc, Rf = self.implicit_tendency(Uc, Uf, pc)
        log("t_imp")
        Rc = tc + FT(dtg) * Rc - Uc
        Rf = tf + FT(dtg) * Rf - Uf
        dc, df = self.ldiv(Jm, Rc, Rf)
        log("ldiv")
        Uc -= dc
        Uf -= df
        pc = self.set_implicit_precomputed_quantities(Uc, Uf)
        log("cache_imp")
        if self.N.energy_upwinding != "none":
            pc_, pf_ = self.correct_implicit_advection_tendency(Uc, Uf, pc)
            log("t_post_imp")
            Uc += FT(dtg) * pc_
            Uf += FT(dtg) * pf_
        return tc, tf

    def step(self, Yc, Yf, trace=None, pool=None, nchunks=1):
        """One IMEX-ARK (ARS343) step with one Newton iteration per implicit stage, hook order
        reconstructed from ClimaTimeSteppers 0.10.6 [UPSTREAM-RECALL] (DESIGN.md
        "Step trace").  Returns the new (Yc, Yf).

        ``pool`` (a ``concurrent.futures.ThreadPoolExecutor``) with ``nchunks`` > 1 runs every element-local
        phase on element chunks in parallel (NumPy releases the GIL); the DSS stays global.  The result is
        bitwise identical to the serial path (checked in tests/test_oracle_identities.py)."""
        from climaatmos_jl_b200.params import ars343

        FT = self.FT
        a_exp, a_imp, b_exp, b_imp, gam = ars343()
        dt = self.N.dt
        uc, uf = Yc, Yf
        nel = Yc.shape[0]
        Texp, Timp, Tlim = [None] * 4, [None] * 4, [None] * 4
        limiter = (bool(getattr(self.N, "apply_sem_quasimonotone_limiter", False))
                   or getattr(self.N, "tracer_nonnegativity_method", None) == "vertical_water_borrowing") and Yc.shape[1] > 4
        log = (lambda s: trace.append(s)) if trace is not None else (lambda s: None)
        nolog = lambda s: None
        if pool is not None and nchunks > 1:
            bounds = np.linspace(0, nel, nchunks + 1).astype(int)
            chunks = [(slice(int(a), int(b)), self.slice(slice(int(a), int(b)))) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]

            def pmap(fn):
                return list(pool.map(lambda cs: fn(cs[1], cs[0]), chunks))
        else:
            chunks = [(slice(0, nel), self)]

            def pmap(fn):
                return [fn(self, slice(0, nel))]

        def increment(i_coefs_exp, i_coefs_imp):
            Uc, Uf = np.empty_like(uc), np.empty_like(uf)
            if limiter:  # CTS update_stage!: U = u + dt Σ a_exp T_lim ; lim!(U, p, t, u) ; then the unlimited increments
                Uc[...] = uc
                Uf[...] = uf
                for j in range(4):
                    if i_coefs_exp[j] != 0 and Tlim[j] is not None:
                        Uc += FT(dt * i_coefs_exp[j]) * Tlim[j]
                if any(i_coefs_exp[j] != 0 and Tlim[j] is not None for j in range(4)):
                    self.limiters_func(Uc, uc)
                    log("lim")

            def work(sub, sl):
                if not limiter:
                    Uc[sl] = uc[sl]
                    Uf[sl] = uf[sl]
                for j in range(4):
                    if i_coefs_exp[j] != 0 and Texp[j] is not None:
                        Uc[sl] += FT(dt * i_coefs_exp[j]) * Texp[j][0][sl]
                        Uf[sl] += FT(dt * i_coefs_exp[j]) * Texp[j][1][sl]
                    if i_coefs_imp[j] != 0 and Timp[j] is not None:
                        Uc[sl] += FT(dt * i_coefs_imp[j]) * Timp[j][0][sl]
                        Uf[sl] += FT(dt * i_coefs_imp[j]) * Timp[j][1][sl]

            pmap(work)
            return Uc, Uf

        def t_exp(Uc, Uf):
            Ytc, Ytf = np.empty_like(Uc), np.empty_like(Uf)
            Ylc = np.zeros_like(Uc) if limiter else None
            nq = Uc.shape[1] - 4
            Ls = [np.empty_like(Uc[:, 0]) for _ in range(4 + nq)] if self.N.hyperdiff else None

            def pre(sub, sl):
                pc = sub.set_implicit_precomputed_quantities(Uc[sl], Uf[sl])
                a, b, L = sub._rt_pre(Uc[sl], Uf[sl], pc)
                lim = np.zeros_like(a)
                sub._tracer_pre(a, lim, Uc[sl], Uf[sl], pc)
                if limiter:
                    Ytc[sl], Ytf[sl], Ylc[sl] = a, b, lim
                else:
                    Ytc[sl], Ytf[sl] = a + lim, b  # lim! is a no-op (no limiter configured): T_lim joins T_exp
                if L is not None:
                    for k in range(4):
                        Ls[k][sl] = L[k]
                    for k, lq in enumerate(sub._tracer_laplacians(Uc[sl], pc)):
                        Ls[4 + k][sl] = lq

            pmap(pre)
            if Ls is not None:
                self.weighted_dss([("c12", [Ls[0], Ls[1]]), ("scalar", [Ls[2]]), ("scalar", [Ls[3]])] + [("scalar", [a]) for a in Ls[4:]])

                def post(sub, sl):
                    pc = sub.set_implicit_precomputed_quantities(Uc[sl], Uf[sl]) if self.moist else None  # (idempotent on a filtered state)
                    sub._rt_post(Ytc[sl], Ytf[sl], Uc[sl], tuple(a[sl] for a in Ls[:4]), pc, Ls[4][sl] if self.moist else None)
                    sub._tracer_post(Ylc[sl] if limiter else Ytc[sl], Uc[sl], [a[sl] for a in Ls[4:]])

                pmap(post)
            return (Ytc, Ytf, Ylc) if limiter else (Ytc, Ytf)

        for i in range(4):
            Uc, Uf = increment(a_exp[i], a_imp[i])
            if i != 0:
                self.dss_state(Uc, Uf)
                log("dss")
            if a_imp[i][i] != 0:
                dtg = dt * a_imp[i][i]
                tc, tf = np.empty_like(Uc), np.empty_like(Uf)

                def imp(sub, sl, first=[True]):
                    lg = log if sl.start == 0 else nolog
                    a, b = sub._implicit_stage_local(Uc[sl], Uf[sl], dtg, lg)
                    tc[sl], tf[sl] = a, b

                pmap(imp)
                self.dss_state(Uc, Uf)
                log("dss")
                log("cache_imp")
                Timp[i] = (np.empty_like(Uc), np.empty_like(Uf))

                def diff(sub, sl):
                    Timp[i][0][sl] = (Uc[sl] - tc[sl]) / FT(dtg)
                    Timp[i][1][sl] = (Uf[sl] - tf[sl]) / FT(dtg)

                pmap(diff)
            else:
                log("cache_imp")
            te = t_exp(Uc, Uf)
            Texp[i] = te[:2]
            if limiter:
                Tlim[i] = te[2]
            log("t_exp")
        uc, uf = increment(b_exp, b_imp)
        self.dss_state(uc, uf)
        log("dss")
        self.set_implicit_precomputed_quantities(uc, uf)
        log("cache")
        return uc, uf
