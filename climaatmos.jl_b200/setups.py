"""Initial conditions (input generators; init-time NumPy, float64 then cast).

* ``dry_baroclinic_wave`` restates ``deep_atmos_barowave_values`` / ``shallow_…``
  (src/setups/DryBaroclinicWave.jl:19-155, Ullrich et al. 2014) and the prognostic-variable
  assembly of src/setups/common/prognostic_variables.jl:44-62, 280-287.
* ``decaying_profile`` restates src/setups/DecayingProfile.jl:39-66 (Held–Suarez default IC).
State layout follows ClimaCore's VIJFH parent arrays: ``Yc[h, f, j, i, v]`` with
f = (ρ, uₕ₁, uₕ₂, ρe_tot) and ``Yf[h, 0, j, i, v]`` = u₃ (covariant).
"""
from __future__ import annotations

import numpy as np


def _barowave_values(z, lat, lon, params, perturb=True, deep=True):
    R_d, MSLP, grav, Om, R = params.R_d, params.MSLP, params.grav, params.Omega, params.planet_radius
    k = 3
    T_e, T_p = 310.0, 240.0
    T_0 = 0.5 * (T_e + T_p)
    Gam = 0.005
    A = 1 / Gam
    B = (T_0 - T_p) / T_0 / T_p
    C = 0.5 * (k + 2) * (T_e - T_p) / T_e / T_p
    b = 2
    H = R_d * T_0 / grav
    z_t, lam_c, phi_c, d_0, V_p = 15e3, 20.0, 40.0, R / 6, 1.0
    cosd = lambda a: np.cos(np.radians(a))
    sind = lambda a: np.sin(np.radians(a))
    e = np.exp(-((z / b / H) ** 2))
    tau1 = A * Gam / T_0 * np.exp(Gam * z / T_0) + B * (1 - 2 * (z / b / H) ** 2) * e
    tau2 = C * (1 - 2 * (z / b / H) ** 2) * e
    itau1 = A * (np.exp(Gam * z / T_0) - 1) + B * z * e
    itau2 = C * z * e
    if deep:
        rr = (z + R) / R
        I_T = (rr * cosd(lat)) ** k - (k / (k + 2)) * (rr * cosd(lat)) ** (k + 2)
        T = (R / (z + R)) ** 2 / (tau1 - tau2 * I_T)
        p = MSLP * np.exp(-grav / R_d * (itau1 - itau2 * I_T))
        U = grav / R * k * T * itau2 * ((rr * cosd(lat)) ** (k - 1) - (rr * cosd(lat)) ** (k + 1))
        u = -Om * (R + z) * cosd(lat) + np.sqrt((Om * (R + z) * cosd(lat)) ** 2 + (R + z) * cosd(lat) * U)
    else:
        I_T = cosd(lat) ** k - (k / (k + 2)) * cosd(lat) ** (k + 2)
        T = 1.0 / (tau1 - tau2 * I_T)
        p = MSLP * np.exp(-grav / R_d * (itau1 - itau2 * I_T))
        U = grav * k / R * itau2 * T * (cosd(lat) ** (k - 1) - cosd(lat) ** (k + 1))
        u = -Om * R * cosd(lat) + np.sqrt((Om * R * cosd(lat)) ** 2 + R * cosd(lat) * U)
    v = np.zeros_like(u)
    if perturb:
        F_z = (1 - 3 * (z / z_t) ** 2 + 2 * (z / z_t) ** 3) * (z <= z_t)
        arg = np.clip(sind(phi_c) * sind(lat) + cosd(phi_c) * cosd(lat) * cosd(lon - lam_c), -1, 1)
        r = R * np.arccos(arg)
        c3 = np.cos(np.pi * r / 2 / d_0) ** 3
        s1 = np.sin(np.pi * r / 2 / d_0)
        cond = ((0 < r) & (r < d_0) & (r != R * np.pi)).astype(float)
        sr = np.where(cond > 0, np.sin(r / R), 1.0)
        u = u + (-16 * V_p / 3 / np.sqrt(3.0) * F_z * c3 * s1
                 * (-sind(phi_c) * cosd(lat) + cosd(phi_c) * sind(lat) * cosd(lon - lam_c)) / sr * cond)
        v = v + (16 * V_p / 3 / np.sqrt(3.0) * F_z * c3 * s1 * cosd(phi_c) * sind(lon - lam_c) / sr * cond)
    return T, p, u, v


def _assemble(grid, params, T, p, u, v):
    """prognostic_variables.jl:44-62: ρ = p/(R_d T); uₕ = C12(UV(u,v)); ρe_tot = ρ(e_int + K + Φ) with the
    dry internal energy e_int = cv_d (T − T_0) − R_d T_0 (docs/src/thermodynamics.md:103-111)."""
    FT = grid.FT
    nel, nq, nv = grid.nelems, grid.nq, grid.nv
    rho = p / (params.R_d * T)
    s = (grid.radius + grid.z_c) / grid.radius if grid.deep else np.ones(nv)
    A = grid.dxdxi[..., None, :, :] * s[None, None, None, :, None, None]  # [h,j,i,v,a,b]
    u1 = A[..., 0, 0] * u + A[..., 1, 0] * v  # (∂x/∂ξ)ᵀ·(u, v)
    u2 = A[..., 0, 1] * u + A[..., 1, 1] * v
    z = grid.z_c[None, None, None, :]
    e_tot = params.cv_d * (T - params.T_0) - params.R_d * params.T_0 + 0.5 * (u * u + v * v) + params.grav * z
    Yc = np.zeros((nel, 4, nq, nq, nv), dtype=FT)
    Yc[:, 0] = rho
    Yc[:, 1] = u1
    Yc[:, 2] = u2
    Yc[:, 3] = rho * e_tot
    Yf = np.zeros((nel, 1, nq, nq, nv + 1), dtype=FT)
    return Yc, Yf


def dry_baroclinic_wave(grid, params, perturb=True):
    z = np.broadcast_to(grid.z_c[None, None, None, :], (grid.nelems, grid.nq, grid.nq, grid.nv))
    lat = grid.lat[..., None]
    lon = grid.lon[..., None]
    T, p, u, v = _barowave_values(z, lat, lon, params, perturb=perturb, deep=grid.deep)
    return _assemble(grid, params, T, p, u, v)


def moist_baroclinic_wave(grid, params, perturb=True, q_0=0.018):
    """MoistBaroclinicWave (src/setups/MoistBaroclinicWave.jl:50-92): the dry wave with the moisture profile
    q_tot = q_0 exp(−(φ/40°)⁴) exp(−((p − MSLP)/340 hPa)²) below 100 hPa (1e-12 above) and T = T_v/(1 + 0.608 q_tot); assembly by
    src/setups/common/prognostic_variables.jl:44-85 with q_liq = q_ice = 0: ρ = p/(R_m T), ρe_tot = ρ(I(T, q_tot) + K + Φ) with the
    moist internal energy of docs/src/thermodynamics.md:103-111, ρq_tot = ρ q_tot (component 4 of Y.c)."""
    z = np.broadcast_to(grid.z_c[None, None, None, :], (grid.nelems, grid.nq, grid.nq, grid.nv))
    lat = grid.lat[..., None]
    lon = grid.lon[..., None]
    T, p, u, v = _barowave_values(z, lat, lon, params, perturb=perturb, deep=grid.deep)
    q = np.where(p <= 1.0e4, 1e-12, q_0 * np.exp(-((lat / 40.0) ** 4)) * np.exp(-(((p - params.MSLP) / 3.4e4) ** 2)))
    T = T / (1 + 0.608 * q)
    P = params
    R_m = P.R_d * (1 - q) + P.R_v * q
    cv_m = P.cv_d + (P.cv_v - P.cv_d) * q
    Yc4, Yf = _assemble(grid, params, T, p, u, v)
    rho = p / (R_m * T)
    e_int = cv_m * (T - P.T_0) + q * P.e_int_v0 - (1 - q) * P.R_d * P.T_0
    e_tot = e_int + 0.5 * (u * u + v * v) + P.grav * grid.z_c[None, None, None, :]
    Yc = np.zeros((grid.nelems, 5, grid.nq, grid.nq, grid.nv), dtype=grid.FT)
    Yc[:, 0] = rho
    Yc[:, 1:3] = Yc4[:, 1:3]
    Yc[:, 3] = rho * e_tot
    Yc[:, 4] = rho * q
    return Yc, Yf


def decaying_profile(grid, params, perturb=True, T_s=290.0, T_min=220.0, H_t=8000.0):
    """Thermodynamics.jl ``DecayingTemperatureProfile`` [UPSTREAM-RECALL] + the 0.1 K
    longitudinal perturbation below 5 km of src/setups/DecayingProfile.jl:39-66; zero wind."""
    z = np.broadcast_to(grid.z_c[None, None, None, :], (grid.nelems, grid.nq, grid.nq, grid.nv)).astype(float)
    dT = T_s - T_min
    zp = z / H_t
    th = np.tanh(zp)
    dTp = dT / T_s
    H_sfc = params.R_d * T_s / params.grav
    Tv = T_s - dT * th
    p = -H_t * (zp + dTp * (np.log(1 - dTp * th) - np.log(1 + th) + zp))
    p = params.MSLP * np.exp(p / (H_sfc * (1 - dTp * dTp)))
    T = Tv.copy()
    if perturb:
        T = T + 0.1 * np.sin(np.radians(grid.lon[..., None])) * (z < 5000.0)
    u = np.zeros_like(T)
    return _assemble(grid, params, T, p, u, u.copy())
