"""ctypes binding of libb200dycore.so (include/b200_dycore.h).

This is the Python twin of the Julia ``ccall`` glue shown in INTEGRATION.md: it passes plain device
pointers, sizes and a stream handle; no torch types cross the boundary.  There is no CPU fallback:
``load()`` raises if the library is missing and every compute call raises on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200dycore.so")
REPO_ROOT = os.path.dirname(_HERE)

SYMBOLS = [
    "b200_create", "b200_destroy", "b200_last_error", "b200_nccl_unique_id", "b200_cache_imp",
    "b200_t_exp_lim", "b200_t_imp", "b200_wfact", "b200_ldiv", "b200_t_post_imp", "b200_dss",
    "b200_axpy_n", "b200_step_ars343", "b200_launch_count", "b200_halo_export", "b200_halo_import", "b200_build_dss_csr", "b200_debug_dss_csr", "b200_t_exp_phase", "b200_implicit_stage", "b200_lim", "b200_debug_jacobian",
]


class Dims(C.Structure):
    _fields_ = [("nh", C.c_int32), ("nh_ghost", C.c_int32), ("nv", C.c_int32), ("nq", C.c_int32),
                ("ft_bytes", C.c_int32), ("deep", C.c_int32), ("n_tracers", C.c_int32)]


class Geometry(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("dxdxi", "J2", "lat", "gll_w", "gll_D", "z_c", "z_f", "dz_c", "dz_f")] + [
        ("radius", C.c_double), ("z_max", C.c_double)]


class Topology(C.Structure):
    _fields_ = [("interior_faces", C.c_void_p), ("n_faces", C.c_int32), ("local_vertices", C.c_void_p),
                ("local_vertex_offset", C.c_void_p), ("n_verts", C.c_int32), ("n_neighbors", C.c_int32),
                ("neighbor_ranks", C.c_void_p), ("send_offset", C.c_void_p), ("send_elems", C.c_void_p),
                ("recv_offset", C.c_void_p), ("elem_gid", C.c_void_p)]


class Params(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("R_d", "cp_d", "cv_d", "T_0", "grav", "Omega", "p_ref_theta", "T_surf_ref",
                                          "T_min_ref", "T_min_sgs", "dt", "nu4_vorticity", "nu4_scalar",
                                          "divergence_damping_factor")] + [
        ("hyperdiff", C.c_int32), ("rayleigh_sponge", C.c_int32), ("zd_rayleigh", C.c_double),
        ("alpha_rayleigh_uh", C.c_double), ("alpha_rayleigh_w", C.c_double), ("viscous_sponge", C.c_int32),
        ("zd_viscous", C.c_double), ("kappa_2_sponge", C.c_double), ("energy_upwinding", C.c_int32),
        ("tracer_upwinding", C.c_int32), ("held_suarez", C.c_int32)] + [(n, C.c_double) for n in ("hs_day", "hs_sigma_b", "hs_dT_y", "hs_T_equator",
                                                               "hs_dtheta_z", "hs_T_min", "MSLP")] + [("sem_quasimonotone_limiter", C.c_int32)] + [
        (n, C.c_int32) for n in ("vert_diff", "implicit_diffusion", "approximate_linear_solve_iters",
                                 "disable_momentum_vertical_diffusion")] + [(n, C.c_double) for n in ("C_E", "H_diffusion", "D_0_diffusion")] + [
        ("vertical_water_borrowing_limiter", C.c_int32), ("microphysics_0M", C.c_int32)] + [
        (n, C.c_double) for n in ("R_v", "cp_v", "cp_l", "cp_i", "LH_v0", "LH_s0", "T_triple", "press_triple", "T_freeze", "T_icenuc",
                                  "pow_icenuc")]


class CachePtrs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("u_c", "u3_f", "K_c", "T_c", "p_c", "h_tot_c")]


NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/capi.cu for sm_100a into libb200dycore.so (in-tree)."""
    src_dir = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(src_dir, f) for f in os.listdir(src_dir)] + [os.path.join(REPO_ROOT, "include", "b200_dycore.h")]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + [os.path.join(src_dir, "capi.cu"), "-o", LIB_PATH, "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB_PATH


_lib = None


def load():
    """dlopen the library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    lib.b200_last_error.restype = C.c_char_p
    lib.b200_last_error.argtypes = [C.c_void_p]
    lib.b200_launch_count.restype = C.c_int64
    lib.b200_launch_count.argtypes = [C.c_void_p]
    vp, dbl, i32 = C.c_void_p, C.c_double, C.c_int32
    lib.b200_create.argtypes = [C.POINTER(vp), C.POINTER(Dims), C.POINTER(Geometry), C.POINTER(Topology), C.POINTER(Params), vp, C.c_int, C.c_int]
    lib.b200_destroy.argtypes = [vp]
    lib.b200_nccl_unique_id.argtypes = [vp]
    lib.b200_halo_export.argtypes = [vp, vp]
    lib.b200_halo_import.argtypes = [vp, vp, vp, vp]
    lib.b200_cache_imp.argtypes = [vp, vp, vp, C.POINTER(CachePtrs), vp]
    lib.b200_t_exp_lim.argtypes = [vp, vp, vp, vp, vp, vp, vp, dbl, vp]
    lib.b200_t_exp_phase.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    lib.b200_t_imp.argtypes = [vp, vp, vp, vp, vp, dbl, vp]
    lib.b200_implicit_stage.argtypes = [vp, vp, vp, vp, vp, dbl, vp]
    lib.b200_lim.argtypes = [vp, vp, vp, vp, vp, dbl, vp]
    lib.b200_wfact.argtypes = [vp, vp, vp, dbl, dbl, vp]
    lib.b200_ldiv.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.b200_t_post_imp.argtypes = [vp, vp, vp, vp, vp, dbl, vp]
    lib.b200_dss.argtypes = [vp, C.POINTER(vp), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), i32, vp]
    lib.b200_axpy_n.argtypes = [vp, vp, vp, vp, vp, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(dbl), vp]
    lib.b200_step_ars343.argtypes = [vp, vp, vp, dbl, i32, vp]
    lib.b200_debug_jacobian.argtypes = [vp, vp, C.c_int64, vp]
    lib.b200_debug_dss_csr.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(i32), C.POINTER(i32)]
    lib.b200_build_dss_csr.argtypes = [C.POINTER(Topology), vp, i32, vp, i32, C.POINTER(i32), C.POINTER(i32)]
    _lib = lib
    return lib


def check(status: int, what: str, ctx=None):
    """Raise on a non-zero status with the message of `ctx` (per-context, include/b200_dycore.h) or of the calling thread."""
    if status != 0:
        raise RuntimeError(f"{what} failed: {load().b200_last_error(ctx).decode()}")


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def make_params(P, N, grid) -> Params:
    h = grid.node_horizontal_length_scale()
    nu4v = N.nu4_vorticity_coeff * h**3 if N.hyperdiff else 0.0
    up = {"none": 0, "first_order": 1, "third_order": 2, "vanleer_limiter": 3}[N.energy_upwinding]
    return Params(
        R_d=P.R_d, cp_d=P.cp_d, cv_d=P.cv_d, T_0=P.T_0, grav=P.grav, Omega=P.Omega, p_ref_theta=P.p_ref_theta,
        T_surf_ref=P.T_surf_ref, T_min_ref=P.T_min_ref, T_min_sgs=P.T_min_sgs, dt=N.dt, nu4_vorticity=nu4v,
        nu4_scalar=nu4v / N.prandtl_number, divergence_damping_factor=N.divergence_damping_factor,
        hyperdiff=int(N.hyperdiff), rayleigh_sponge=int(N.rayleigh_sponge), zd_rayleigh=P.zd_rayleigh,
        alpha_rayleigh_uh=P.alpha_rayleigh_uh, alpha_rayleigh_w=P.alpha_rayleigh_w,
        viscous_sponge=int(N.viscous_sponge), zd_viscous=P.zd_viscous, kappa_2_sponge=P.kappa_2_sponge,
        energy_upwinding=up, tracer_upwinding={"none": 0, "first_order": 1, "third_order": 2, "vanleer_limiter": 3}[N.tracer_upwinding], held_suarez=int(N.held_suarez), hs_day=P.day, hs_sigma_b=P.sigma_b, hs_dT_y=P.dT_y_dry,
        hs_T_equator=P.T_equator_dry, hs_dtheta_z=P.dtheta_z, hs_T_min=P.T_min_hs, MSLP=P.MSLP,
        sem_quasimonotone_limiter=int(getattr(N, "apply_sem_quasimonotone_limiter", False)),
        vert_diff={None: 0, "VerticalDiffusion": 1, "DecayWithHeightDiffusion": 2}[getattr(N, "vert_diff", None)],
        implicit_diffusion=int(getattr(N, "implicit_diffusion", False)),
        approximate_linear_solve_iters=int(getattr(N, "approximate_linear_solve_iters", 1)),
        disable_momentum_vertical_diffusion=int(getattr(N, "disable_momentum_vertical_diffusion", False)),
        C_E=getattr(P, "C_E", 0.0), H_diffusion=getattr(P, "H_diffusion", 1.0), D_0_diffusion=getattr(P, "D_0_diffusion", 0.0),
        vertical_water_borrowing_limiter=int(getattr(N, "tracer_nonnegativity_method", None) == "vertical_water_borrowing"),
        microphysics_0M=int(getattr(N, "microphysics_model", None) == "0M"),
        R_v=P.R_v, cp_v=P.cp_v, cp_l=P.cp_l, cp_i=P.cp_i, LH_v0=P.LH_v0, LH_s0=P.LH_s0, T_triple=P.T_triple,
        press_triple=P.press_triple, T_freeze=P.T_freeze, T_icenuc=P.T_icenuc, pow_icenuc=P.pow_icenuc)


def create_context(grid, P, N, part=None, nccl_id: bytes | None = None, rank: int = 0, nranks: int = 1, n_tracers: int = 0):
    """b200_create from a ``SphereGrid`` (or one rank's partition of it, see partition.py)."""
    lib = load()
    ft = 4 if np.dtype(grid.FT) == np.float32 else 8
    keep = []  # keep host arrays alive during the call

    def arr(a, dt):
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return _ptr(a)

    if part is None:
        elems = np.arange(grid.nelems)
        nh, ng = grid.nelems, 0
        topo = grid.topology
        faces, lv, lvo = topo.interior_faces, topo.local_vertices, topo.local_vertex_offset
        T = Topology(interior_faces=arr(faces, np.int32), n_faces=len(faces), local_vertices=arr(lv, np.int32),
                     local_vertex_offset=arr(lvo, np.int32), n_verts=len(lvo) - 1, n_neighbors=0)
    else:
        elems = part.elems_ext
        nh, ng = part.nh, part.nh_ghost
        T = Topology(interior_faces=arr(part.interior_faces, np.int32), n_faces=len(part.interior_faces),
                     local_vertices=arr(part.local_vertices, np.int32),
                     local_vertex_offset=arr(part.local_vertex_offset, np.int32), n_verts=len(part.local_vertex_offset) - 1,
                     n_neighbors=len(part.neighbor_ranks), neighbor_ranks=arr(part.neighbor_ranks, np.int32),
                     send_offset=arr(part.send_offset, np.int32), send_elems=arr(part.send_elems, np.int32),
                     recv_offset=arr(part.recv_offset, np.int32), elem_gid=arr(part.elems_ext, np.int64))
    dims = Dims(nh=nh, nh_ghost=ng, nv=grid.nv, nq=grid.nq, ft_bytes=ft, deep=int(grid.deep), n_tracers=int(n_tracers))
    G = Geometry(dxdxi=arr(grid.dxdxi[elems], np.float64), J2=arr(grid.J2[elems], np.float64),
                 lat=arr(grid.lat[elems], np.float64), gll_w=arr(grid.wq, np.float64), gll_D=arr(grid.D, np.float64),
                 z_c=arr(grid.z_c, np.float64), z_f=arr(grid.z_f, np.float64), dz_c=arr(grid.dz_c, np.float64),
                 dz_f=arr(grid.dz_f, np.float64), radius=grid.radius, z_max=grid.z_max)
    prm = make_params(P, N, grid)
    ctx = C.c_void_p()
    idbuf = C.create_string_buffer(nccl_id, 128) if nccl_id is not None else None
    check(lib.b200_create(C.byref(ctx), C.byref(dims), C.byref(G), C.byref(T), C.byref(prm),
                          C.cast(idbuf, C.c_void_p) if idbuf is not None else None, rank, nranks), "b200_create")
    return ctx


def build_dss_csr(topo, elem_gid=None):
    """Host-only CSR build through the C-ABI (no GPU needed); returns (offsets, members[elem*16+j*4+i])."""
    lib = load()
    faces = np.ascontiguousarray(topo.interior_faces, dtype=np.int32)
    lv = np.ascontiguousarray(topo.local_vertices, dtype=np.int32)
    lvo = np.ascontiguousarray(topo.local_vertex_offset, dtype=np.int32)
    T = Topology(interior_faces=_ptr(faces), n_faces=len(faces), local_vertices=_ptr(lv), local_vertex_offset=_ptr(lvo),
                 n_verts=len(lvo) - 1, n_neighbors=0)
    if elem_gid is not None:
        gid = np.ascontiguousarray(elem_gid, dtype=np.int64)
        T.elem_gid = _ptr(gid)
    cap_n = len(lvo) - 1 + 2 * len(faces)
    cap_m = len(lv) + 4 * len(faces)
    off = np.zeros(cap_n + 1, dtype=np.int32)
    mem = np.zeros(cap_m, dtype=np.int32)
    nn, nm = C.c_int32(), C.c_int32()
    check(lib.b200_build_dss_csr(C.byref(T), _ptr(off), cap_n, _ptr(mem), cap_m, C.byref(nn), C.byref(nm)), "b200_build_dss_csr")
    return off[: nn.value + 1].copy(), mem[: nm.value].copy()


def setup_peer_halo(ctx, part, comms):
    """NVLink peer-memory halo: exchange cudaIpc handles of the ghost buffers over torch.distributed and map
    the neighbours' buffers (falls back to the NCCL send/recv halo if IPC mapping is not possible)."""
    lib = load()
    buf = C.create_string_buffer(64)
    check(lib.b200_halo_export(ctx, C.cast(buf, C.c_void_p)), "b200_halo_export")
    mine = dict(handle=buf.raw, nh_ghost=int(part.nh_ghost), nbrs=[int(x) for x in part.neighbor_ranks],
                recv_offset=[int(x) for x in part.recv_offset])
    allinfo = [None] * comms.nranks
    comms.dist.all_gather_object(allinfo, mine)
    nn = len(part.neighbor_ranks)
    handles = b"".join(allinfo[int(q)]["handle"] for q in part.neighbor_ranks)
    their_off = np.array([allinfo[int(q)]["recv_offset"][allinfo[int(q)]["nbrs"].index(comms.rank)] for q in part.neighbor_ranks], dtype=np.int32)
    their_nhg = np.array([allinfo[int(q)]["nh_ghost"] for q in part.neighbor_ranks], dtype=np.int32)
    hb = C.create_string_buffer(handles, 64 * nn)
    rc = lib.b200_halo_import(ctx, C.cast(hb, C.c_void_p), _ptr(their_off), _ptr(their_nhg))
    ok = comms.all_true(rc == 0)
    if not ok and comms.rank == 0:
        print("warning: peer-memory halo unavailable (" + lib.b200_last_error(ctx).decode() + "); using NCCL send/recv")
    return ok


def debug_dss_csr(ctx):
    lib = load()
    off, mem = C.c_void_p(), C.c_void_p()
    nn, nm = C.c_int32(), C.c_int32()
    lib.b200_debug_dss_csr(ctx, C.byref(off), C.byref(mem), C.byref(nn), C.byref(nm))
    o = np.ctypeslib.as_array(C.cast(off, C.POINTER(C.c_int32)), shape=(nn.value + 1,)).copy()
    m = np.ctypeslib.as_array(C.cast(mem, C.POINTER(C.c_int32)), shape=(nm.value,)).copy()
    return o, m
