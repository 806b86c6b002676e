"""CPU tests: grid construction, quadrature, topology, SFC, partition and stepper host logic."""
import re
import os
import numpy as np
import pytest

from climaatmos_jl_b200 import grid as G
from climaatmos_jl_b200 import partition, params, dycore


def test_gll_quadrature_and_derivative_matrix():
    x, w = G.gll_points_weights(4)
    assert np.allclose(x, [-1, -1 / np.sqrt(5), 1 / np.sqrt(5), 1])
    assert np.allclose(w, [1 / 6, 5 / 6, 5 / 6, 1 / 6])
    D = G.differentiation_matrix(x)
    # exact for polynomials up to degree 3; rows sum to zero
    for p in range(4):
        assert np.allclose(D @ x**p, p * x ** max(p - 1, 0) if p else 0 * x, atol=1e-13)
    assert np.abs(D.sum(axis=1)).max() < 1e-14


@pytest.mark.parametrize("ne", [2, 3, 6])
def test_cubed_sphere_area_and_counts(ne):
    g = G.make_sphere_grid(h_elem=ne, z_elem=4)
    assert g.nelems == 6 * ne * ne
    area = np.sum(g.J2 * g.W)
    assert abs(area / (4 * np.pi * g.radius**2) - 1) < 5e-4 / ne**4 + 1e-6
    t = g.topology
    assert len(t.interior_faces) == 12 * ne * ne  # every element has 4 faces, each shared by 2
    assert len(t.local_vertex_offset) - 1 == 6 * ne * ne + 2
    counts = np.diff(t.local_vertex_offset)
    assert sorted(np.unique(counts)) == [3, 4] and (counts == 3).sum() == 8  # cube corners
    # all panels right-handed ⇒ shared faces are always traversed in opposite directions
    assert np.all(t.interior_faces[:, 4] == 1)


def test_collocated_nodes_coincide_in_space():
    g = G.make_sphere_grid(h_elem=3, z_elem=4)
    off, mem = G.dss_node_csr(g.topology, 4)
    for n in range(len(off) - 1):
        pts = np.array([g.xyz[e, j, i] for e, i, j in mem[off[n]:off[n + 1]]])
        assert np.abs(pts - pts[0]).max() < 1e-12
    # every perimeter node appears exactly once
    keys = mem[:, 0] * 16 + mem[:, 2] * 4 + mem[:, 1]
    assert len(np.unique(keys)) == len(keys) == g.nelems * 12


def test_spacefillingcurve_selfconsistent():
    """Mirror of the reference's SFC test (test/grids.jl:6-21)."""
    order = G.spacefillingcurve(3)
    index = {c: k for k, c in enumerate(order)}
    for k, c in enumerate(order):
        assert index[c] == k
    linear = [(x, y, p) for p in range(6) for y in range(3) for x in range(3)]
    assert sorted(order) == sorted(linear) and order != linear
    # consecutive elements inside a panel are face neighbours (locality of the curve)
    for a, b in zip(order[:8], order[1:9]):
        assert abs(a[0] - b[0]) + abs(a[1] - b[1]) == 1


def test_vertical_stretching():
    zf = G.hyperbolic_tangent_stretching(60000.0, 63, 30.0)
    assert zf[0] == 0 and zf[-1] == 60000.0 and abs((zf[1] - zf[0]) - 30.0) < 1e-6
    assert np.all(np.diff(zf) > 0) and np.all(np.diff(np.diff(zf)) > -1e-9)
    g = G.make_sphere_grid(h_elem=2, z_elem=10)
    assert np.allclose(g.dz_f[1:-1], np.diff(g.z_c)) and np.isclose(g.dz_f[0], g.dz_c[0])


def test_ars343_tableau():
    a_exp, a_imp, b_exp, b_imp, gam = params.ars343()
    c = [0, gam, (1 + gam) / 2, 1]
    for i in range(4):
        assert abs(sum(a_exp[i]) - c[i]) < 1e-14 and abs(sum(a_imp[i]) - c[i]) < 1e-14
    assert abs(sum(b_exp) - 1) < 1e-14 and b_exp == b_imp
    assert a_imp[3] == b_imp  # stiffly accurate


def test_sypd_definition():
    # solve.jl:40-45: 365-day years per wall-clock day
    assert np.isclose(dycore.sypd(365 * 86400.0, 86400.0), 1.0)


@pytest.mark.parametrize("nranks", [2, 3, 8])
def test_partition_is_consistent(nranks):
    g = G.make_sphere_grid(h_elem=4, z_elem=4)
    parts = [partition.partition_grid(g, r, nranks) for r in range(nranks)]
    assert sum(p.nh for p in parts) == g.nelems
    assert max(p.nh for p in parts) - min(p.nh for p in parts) <= 1
    owned = np.concatenate([p.elems_ext[: p.nh] for p in parts])
    assert np.array_equal(np.sort(owned), np.arange(g.nelems))
    for p in parts:
        for k, q in enumerate(p.neighbor_ranks):
            s = p.elems_ext[p.send_elems[p.send_offset[k]:p.send_offset[k + 1]]]
            pq = parts[q]
            kk = list(pq.neighbor_ranks).index(p.rank)
            r = pq.elems_ext[pq.nh + pq.recv_offset[kk]: pq.nh + pq.recv_offset[kk + 1]]
            assert np.array_equal(s, r)
        # every node with a local member keeps all of its members
        off, mem = G.dss_node_csr(G.Topology2D(p.nh + p.nh_ghost, [], p.interior_faces, p.local_vertices, p.local_vertex_offset), 4)
        goff, gmem = G.dss_node_csr(g.topology, 4)
        glob = {}
        for n in range(len(goff) - 1):
            m = frozenset((int(e), int(i), int(j)) for e, i, j in gmem[goff[n]:goff[n + 1]])
            for x in m:
                glob[x] = m
        for n in range(len(off) - 1):
            m = frozenset((int(p.elems_ext[e]), int(i), int(j)) for e, i, j in mem[off[n]:off[n + 1]])
            if any(e < p.nh for e, _, _ in mem[off[n]:off[n + 1]]):
                assert m == glob[next(iter(m))]


def test_header_symbols_exported(lib):
    """The C-ABI library loads and exports every function include/b200_dycore.h declares."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "b200_dycore.h")).read()
    names = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n


@pytest.mark.parametrize("ne", [2, 5, 6])
def test_dss_index_map_bit_exact(lib, ne):
    """DSS element/node index maps built by the library from ClimaCore-shaped Topology2D tables are
    bit-identical to the independent Python construction."""
    from climaatmos_jl_b200 import capi

    g = G.make_sphere_grid(h_elem=ne, z_elem=4)
    off, mem = capi.build_dss_csr(g.topology)
    o2, m2 = G.dss_node_csr(g.topology, 4)
    assert np.array_equal(off, o2)
    assert np.array_equal(mem, m2[:, 0] * 16 + m2[:, 2] * 4 + m2[:, 1])


def test_create_without_gpu_fails_loudly(lib):
    """No CPU fallback: creating a context without a CUDA device is an error, not a silent path."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from climaatmos_jl_b200 import capi

    g = G.make_sphere_grid(h_elem=2, z_elem=4)
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        capi.create_context(g, params.DycoreParams(), params.DycoreNumerics())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dycore.AtmosSimulation(h_elem=2, z_elem=4)


def test_create_validates_parameters_before_touching_a_device(lib):
    """Configuration errors are reported by b200_create itself (no device needed): implicit_diffusion without a vert_diff model — the
    reference's update_diffusion_jacobian! has no diffusivity then — and more tracers than the documented maximum."""
    from climaatmos_jl_b200 import capi

    g = G.make_sphere_grid(h_elem=2, z_elem=4)
    with pytest.raises(RuntimeError, match="implicit_diffusion needs a vert_diff model"):
        capi.create_context(g, params.DycoreParams(), params.DycoreNumerics(implicit_diffusion=True))
    with pytest.raises(RuntimeError, match="n_tracers"):
        capi.create_context(g, params.DycoreParams(), params.DycoreNumerics(), n_tracers=5)


def test_entry_points_reject_a_null_context(lib):
    """Error convention of the C-ABI (include/b200_dycore.h): <0 and a message via b200_last_error, nothing crosses as a crash —
    the glue turns it into the exception solve_atmos! catches (src/simulation/solve.jl:140-149).  No device is touched."""
    import ctypes as C
    from climaatmos_jl_b200 import capi

    lib = C.CDLL(capi.LIB_PATH)  # a handle without the argtypes of capi.load(): every argument is passed as a raw pointer / scalar
    z = C.c_void_p(None)
    calls = {
        "b200_cache_imp": (z, z, z, z, z), "b200_t_imp": (z, z, z, z, z, C.c_double(0), z),
        "b200_wfact": (z, z, z, C.c_double(1), C.c_double(0), z), "b200_ldiv": (z, z, z, z, z, z),
        "b200_t_post_imp": (z, z, z, z, z, C.c_double(0), z), "b200_t_exp_lim": (z, z, z, z, z, z, z, C.c_double(0), z),
        "b200_dss": (z, z, z, z, z, C.c_int32(0), z), "b200_axpy_n": (z, z, z, z, z, C.c_int32(0), z, z, z, z),
        "b200_lim": (z, z, z, z, z, C.c_double(0), z), "b200_implicit_stage": (z, z, z, z, z, C.c_double(1), z),
        "b200_step_ars343": (z, z, z, C.c_double(0), C.c_int32(1), z), "b200_halo_export": (z, z), "b200_halo_import": (z, z, z, z),
    }
    lib.b200_last_error.restype = C.c_char_p
    for name, args in calls.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        assert fn(*args) < 0, name
        assert name.encode() in lib.b200_last_error(z) and b"null context" in lib.b200_last_error(z)
    assert lib.b200_destroy(z) == 0


def test_ctypes_mirrors_match_the_c_structs(tmp_path):
    """The drop-in boundary is plain C structs: the ctypes mirrors in capi.py (what the tests and bench.py bind through, and the
    template for the Julia `struct` mirrors of INTEGRATION.md) must have exactly the size and field offsets gcc gives the structs of
    include/b200_dycore.h."""
    import ctypes as C
    import re
    import subprocess

    from climaatmos_jl_b200 import capi

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pairs = {"b200_dims": capi.Dims, "b200_geometry": capi.Geometry, "b200_topology": capi.Topology, "b200_params": capi.Params,
             "b200_cacheptrs": capi.CachePtrs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "b200_dycore.h"', "int main(void) {"]
    for cname, T in pairs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in T._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    got = dict(re.findall(r"(\S+) (\d+)", out))
    for cname, T in pairs.items():
        assert int(got[cname]) == C.sizeof(T), cname
        for fname, _ in T._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(T, fname).offset, f"{cname}.{fname}"


def test_julia_glue_is_structurally_complete():
    """julia/B200Dycore.jl cannot be executed here (no julia in the image or on the GPU box), so it is checked structurally: no bodiless
    `function … end` stubs, block keywords balance, every symbol it `ccall`s is exported by the library, every ccall passes exactly
    as many argument types and values as the C prototype in include/b200_dycore.h has parameters, and the struct mirrors list the same
    field names in the same order as the C structs."""
    import re

    from climaatmos_jl_b200 import capi

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "julia", "B200Dycore.jl"), encoding="utf-8").read()
    hdr = open(os.path.join(root, "include", "b200_dycore.h"), encoding="utf-8").read()
    assert not re.search(r"^\s*function\s+[\w!.]+\s+end\s*$", src, flags=re.M), "bodiless function stub"
    # strip docstrings, strings and comments, then balance block openers against `end`
    code = re.sub(r'"""(.|\n)*?"""', '""', src)
    code = re.sub(r'"(\\.|[^"\\\n])*"', '""', code)
    code = re.sub(r"#.*", "", code)
    openers = len(re.findall(r"(?m)^\s*(?:function|struct|mutable struct|module|if|for|while|let|try)\b", code))
    openers += len(re.findall(r"\bbegin\s*$", code, flags=re.M)) + len(re.findall(r"\bdo\b[^\n]*$", code, flags=re.M))
    ends = len(re.findall(r"(?m)^\s*end\b", code))
    assert openers == ends, (openers, ends)
    # C prototypes: name → number of parameters
    protos = {m.group(1): (0 if m.group(2).strip() in ("", "void") else m.group(2).count(",") + 1)
              for m in re.finditer(r"\b(b200_\w+)\s*\(([^;{]*?)\)\s*;", re.sub(r"/\*(.|\n)*?\*/", "", hdr))}
    calls = re.findall(r"ccall\(\(:(b200_\w+), lib\), \w+,\s*\(([^()]*(?:\{[^{}]*\}[^()]*)*)\)", src)
    assert len(calls) >= 14
    for name, types in calls:
        assert name in capi.SYMBOLS, name
        ntypes = len([t for t in types.split(",") if t.strip()])
        assert ntypes == protos[name], (name, ntypes, protos[name])
    # struct mirrors: same field order as the C structs
    def c_fields(cname):
        h = re.sub(r"/\*(.|\n)*?\*/", "", hdr)
        end = re.search(r"\}\s*" + cname + r"\s*;", h).start()
        body = h[h.rindex("typedef struct {", 0, end) + len("typedef struct {"):end]
        names = []
        for stmt in body.split(";"):
            stmt = stmt.strip()
            if not stmt:
                continue
            for part in stmt.split(","):
                names.append(re.findall(r"(\w+)\s*$", part.strip())[0])
        return names

    def jl_fields(jname):
        body = re.search(r"struct " + jname + r"\n(.*?)\nend", src, flags=re.S).group(1)
        body = re.sub(r"#.*", "", body)
        return re.findall(r"(\w+)::", body)

    for cname, jname in (("b200_dims", "Dims"), ("b200_geometry", "GeometryC"), ("b200_topology", "TopologyC"), ("b200_params", "ParamsC"),
                         ("b200_cacheptrs", "CachePtrs")):
        assert c_fields(cname) == jl_fields(jname), (cname, c_fields(cname), jl_fields(jname))


def test_bench_measurement_model_matches_the_survey_numbers():
    """bench.py's host-side pieces (no GPU): the weak series of SURVEY.md §8d.4 (h_elem 30/42/60/85 at N = 1/2/4/8, dt ∝ 1/h_elem), the
    he30-equivalent SYPD factor (1 at N = 1; value_N/(N·value_1) is then the efficiency normalised by elements per GPU), the byte model
    54.5 S + 14 H (7.17 GB per dry he30/ze63 step, 8.66 GB with one tracer, 2.04 GB at he16 — the figures of §8d), and the NUMA probe's
    (node, diagnosis) contract when no GPU is visible."""
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    hs = {n: bench.workload(n)["h_elem"] for n in (1, 2, 4, 8)}
    assert hs == {1: 30, 2: 42, 4: 60, 8: 85}
    for n, h in hs.items():
        w = bench.workload(n)
        assert w["z_elem"] == 63 and w["scaling"] == "weak" and w["dt"] == float(round(90.0 * 30 / h))
        eq = bench.equiv_factor(w, 6 * h * h)
        assert abs(eq - (90.0 / w["dt"]) * (6 * h * h / 5400.0)) < 1e-12
    assert bench.equiv_factor(bench.workload(1), 5400) == 1.0
    assert bench.equiv_factor(bench.workload(8, "strong"), 21600) == 1.0 and bench.workload(8, "strong")["h_elem"] == 60
    assert abs(bench.model_bytes_per_step(86400, 63) / 1e9 - 7.17) < 0.01
    assert abs(bench.model_bytes_per_step(86400, 63, k=1) / 1e9 - 8.66) < 0.01
    assert abs(bench.model_bytes_per_step(24576, 63) / 1e9 - 2.04) < 0.01
    m = bench.workload(1, "moist")
    assert m["moist"] and m["ic"] == "MoistBaroclinicWave" and m["h_elem"] == 30
    node, diag = bench.pin_to_gpu_numa_node(0)
    assert (node is None or isinstance(node, int)) and isinstance(diag, str) and diag
