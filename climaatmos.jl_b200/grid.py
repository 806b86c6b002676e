"""Host-side grid construction (init-time only, NumPy float64 → cast to FT).

Mirrors what the reference obtains from ``SphereGrid`` (src/simulation/grids.jl:42-90):
an equiangular cubed sphere with ``h_elem`` elements per panel edge, GLL quadrature with
``nh_poly + 1`` nodes, elements ordered along a space-filling curve (grids.jl:73-75), a
hyperbolic-tangent stretched vertical mesh (grids.jl:356-357) and deep/shallow spherical
shell geometry.  The arithmetic of those pieces lives in ClimaCore 0.15.1 (un-vendored,
[UPSTREAM-RECALL]); in a real drop-in deployment the arrays produced here are *copied from
the live ClimaCore objects* (see INTEGRATION.md).  This module exists so that the tests,
``smoke()`` and ``bench.py`` can build the same data contract without Julia.

Everything the CUDA library needs is exported as plain arrays:

* horizontal (per element ``h``, node ``(j, i)``; C order ``[h][j][i]``):
  ``dxdxi[h, j, i, a, b]`` = ∂x_a/∂ξ_b in the local (east, north) basis at radius R,
  ``J2`` (2-D Jacobian), ``W`` (w_i·w_j), ``lat``/``lon`` (degrees);
* vertical: ``z_c[Nv]``, ``z_f[Nv+1]``, ``dz_c`` (vertical J at centres), ``dz_f``
  (vertical J at faces);
* topology in ClimaCore ``Topology2D`` form: ``interior_faces`` rows
  ``(e1, f1, e2, f2, reversed)`` and ``local_vertices``/``local_vertex_offset``.
"""
from __future__ import annotations

import dataclasses
import numpy as np

# ----------------------------------------------------------------------------------------------
# GLL quadrature (Nq = 4 is the only size the hot path uses; general Nq supported for tests)
# ----------------------------------------------------------------------------------------------


def gll_points_weights(nq: int):
    """Gauss–Lobatto–Legendre nodes/weights on [-1, 1] (ClimaCore ``Quadratures.GLL{Nq}``)."""
    if nq == 2:
        return np.array([-1.0, 1.0]), np.array([1.0, 1.0])
    n = nq - 1
    # interior nodes = roots of P'_n
    pn = np.polynomial.legendre.Legendre.basis(n)
    xi = np.sort(np.real(pn.deriv().roots()))
    x = np.concatenate(([-1.0], xi, [1.0]))
    w = 2.0 / (n * (n + 1) * pn(x) ** 2)
    # symmetrise
    x = 0.5 * (x - x[::-1])
    w = 0.5 * (w + w[::-1])
    return x, w


def differentiation_matrix(x: np.ndarray) -> np.ndarray:
    """D[i, k] = ℓ'_k(ξ_i) for the Lagrange basis on nodes ``x`` (barycentric form)."""
    n = len(x)
    c = np.ones(n)
    for i in range(n):
        for k in range(n):
            if k != i:
                c[i] *= x[i] - x[k]
    D = np.zeros((n, n))
    for i in range(n):
        for k in range(n):
            if i != k:
                D[i, k] = c[i] / (c[k] * (x[i] - x[k]))
        D[i, i] = -np.sum(D[i, :])
    D[np.abs(D) < 1e-14] = 0.0
    return D


# ----------------------------------------------------------------------------------------------
# Space-filling curve (generalised Hilbert curve; per panel)
# ----------------------------------------------------------------------------------------------


def _sgn(x):
    return (x > 0) - (x < 0)


def _gilbert(x, y, ax, ay, bx, by, out):
    w = abs(ax + ay)
    h = abs(bx + by)
    dax, day = _sgn(ax), _sgn(ay)
    dbx, dby = _sgn(bx), _sgn(by)
    if h == 1:
        for _ in range(w):
            out.append((x, y))
            x, y = x + dax, y + day
        return
    if w == 1:
        for _ in range(h):
            out.append((x, y))
            x, y = x + dbx, y + dby
        return
    ax2, ay2 = ax // 2, ay // 2
    bx2, by2 = bx // 2, by // 2
    w2 = abs(ax2 + ay2)
    h2 = abs(bx2 + by2)
    if 2 * w > 3 * h:
        if (w2 % 2) and (w > 2):
            ax2, ay2 = ax2 + dax, ay2 + day
        _gilbert(x, y, ax2, ay2, bx, by, out)
        _gilbert(x + ax2, y + ay2, ax - ax2, ay - ay2, bx, by, out)
    else:
        if (h2 % 2) and (h > 2):
            bx2, by2 = bx2 + dbx, by2 + dby
        _gilbert(x, y, bx2, by2, ax2, ay2, out)
        _gilbert(x + bx2, y + by2, ax, ay, bx - bx2, by - by2, out)
        _gilbert(
            x + (ax - dax) + (bx2 - dbx),
            y + (ay - day) + (by2 - dby),
            -bx2,
            -by2,
            -(ax - ax2),
            -(ay - ay2),
            out,
        )


def spacefillingcurve(ne: int):
    """Element order ``[(ex, ey, panel), ...]``: a generalised Hilbert curve inside each panel,
    panels visited in order (role of ``Topologies.spacefillingcurve``, grids.jl:73-75)."""
    cells = []
    _gilbert(0, 0, ne, 0, 0, ne, cells)
    assert len(cells) == ne * ne and len(set(cells)) == ne * ne
    order = []
    for p in range(6):
        seq = cells if p % 2 == 0 else cells[::-1]
        order.extend((ex, ey, p) for (ex, ey) in seq)
    return order


# ----------------------------------------------------------------------------------------------
# Cubed sphere
# ----------------------------------------------------------------------------------------------

# panel p maps the gnomonic point (1, X, Y) to Cartesian via these column selections/signs
_PANEL = [
    lambda X, Y: (np.ones_like(X), X, Y),  # +x
    lambda X, Y: (-X, np.ones_like(X), Y),  # +y
    lambda X, Y: (-np.ones_like(X), -X, Y),  # -x
    lambda X, Y: (X, -np.ones_like(X), Y),  # -y
    lambda X, Y: (-Y, X, np.ones_like(X)),  # +z
    lambda X, Y: (Y, X, -np.ones_like(X)),  # -z
]


def _panel_xyz(p, X, Y):
    a, b, c = _PANEL[p](X, Y)
    d = np.sqrt(a * a + b * b + c * c)
    return a / d, b / d, c / d


def _panel_xyz_deriv(p, X, Y):
    """Unit-sphere position and its derivatives w.r.t. the gnomonic coordinates X and Y."""
    one = np.ones_like(X)
    zero = np.zeros_like(X)
    a, b, c = _PANEL[p](X, Y)
    # derivative of the un-normalised vector: linear in X, Y
    aX, bX, cX = [q1 - q0 for q1, q0 in zip(_PANEL[p](one, zero), _PANEL[p](zero, zero))]
    aY, bY, cY = [q1 - q0 for q1, q0 in zip(_PANEL[p](zero, one), _PANEL[p](zero, zero))]
    d2 = a * a + b * b + c * c
    d = np.sqrt(d2)
    r = np.stack([a, b, c], -1) / d[..., None]

    def dd(da, db, dc):
        v = np.stack([da, db, dc], -1) / d[..., None]
        dot = (a * da + b * db + c * dc) / d2
        return v - r * dot[..., None]

    return r, dd(aX, bX, cX), dd(aY, bY, cY)


@dataclasses.dataclass
class Topology2D:
    """Connectivity in the shape of ClimaCore's ``Topologies.Topology2D`` [UPSTREAM-RECALL].

    Faces are numbered 1..4 = (ξ2=-1, ξ1=+1, ξ2=+1, ξ1=-1), traversed counter-clockwise;
    vertices 1..4 = (-1,-1), (+1,-1), (+1,+1), (-1,+1).  All indices here are 0-based.
    """

    nelems: int
    elemorder: list  # [(ex, ey, panel)] in SFC order; index = element id
    interior_faces: np.ndarray  # (nfaces, 5) int32: e1, f1, e2, f2, reversed
    local_vertices: np.ndarray  # (nlv, 2) int32: elem, vert
    local_vertex_offset: np.ndarray  # (nverts + 1,) int32


@dataclasses.dataclass
class SphereGrid:
    FT: type
    h_elem: int
    nq: int
    z_elem: int
    radius: float
    z_max: float
    deep: bool
    nelems: int
    # quadrature
    xi: np.ndarray
    wq: np.ndarray
    D: np.ndarray
    # horizontal geometry [h, j, i, ...]
    dxdxi: np.ndarray
    J2: np.ndarray
    W: np.ndarray
    lat: np.ndarray
    lon: np.ndarray
    xyz: np.ndarray
    # vertical
    z_f: np.ndarray
    z_c: np.ndarray
    dz_c: np.ndarray
    dz_f: np.ndarray
    topology: Topology2D

    @property
    def nv(self):
        return self.z_elem

    @property
    def ncols(self):
        return self.nelems * self.nq * self.nq

    def node_horizontal_length_scale(self) -> float:
        """``Spaces.node_horizontal_length_scale`` (used by ν₄, hyperdiffusion.jl:21-28):
        sqrt(sphere area / n_elems) / (Nq - 1)."""
        return float(np.sqrt(4 * np.pi * self.radius**2 / (6 * self.h_elem**2)) / (self.nq - 1))


def hyperbolic_tangent_stretching(z_max: float, nelems: int, dz_bottom: float) -> np.ndarray:
    """Face heights of ``Meshes.HyperbolicTangentStretching(dz_bottom)`` on [0, z_max]
    (grids.jl:356-357) [UPSTREAM-RECALL]: η(ζ) = 1 + tanh(γ(ζ-1))/tanh(γ) with γ chosen so the
    first layer is ``dz_bottom`` thick."""
    h = dz_bottom / z_max
    zeta = np.linspace(0.0, 1.0, nelems + 1)
    if h >= 1.0 / nelems:  # no stretching possible/needed
        return z_max * zeta

    def first(gamma):
        return 1.0 + np.tanh(gamma * (1.0 / nelems - 1.0)) / np.tanh(gamma) - h

    lo, hi = 1e-8, 50.0
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if first(mid) > 0:
            lo = mid
        else:
            hi = mid
    gamma = 0.5 * (lo + hi)
    eta = 1.0 + np.tanh(gamma * (zeta - 1.0)) / np.tanh(gamma)
    eta[0], eta[-1] = 0.0, 1.0
    return z_max * eta


def uniform_faces(z_max: float, nelems: int) -> np.ndarray:
    return np.linspace(0.0, z_max, nelems + 1)


FACE_VERTS = ((0, 1), (1, 2), (2, 3), (3, 0))  # face f goes from vertex a to vertex b (ccw)


def face_nodes(f: int, nq: int):
    """(i, j) of the nq nodes of face ``f`` in ccw traversal order (vertex a → vertex b)."""
    r = range(nq)
    if f == 0:
        return [(q, 0) for q in r]
    if f == 1:
        return [(nq - 1, q) for q in r]
    if f == 2:
        return [(nq - 1 - q, nq - 1) for q in r]
    return [(0, nq - 1 - q) for q in r]


def vertex_node(v: int, nq: int):
    return ((0, 0), (nq - 1, 0), (nq - 1, nq - 1), (0, nq - 1))[v]


def perimeter_nodes(nq: int):
    """ClimaCore ``Perimeter2D(Nq)`` enumeration [UPSTREAM-RECALL]: the 4 vertices, then the
    interior nodes of faces 1..4 in ccw order. Returns [(i, j)] of length 4(nq-1)."""
    out = [vertex_node(v, nq) for v in range(4)]
    for f in range(4):
        out.extend(face_nodes(f, nq)[1:-1])
    return out


def build_topology(ne: int, elemorder) -> Topology2D:
    """Derive face/vertex connectivity geometrically from the cube (robust to panel layout)."""
    nel = len(elemorder)
    # corner key: quantised Cartesian coordinates of element corners on the cube surface
    def corner_key(ex, ey, p, cx, cy):
        X = np.tan(-np.pi / 4 + (np.pi / 2) * (ex + cx) / ne)
        Y = np.tan(-np.pi / 4 + (np.pi / 2) * (ey + cy) / ne)
        x, y, z = _panel_xyz(p, np.float64(X), np.float64(Y))
        return (int(round(float(x) * 1e7)), int(round(float(y) * 1e7)), int(round(float(z) * 1e7)))

    corner_off = ((0, 0), (1, 0), (1, 1), (0, 1))
    vkeys = {}
    elem_vid = np.zeros((nel, 4), dtype=np.int64)
    for e, (ex, ey, p) in enumerate(elemorder):
        for v, (cx, cy) in enumerate(corner_off):
            k = corner_key(ex, ey, p, cx, cy)
            if k not in vkeys:
                vkeys[k] = len(vkeys)
            elem_vid[e, v] = vkeys[k]
    nverts = len(vkeys)
    assert nverts == 6 * ne * ne + 2
    # vertices: group (elem, vert) by unique vertex, ordered by first appearance (elem asc)
    members = [[] for _ in range(nverts)]
    for e in range(nel):
        for v in range(4):
            members[elem_vid[e, v]].append((e, v))
    offs = [0]
    lv = []
    for m in members:
        lv.extend(m)
        offs.append(len(lv))
    # faces
    edges = {}
    faces = []
    for e in range(nel):
        for f, (a, b) in enumerate(FACE_VERTS):
            va, vb = elem_vid[e, a], elem_vid[e, b]
            key = (min(va, vb), max(va, vb))
            if key in edges:
                e1, f1, va1, vb1 = edges.pop(key)
                # reversed: traversing face f of e (a→b) runs opposite to face f1 of e1
                rev = 1 if (va1 == vb and vb1 == va) else 0
                faces.append((e1, f1, e, f, rev))
            else:
                edges[key] = (e, f, va, vb)
    assert not edges, "cubed sphere has no boundary faces"
    return Topology2D(
        nelems=nel,
        elemorder=list(elemorder),
        interior_faces=np.asarray(faces, dtype=np.int32).reshape(-1, 5),
        local_vertices=np.asarray(lv, dtype=np.int32).reshape(-1, 2),
        local_vertex_offset=np.asarray(offs, dtype=np.int32),
    )


def make_sphere_grid(
    FT=np.float32,
    h_elem: int = 6,
    nh_poly: int = 3,
    z_elem: int = 10,
    z_max: float = 30000.0,
    dz_bottom: float = 500.0,
    z_stretch: bool = True,
    radius: float = 6.371229e6,
    deep_atmosphere: bool = True,
) -> SphereGrid:
    """Keyword surface of ``SphereGrid`` (src/simulation/grids.jl:42-62)."""
    nq = nh_poly + 1
    ne = h_elem
    xi, wq = gll_points_weights(nq)
    D = differentiation_matrix(xi)
    order = spacefillingcurve(ne)
    nel = len(order)

    ex = np.array([o[0] for o in order])
    ey = np.array([o[1] for o in order])
    pn = np.array([o[2] for o in order])
    # angles of every node [h, j, i]
    dalpha = (np.pi / 2) / ne
    alpha = -np.pi / 4 + dalpha * (ex[:, None, None] + 0.5 * (xi[None, None, :] + 1.0))
    beta = -np.pi / 4 + dalpha * (ey[:, None, None] + 0.5 * (xi[None, :, None] + 1.0))
    alpha = np.broadcast_to(alpha, (nel, nq, nq)).copy()
    beta = np.broadcast_to(beta, (nel, nq, nq)).copy()
    X, Y = np.tan(alpha), np.tan(beta)
    # exact zeros at panel centres so pole nodes agree across elements
    X[np.abs(alpha) < 1e-14] = 0.0
    Y[np.abs(beta) < 1e-14] = 0.0
    r = np.zeros((nel, nq, nq, 3))
    rX = np.zeros_like(r)
    rY = np.zeros_like(r)
    for p in range(6):
        m = pn == p
        r[m], rX[m], rY[m] = _panel_xyz_deriv(p, X[m], Y[m])
    # chain rule: dX/dξ1 = (1 + X²)·dα/dξ1, dα/dξ1 = dalpha/2
    r1 = rX * ((1 + X**2) * (dalpha / 2))[..., None]
    r2 = rY * ((1 + Y**2) * (dalpha / 2))[..., None]
    x, y, z = r[..., 0], r[..., 1], r[..., 2]
    x = np.where(np.abs(x) < 1e-15, 0.0, x)
    y = np.where(np.abs(y) < 1e-15, 0.0, y)
    lon = np.arctan2(y, x)
    lat = np.arctan2(z, np.hypot(x, y))
    east = np.stack([-np.sin(lon), np.cos(lon), np.zeros_like(lon)], -1)
    north = np.stack([-np.sin(lat) * np.cos(lon), -np.sin(lat) * np.sin(lon), np.cos(lat)], -1)
    dxdxi = np.zeros((nel, nq, nq, 2, 2))
    dxdxi[..., 0, 0] = radius * np.sum(r1 * east, -1)
    dxdxi[..., 1, 0] = radius * np.sum(r1 * north, -1)
    dxdxi[..., 0, 1] = radius * np.sum(r2 * east, -1)
    dxdxi[..., 1, 1] = radius * np.sum(r2 * north, -1)
    J2 = dxdxi[..., 0, 0] * dxdxi[..., 1, 1] - dxdxi[..., 0, 1] * dxdxi[..., 1, 0]
    assert np.all(J2 > 0)
    W = np.broadcast_to(wq[None, :, None] * wq[None, None, :], (nel, nq, nq)).copy()

    z_f = (
        hyperbolic_tangent_stretching(z_max, z_elem, dz_bottom)
        if z_stretch
        else uniform_faces(z_max, z_elem)
    )
    z_c = 0.5 * (z_f[1:] + z_f[:-1])
    dz_c = z_f[1:] - z_f[:-1]
    # face "J": centre-to-centre spacing in the interior, 2 × half-cell at the boundaries
    dz_f = np.empty(z_elem + 1)
    dz_f[1:-1] = z_c[1:] - z_c[:-1]
    dz_f[0] = 2.0 * (z_c[0] - z_f[0])
    dz_f[-1] = 2.0 * (z_f[-1] - z_c[-1])

    topo = build_topology(ne, order)
    return SphereGrid(
        FT=FT,
        h_elem=h_elem,
        nq=nq,
        z_elem=z_elem,
        radius=radius,
        z_max=z_max,
        deep=deep_atmosphere,
        nelems=nel,
        xi=xi,
        wq=wq,
        D=D,
        dxdxi=dxdxi,
        J2=J2,
        W=W,
        lat=np.degrees(lat),
        lon=np.degrees(lon),
        xyz=r,
        z_f=z_f,
        z_c=z_c,
        dz_c=dz_c,
        dz_f=dz_f,
        topology=topo,
    )


# ----------------------------------------------------------------------------------------------
# DSS node map (the index tables the CUDA gather kernel consumes)
# ----------------------------------------------------------------------------------------------


def dss_node_csr(topo: Topology2D, nq: int):
    """Build the unique-perimeter-node → [(elem, i, j)] CSR from the ClimaCore-shaped tables.

    Ordering contract (bit-exact, checked against the C++ builder in tests): vertices first in
    ``local_vertex_offset`` order, members in ``local_vertices`` order; then for each interior
    face in table order, its nq-2 interior nodes in the traversal order of element 1, members
    (e1, e2).  Returns (offsets int32[n+1], members int32[m, 3] = elem, i, j).
    """
    offs = [0]
    mem = []
    lv, lo = topo.local_vertices, topo.local_vertex_offset
    for v in range(len(lo) - 1):
        for q in range(lo[v], lo[v + 1]):
            e, vert = lv[q]
            i, j = vertex_node(int(vert), nq)
            mem.append((int(e), i, j))
        offs.append(len(mem))
    for e1, f1, e2, f2, rev in topo.interior_faces:
        n1 = face_nodes(int(f1), nq)
        n2 = face_nodes(int(f2), nq)
        for q in range(1, nq - 1):
            q2 = nq - 1 - q if rev else q
            mem.append((int(e1), *n1[q]))
            mem.append((int(e2), *n2[q2]))
            offs.append(len(mem))
    return np.asarray(offs, dtype=np.int32), np.asarray(mem, dtype=np.int32).reshape(-1, 3)
