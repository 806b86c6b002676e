"""Horizontal domain decomposition along the space-filling curve (host logic, init time).

Mirrors how the reference shards the path: elements in space-filling-curve order are split into
contiguous, near-equal ranges, one per rank (``Topologies.Topology2D(context, mesh, elemorder)``,
src/simulation/grids.jl:73-75; docs/src/gpu_and_mpi.md:82-93) [UPSTREAM-RECALL for the exact split].
Columns are never split.  The only exchange step is the DSS halo: every rank receives the element
slabs of its *ghost* elements (elements of other ranks sharing a vertex with a local one) and then
sums collocated nodes in ascending global-element order, which makes the result independent of the
rank count.
"""
from __future__ import annotations

import dataclasses

import numpy as np


@dataclasses.dataclass
class Partition:
    rank: int
    nranks: int
    nh: int
    nh_ghost: int
    elems_ext: np.ndarray  # global ids: local elements (ascending) then ghosts (by owner rank, ascending gid)
    interior_faces: np.ndarray  # local-index copies of the Topology2D tables restricted to this rank
    local_vertices: np.ndarray
    local_vertex_offset: np.ndarray
    neighbor_ranks: np.ndarray
    send_offset: np.ndarray
    send_elems: np.ndarray  # local indices
    recv_offset: np.ndarray  # ghost-slot ranges per neighbour


def rank_ranges(nelems: int, nranks: int):
    """Contiguous near-equal split of the SFC order: first (nelems % nranks) ranks get one extra."""
    base, extra = divmod(nelems, nranks)
    starts = [r * base + min(r, extra) for r in range(nranks + 1)]
    return starts


def owner_of(gid, starts):
    return int(np.searchsorted(starts, gid, side="right") - 1)


def partition_grid(grid, rank: int, nranks: int) -> Partition:
    topo = grid.topology
    nel = grid.nelems
    starts = rank_ranges(nel, nranks)
    lo, hi = starts[rank], starts[rank + 1]
    lv, lvo = topo.local_vertices, topo.local_vertex_offset
    is_local = lambda e: lo <= e < hi
    # vertex neighbours → ghosts
    ghosts = set()
    keep_verts = []
    for v in range(len(lvo) - 1):
        mem = lv[lvo[v]:lvo[v + 1], 0]
        if any(is_local(e) for e in mem):
            keep_verts.append(v)
            ghosts.update(int(e) for e in mem if not is_local(e))
    ghost_list = sorted(ghosts, key=lambda e: (owner_of(e, starts), e))
    owners = [owner_of(e, starts) for e in ghost_list]
    nbrs = sorted(set(owners))
    recv_offset = [0]
    for r in nbrs:
        recv_offset.append(recv_offset[-1] + owners.count(r))
    elems_ext = np.array(list(range(lo, hi)) + ghost_list, dtype=np.int64)
    g2l = {int(g): k for k, g in enumerate(elems_ext)}
    # send lists: my local elements that touch (by vertex) any local element of rank r, ascending gid
    send = {r: set() for r in nbrs}
    for v in keep_verts:
        mem = [int(e) for e in lv[lvo[v]:lvo[v + 1], 0]]
        mine = [e for e in mem if is_local(e)]
        for e in mem:
            if not is_local(e):
                send[owner_of(e, starts)].update(mine)
    send_offset, send_elems = [0], []
    for r in nbrs:
        s = sorted(send[r])
        send_elems.extend(g2l[e] for e in s)
        send_offset.append(len(send_elems))
    # restricted tables in local numbering
    faces = [(g2l[int(e1)], f1, g2l[int(e2)], f2, rev) for e1, f1, e2, f2, rev in topo.interior_faces
             if (is_local(e1) or is_local(e2))]
    lv2, off2 = [], [0]
    for v in keep_verts:
        for e, vert in lv[lvo[v]:lvo[v + 1]]:
            lv2.append((g2l[int(e)], int(vert)))
        off2.append(len(lv2))
    return Partition(
        rank=rank, nranks=nranks, nh=hi - lo, nh_ghost=len(ghost_list), elems_ext=elems_ext,
        interior_faces=np.asarray(faces, dtype=np.int32).reshape(-1, 5),
        local_vertices=np.asarray(lv2, dtype=np.int32).reshape(-1, 2),
        local_vertex_offset=np.asarray(off2, dtype=np.int32),
        neighbor_ranks=np.asarray(nbrs, dtype=np.int32), send_offset=np.asarray(send_offset, dtype=np.int32),
        send_elems=np.asarray(send_elems, dtype=np.int32), recv_offset=np.asarray(recv_offset, dtype=np.int32))
