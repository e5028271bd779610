"""Mesh decomposition for multi-GPU runs: METIS partition + owned/halo renumbering.

Restates the reference's mesh splitter (class ``Decomp``, ``src/methods/decomp.cpp:69-335``) so that
the integer maps are IDENTICAL to the ``mesh/mesh.NNNN.proc`` files it writes:

  * dual graph from ``Cell::neigh`` (neigh >= 0 only), CSR xadj/adjncy                (:86-102)
  * ``METIS_PartGraphRecursive(n, ncon=1, xadj, adjncy, NULL.., nparts, NULL.., part)`` with the
    METIS 5.1.0 the reference bundles (32-bit idx_t, default options => seed 4321)      (:104)
  * per rank: owned cells = ascending global id; halo = non-owned face neighbours in discovery
    order (cell ascending, neigh slot 0..2; a cell touching two owned cells appears TWICE, exactly
    as in the reference), then stably sorted by owner rank                             (:165-208)
  * lCells = global -> LAST local occurrence; recvCount per owner                      (:209-214)
  * edges: every edge touching an owned cell, ascending global id                      (:217-236)
  * sendInd[q][p] = owner-side local ids in the receiver p's halo order                (:284-292)

What is NOT taken from the .proc convention: the reader of those files (``Grid::readMeshFiles``,
``src/mesh/grid.cpp:362-411``) re-derives c1/c2 and flips normals per rank.  Here every rank keeps
the GLOBAL orientation, normal and Gauss-point order of each edge, so the
multi-GPU result equals the single-GPU result bit for bit.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import os

import numpy as np

from . import mesh as _mesh

HERE = os.path.dirname(os.path.abspath(__file__))
METIS_LIB = os.path.join(HERE, "third_party", "_metis", "libmetis.so")


def dual_graph(m: _mesh.Mesh):
    """xadj/adjncy exactly as decomp.cpp:86-102 (neighbour slots in Cell::neigh order)."""
    nb = m.cell_neigh
    ok = nb >= 0
    xadj = np.zeros(m.nc + 1, dtype=np.int32)
    xadj[1:] = np.cumsum(ok.sum(axis=1))
    adjncy = nb[ok].astype(np.int32)          # row-major boolean indexing keeps (cell, slot) order
    return xadj, adjncy


def metis_part(m: _mesh.Mesh, nparts: int) -> np.ndarray:
    """part[] from the reference's own METIS call (decomp.cpp:104)."""
    if nparts <= 1:
        return np.zeros(m.nc, dtype=np.int32)
    if not os.path.exists(METIS_LIB):
        raise RuntimeError(f"{METIS_LIB} missing: run cfd-2d_b200/third_party/build_metis.sh "
                           "(compiles the METIS 5.1.0 bundled with the reference)")
    lib = C.CDLL(METIS_LIB, mode=os.RTLD_LOCAL)
    xadj, adjncy = dual_graph(m)
    n = C.c_int32(m.nc)
    ncon = C.c_int32(1)
    npart = C.c_int32(nparts)
    objval = C.c_int32(0)
    part = np.zeros(m.nc, dtype=np.int32)
    ip = C.POINTER(C.c_int32)
    rc = lib.METIS_PartGraphRecursive(C.byref(n), C.byref(ncon), xadj.ctypes.data_as(ip), adjncy.ctypes.data_as(ip),
                                      None, None, None, C.byref(npart), None, None, None, C.byref(objval),
                                      part.ctypes.data_as(ip))
    if rc != 1:  # METIS_OK
        raise RuntimeError(f"METIS_PartGraphRecursive failed: {rc}")
    return part


def slab_part(m: _mesh.Mesh, nparts: int) -> np.ndarray:
    """Geometric fallback partition (equal-count slabs along x) for meshes too large for the
    32-bit METIS build or when a reproducible strip layout is wanted (weak-scaling bench)."""
    order = np.argsort(m.cell_cx, kind="stable")
    part = np.empty(m.nc, dtype=np.int32)
    bounds = (np.arange(nparts + 1, dtype=np.int64) * m.nc) // nparts
    for p in range(nparts):
        part[order[bounds[p]:bounds[p + 1]]] = p
    return part


@dataclasses.dataclass
class RankMesh:
    """One rank's share: Decomp's ProcMesh (decomp.cpp:17-28) plus the flattened local mesh."""
    rank: int
    nranks: int
    nc: int                    # cCount
    nc_ex: int                 # cCountEx
    g_cells: np.ndarray        # [nc_ex] global cell ids (gCells)
    g_edges: np.ndarray        # [ne]    global edge ids (gEdges[:eCount])
    ne_ex: int                 # eCountEx (the .proc file also lists halo-only edges)
    g_edges_ex: np.ndarray     # [ne_ex - ne] global ids of the halo-only edges
    recv_count: np.ndarray     # [nranks]
    send_ind: list             # per destination rank: owned local ids
    local: dict                # arrays for fvm.Packed / cfd2d_mesh

    @property
    def send_count(self):
        return np.array([len(s) for s in self.send_ind], dtype=np.int32)

    def halo_dict(self, nccl_unique_id: bytes):
        flat = np.concatenate([np.asarray(s, dtype=np.int32) for s in self.send_ind]) if self.nranks else np.empty(0, np.int32)
        return dict(rank=self.rank, nranks=self.nranks, recv_count=self.recv_count, send_count=self.send_count,
                    send_ind=flat, nccl_unique_id=nccl_unique_id, cell_gid=self.g_cells)


def decompose(m: _mesh.Mesh, part: np.ndarray, nranks: int, only_rank: int | None = None) -> list[RankMesh]:
    """Owned/halo renumbering of every rank (or of ``only_rank`` plus the send lists it needs)."""
    part = np.asarray(part, dtype=np.int32)
    rms: dict[int, RankMesh] = {}
    lcells: dict[int, np.ndarray] = {}
    for p in range(nranks):
        owned = np.nonzero(part == p)[0].astype(np.int32)
        nb = m.cell_neigh[owned].ravel()                         # discovery order (cell, slot)
        nb = nb[nb >= 0]
        halo = nb[part[nb] != p]
        halo = halo[np.argsort(part[halo], kind="stable")]       # bubble sort by owner == stable sort
        g_cells = np.concatenate([owned, halo]).astype(np.int32)
        nc, nc_ex = owned.shape[0], g_cells.shape[0]
        # lCells[g] = i for i ascending => the LAST occurrence wins (std::map assignment, :209-211)
        l = np.full(m.nc, -1, dtype=np.int32)
        u, idx = np.unique(g_cells[::-1], return_index=True)
        l[u] = (nc_ex - 1 - idx).astype(np.int32)
        recv_count = np.bincount(part[halo], minlength=nranks).astype(np.int32)
        flag_in = np.zeros(m.ne, dtype=bool)
        flag_in[m.cell_edges[owned].ravel()] = True
        flag_ex = np.zeros(m.ne, dtype=bool)
        if halo.size:
            flag_ex[m.cell_edges[halo].ravel()] = True
        g_edges = np.nonzero(flag_in)[0].astype(np.int32)
        g_edges_ex = np.nonzero(flag_ex & ~flag_in)[0].astype(np.int32)
        lcells[p] = l
        rms[p] = RankMesh(rank=p, nranks=nranks, nc=nc, nc_ex=nc_ex, g_cells=g_cells, g_edges=g_edges,
                          ne_ex=g_edges.shape[0] + g_edges_ex.shape[0], g_edges_ex=g_edges_ex,
                          recv_count=recv_count, send_ind=[[] for _ in range(nranks)], local={})
    # second pass (decomp.cpp:284-292): owner-side send lists in the receiver's halo order
    for p in range(nranks):
        rm = rms[p]
        halo = rm.g_cells[rm.nc:]
        owner = part[halo]
        for q in np.unique(owner):
            rms[int(q)].send_ind[p] = lcells[int(q)][halo[owner == q]].astype(np.int32)
    for p in range(nranks):
        rms[p].send_ind = [np.asarray(s, dtype=np.int32) for s in rms[p].send_ind]
        if only_rank is None or p == only_rank:
            rms[p].local = _local_mesh(m, rms[p], lcells[p])
    return [rms[p] for p in range(nranks)]


def _local_mesh(m: _mesh.Mesh, rm: RankMesh, l: np.ndarray) -> dict:
    gc, ge = rm.g_cells, rm.g_edges
    ledge = np.full(m.ne, -1, dtype=np.int32)
    ledge[ge] = np.arange(ge.shape[0], dtype=np.int32)
    c2g = m.edge_c2[ge]
    loc = dict(
        cell_S=m.cell_S[gc], cell_cx=m.cell_cx[gc], cell_cy=m.cell_cy[gc], cell_mat=m.cell_mat[gc],
        cell_edges=ledge[m.cell_edges[gc[:rm.nc]]],
        edge_c1=l[m.edge_c1[ge]], edge_c2=np.where(c2g >= 0, l[np.maximum(c2g, 0)], -1).astype(np.int32),
        edge_nx=m.edge_nx[ge], edge_ny=m.edge_ny[ge], edge_l=m.edge_l[ge], edge_gp=m.edge_gp[ge],
        edge_bc=m.edge_bc[ge])
    assert (loc["cell_edges"] >= 0).all() and (loc["edge_c1"] >= 0).all()
    assert ((loc["edge_c2"] >= 0) == (c2g >= 0)).all()
    return loc


def slab_rank_mesh(nx: int, ny: int, rank: int, world: int, h: float = 1.0):
    """The share of ``rank`` in the slab partition of the (world*nx) x ny channel, built WITHOUT the
    global mesh: the rank's nx columns plus one column of quads on each interior side, decomposed
    three ways (left pad | mine | right pad).  Cells, edges, halo order and send lists come out
    exactly as ``decompose(channel(world*nx, ny), slab_part(..))`` gives them (the numberings of the
    window and of the whole channel are order-isomorphic, and with h = 1 the coordinates are the same
    doubles) -- tests/test_decomp.py::test_slab_window_equals_global_decomposition -- at O(nx*ny)
    instead of O(world*nx*ny) time and memory per rank.  Returns (window case, RankMesh)."""
    from . import cases
    pad_l = 1 if rank > 0 else 0
    pad_r = 1 if rank < world - 1 else 0
    x_lo = rank * nx * h
    c = cases.channel(nx + pad_l + pad_r, ny, h=h, x0=x_lo - pad_l * h)
    cx = c.mesh.cell_cx
    part = np.where(cx < x_lo, 0, np.where(cx < x_lo + nx * h, 1, 2)).astype(np.int32)
    rm3 = decompose(c.mesh, part, 3, only_rank=1)[1]
    recv = np.zeros(world, dtype=np.int32)
    send = [np.empty(0, np.int32) for _ in range(world)]
    if pad_l:
        recv[rank - 1] = rm3.recv_count[0]; send[rank - 1] = rm3.send_ind[0]
    if pad_r:
        recv[rank + 1] = rm3.recv_count[2]; send[rank + 1] = rm3.send_ind[2]
    rm = dataclasses.replace(rm3, rank=rank, nranks=world, recv_count=recv, send_ind=send)
    return c, rm


def parse_proc_file(path: str) -> dict:
    """Integer maps of a reference ``mesh.NNNN.proc`` file (format: decomp.cpp:295-334)."""
    with open(path) as f:
        tok = f.read().split("\n")
    it = iter(tok)

    def nonempty():
        for ln in it:
            if ln.strip():
                return ln
        raise StopIteration

    n, n_ex = map(int, nonempty().split())
    for _ in range(n_ex):
        next(it)
    c, c_ex = map(int, nonempty().split())
    cells = np.array([list(map(int, next(it).split()[:4])) for _ in range(c_ex)], dtype=np.int32).reshape(-1, 4)
    e, e_ex = map(int, nonempty().split())
    edges = np.array([list(map(int, next(it).split()[:4])) for _ in range(e_ex)], dtype=np.int32).reshape(-1, 4)
    recv = np.array(list(map(int, nonempty().split())), dtype=np.int32)
    send = []
    for _ in range(recv.shape[0]):
        p, cnt = map(int, nonempty().split())
        send.append(np.array(list(map(int, nonempty().split())), dtype=np.int32) if cnt else np.empty(0, np.int32))
    return dict(nCount=n, nCountEx=n_ex, cCount=c, cCountEx=c_ex, eCount=e, eCountEx=e_ex, cells=cells, edges=edges,
                recv_count=recv, send_ind=send)


# ------------------------------------------------------------------------------------------------
# per-rank solver for torchrun launches (bench.py, tests): every rank builds the same global mesh
# deterministically, partitions it, and keeps its own share.
# ------------------------------------------------------------------------------------------------
def nccl_unique_id(dist, rank: int) -> bytes:
    """ncclUniqueId created on rank 0 by the library and broadcast through torch.distributed."""
    import torch
    from . import fvm
    lib = fvm.load_library()
    buf = (C.c_char * 128)()
    if rank == 0:
        lib.cfd2d_nccl_get_unique_id.argtypes = [C.c_void_p]
        rc = lib.cfd2d_nccl_get_unique_id(C.cast(buf, C.c_void_p))
        if rc != 0:
            raise fvm.CFDError(rc, "cfd2d_nccl_get_unique_id")
    t = torch.frombuffer(bytearray(bytes(buf)), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


def make_rank_solver(nx, ny, rank, world, device, flux, order, dist, partition="slab", case=None, state=None, tiles=None):
    """Weak-scaling layout: a (world*nx) x ny x 2 channel, partitioned `world` ways."""
    from . import cases, fvm
    if case is None and partition == "slab":
        # weak-scaling bench: build only this rank's window of the channel
        c, rm = slab_rank_mesh(nx, ny, rank, world)
        st = c.smooth_state(tiles=tiles or world, extent=(0.0, float(world * nx), 0.0, float(ny)))
        uid = nccl_unique_id(dist, rank)
        s = fvm.Solver(rm.local, c.task, flux, order, device=device, nc_owned=rm.nc, halo=rm.halo_dict(uid))
        own = rm.g_cells[:rm.nc]
        s.rank_mesh = rm
        return s, tuple(x[own] for x in st), rm.nc, 2 * nx * ny * world
    c = case if case is not None else cases.channel(nx * world, ny)
    st = state if state is not None else c.smooth_state(tiles=(tiles or world) if case is None else 1)
    part = slab_part(c.mesh, world) if partition == "slab" else metis_part(c.mesh, world)
    rm = decompose(c.mesh, part, world, only_rank=rank)[rank]
    uid = nccl_unique_id(dist, rank)
    s = fvm.Solver(rm.local, c.task, flux, order, device=device, nc_owned=rm.nc, halo=rm.halo_dict(uid))
    own = rm.g_cells[:rm.nc]
    s.rank_mesh = rm
    return s, tuple(x[own] for x in st), rm.nc, c.mesh.nc
