"""Mesh construction for the FVM_TVD path: node/triangle lists -> the flat SoA the C-ABI consumes.

This is the host-side restatement of what the reference's Salome-UNV reader builds into its AoS
``Grid`` (``src/mesh/MeshReaderSalomeUnv.cpp:25-264``, ``src/mesh/grid.h:18-100``) -- same edge
numbering, same edge orientation, same floating-point formulas, evaluated in the same order --
vectorised with numpy so 4 M..32 M-cell meshes are built in seconds instead of going through a
multi-GB text file.  ``tests/test_mesh_vs_reference.py`` pins it bit-for-bit against the compiled
reference reader on UNV files written by :func:`write_unv`.

Rules reproduced (file:line in the reference):
  * neighbour k of a cell = the other cell sharing nodes (k, k+1)            (:93-108)
  * an edge is created by the lower-numbered cell, ids ascend in (cell, k)   (:112-149)
  * n1,n2 = nodes (k, k+1) of the creating cell c1                           (:145-146)
  * Gauss points  mid -/+ (1/sqrt 3) (n2-n1)/2                               (:153-163)
  * normal (y2-y1, x1-x2)/l, flipped to point out of c1                      (:164-187)
  * cells[].edgesInd filled in edge-creation (= ascending id) order          (:170-171,193-194)
  * area by Heron from the three edge lengths in edgesInd order              (:245-251)
  * centre = (x0+x1+x2)/3.0                                                  (:86-87)
"""
from __future__ import annotations

import dataclasses
import numpy as np


@dataclasses.dataclass
class Mesh:
    """Flat mesh; field names follow the reference's Grid/Cell/Edge members."""
    nodes: np.ndarray        # (nn, 2) float64
    cell_nodes: np.ndarray   # (nc, 3) int32   Cell::nodesInd
    cell_edges: np.ndarray   # (nc, 3) int32   Cell::edgesInd (ascending edge id)
    cell_neigh: np.ndarray   # (nc, 3) int32   Cell::neigh   (-2 = boundary)
    cell_S: np.ndarray       # (nc,)  float64  Cell::S
    cell_cx: np.ndarray      # (nc,)  float64  Cell::c.x
    cell_cy: np.ndarray      # (nc,)  float64  Cell::c.y
    edge_n1: np.ndarray      # (ne,)  int32
    edge_n2: np.ndarray      # (ne,)  int32
    edge_c1: np.ndarray      # (ne,)  int32
    edge_c2: np.ndarray      # (ne,)  int32    -1 on a boundary edge
    edge_nx: np.ndarray      # (ne,)  float64  Edge::n.x (out of c1)
    edge_ny: np.ndarray      # (ne,)  float64
    edge_l: np.ndarray       # (ne,)  float64
    edge_gp: np.ndarray      # (ne, 4) float64 Edge::c[1].x, c[1].y, c[2].x, c[2].y
    # bound later from the task description:
    cell_mat: np.ndarray | None = None   # (nc,) int32 material index of the cell's region
    edge_bc: np.ndarray | None = None    # (ne,) int32 index into the boundary table, -1 inner
    cell_region: np.ndarray | None = None  # (nc,) int32 region index (initial state)

    @property
    def nc(self) -> int:
        return int(self.cell_nodes.shape[0])

    @property
    def ne(self) -> int:
        return int(self.edge_c1.shape[0])

    @property
    def nn(self) -> int:
        return int(self.nodes.shape[0])

    def boundary_edges(self) -> np.ndarray:
        return np.nonzero(self.edge_c2 < 0)[0].astype(np.int32)

    def edge_key_lookup(self, pairs: np.ndarray) -> np.ndarray:
        """Edge ids of the given (m, 2) node pairs (either orientation); -1 if absent."""
        nn = np.int64(self.nn)
        a = np.minimum(self.edge_n1, self.edge_n2).astype(np.int64)
        b = np.maximum(self.edge_n1, self.edge_n2).astype(np.int64)
        keys = a * nn + b
        order = np.argsort(keys, kind="stable")
        skeys = keys[order]
        pa = np.minimum(pairs[:, 0], pairs[:, 1]).astype(np.int64)
        pb = np.maximum(pairs[:, 0], pairs[:, 1]).astype(np.int64)
        q = pa * nn + pb
        pos = np.searchsorted(skeys, q)
        pos = np.clip(pos, 0, len(skeys) - 1)
        hit = skeys[pos] == q
        out = np.where(hit, order[pos], -1)
        return out.astype(np.int32)


def build_mesh(nodes: np.ndarray, tris: np.ndarray) -> Mesh:
    """Nodes (nn,2) + triangles (nc,3) -> Mesh, reproducing MeshReaderSalomeUnv::read."""
    nodes = np.ascontiguousarray(nodes, dtype=np.float64)
    tris = np.ascontiguousarray(tris, dtype=np.int32)
    nc = tris.shape[0]
    nn = nodes.shape[0]
    x = nodes[:, 0]
    y = nodes[:, 1]

    # ---- neighbours through half-edge keys (reference: set_intersection of node->cells maps)
    ha = tris.astype(np.int64)                     # node k
    hb = np.roll(tris, -1, axis=1).astype(np.int64)  # node k+1
    key = (np.minimum(ha, hb) * np.int64(nn) + np.maximum(ha, hb)).ravel()   # (3nc,) order (cell,k)
    order = np.argsort(key, kind="stable")
    sk = key[order]
    same_next = np.zeros(3 * nc, dtype=bool)
    same_next[:-1] = sk[1:] == sk[:-1]
    same_prev = np.zeros(3 * nc, dtype=bool)
    same_prev[1:] = same_next[:-1]
    partner = np.full(3 * nc, -1, dtype=np.int64)   # half-edge index of the twin
    idx = np.nonzero(same_next)[0]
    partner[order[idx]] = order[idx + 1]
    partner[order[idx + 1]] = order[idx]
    neigh = np.where(partner >= 0, partner // 3, -2).astype(np.int32).reshape(nc, 3)
    del same_prev

    # ---- edge creation: half-edge (i,k) creates an edge iff boundary or neighbour > i
    cell_of_he = np.repeat(np.arange(nc, dtype=np.int32), 3)
    nb = neigh.ravel()
    creates = (nb == -2) | (nb > cell_of_he)
    he_ids = np.nonzero(creates)[0]                 # ascending (cell,k) == ascending edge id
    ne = he_ids.shape[0]
    edge_of_he = np.full(3 * nc, -1, dtype=np.int64)
    edge_of_he[he_ids] = np.arange(ne, dtype=np.int64)
    twin = partner[he_ids]
    has_twin = twin >= 0
    edge_of_he[twin[has_twin]] = np.nonzero(has_twin)[0]

    c1 = cell_of_he[he_ids].astype(np.int32)
    c2 = np.where(has_twin, nb[he_ids], -1).astype(np.int32)
    n1 = ha.ravel()[he_ids].astype(np.int32)
    n2 = hb.ravel()[he_ids].astype(np.int32)

    # ---- cell centres
    cx = (x[tris[:, 0]] + x[tris[:, 1]] + x[tris[:, 2]]) / 3.0
    cy = (y[tris[:, 0]] + y[tris[:, 1]] + y[tris[:, 2]]) / 3.0

    # ---- edge geometry (formulas evaluated in the reference's order)
    x1, y1, x2, y2 = x[n1], y[n1], x[n2], y[n2]
    s3 = 1.0 / np.sqrt(3.0)
    mx = (x1 + x2) / 2.0
    my = (y1 + y2) / 2.0
    gp = np.empty((ne, 4), dtype=np.float64)
    gp[:, 0] = mx - s3 * (x2 - x1) / 2.0
    gp[:, 1] = my - s3 * (y2 - y1) / 2.0
    gp[:, 2] = mx + s3 * (x2 - x1) / 2.0
    gp[:, 3] = my + s3 * (y2 - y1) / 2.0
    nx_ = y2 - y1
    ny_ = x1 - x2
    l = np.sqrt(nx_ * nx_ + ny_ * ny_)
    nx_ = nx_ / l
    ny_ = ny_ / l
    vcx = cx[c1] - mx
    vcy = cy[c1] - my
    flip = (vcx * nx_ + vcy * ny_) > 0
    nx_ = np.where(flip, nx_ * -1, nx_)
    ny_ = np.where(flip, ny_ * -1, ny_)

    # ---- cells[].edgesInd: the cell's three edges in ascending id
    cell_edges = np.sort(edge_of_he.reshape(nc, 3), axis=1).astype(np.int32)

    # ---- Heron area from edge lengths in edgesInd order
    a = l[cell_edges[:, 0]]
    b = l[cell_edges[:, 1]]
    c = l[cell_edges[:, 2]]
    p = (a + b + c) / 2.0
    S = np.sqrt(p * (p - a) * (p - b) * (p - c))

    return Mesh(nodes=nodes, cell_nodes=tris, cell_edges=cell_edges, cell_neigh=neigh,
                cell_S=S, cell_cx=cx, cell_cy=cy, edge_n1=n1, edge_n2=n2, edge_c1=c1, edge_c2=c2,
                edge_nx=nx_, edge_ny=ny_, edge_l=l, edge_gp=gp)


# --------------------------------------------------------------------------------------------
# synthetic generators (structured-triangulated rectangle)
# --------------------------------------------------------------------------------------------

def rect_tri_nodes(nx: int, ny: int, lx: float, ly: float, jitter: float = 0.0, seed: int = 1234,
                   x0: float = 0.0, y0: float = 0.0):
    """(nx x ny) quads on [x0,x0+lx]x[y0,y0+ly], each split on the same diagonal into two CCW
    triangles -> N = 2 nx ny cells.  Interior nodes may be jittered by <= jitter*h (seeded).
    Returns nodes (nn,2), tris (nc,3), and a dict side-name -> (m,2) boundary node pairs."""
    xs = x0 + lx * np.arange(nx + 1, dtype=np.float64) / nx
    ys = y0 + ly * np.arange(ny + 1, dtype=np.float64) / ny
    X, Y = np.meshgrid(xs, ys)          # (ny+1, nx+1), node id = j*(nx+1)+i
    if jitter > 0.0:
        rng = np.random.default_rng(seed)
        hx, hy = lx / nx, ly / ny
        dx = (rng.random(X.shape) - 0.5) * 2.0 * jitter * hx
        dy = (rng.random(Y.shape) - 0.5) * 2.0 * jitter * hy
        dx[0, :] = dx[-1, :] = 0.0
        dy[0, :] = dy[-1, :] = 0.0
        dx[:, 0] = dx[:, -1] = 0.0
        dy[:, 0] = dy[:, -1] = 0.0
        X = X + dx
        Y = Y + dy
    nodes = np.stack([X.ravel(), Y.ravel()], axis=1)
    i = np.arange(nx, dtype=np.int64)
    j = np.arange(ny, dtype=np.int64)
    I, J = np.meshgrid(i, j)            # (ny, nx)
    n00 = (J * (nx + 1) + I).ravel()
    n10 = n00 + 1
    n01 = n00 + (nx + 1)
    n11 = n01 + 1
    tris = np.empty((2 * nx * ny, 3), dtype=np.int32)
    tris[0::2, 0] = n00; tris[0::2, 1] = n10; tris[0::2, 2] = n11
    tris[1::2, 0] = n00; tris[1::2, 1] = n11; tris[1::2, 2] = n01
    bi = np.arange(nx, dtype=np.int32)
    bj = np.arange(ny, dtype=np.int32)
    sides = {
        "bottom": np.stack([bi, bi + 1], axis=1),
        "top": np.stack([ny * (nx + 1) + bi, ny * (nx + 1) + bi + 1], axis=1),
        "left": np.stack([bj * (nx + 1), (bj + 1) * (nx + 1)], axis=1),
        "right": np.stack([bj * (nx + 1) + nx, (bj + 1) * (nx + 1) + nx], axis=1),
    }
    return nodes, tris, sides


def step_channel(nx: int, ny: int, lx: float, ly: float, step_x: float, step_h: float,
                 jitter: float = 0.0, seed: int = 1234):
    """Forward-facing-step channel (BASELINE config 2 outline): the rectangle with the block
    [step_x, lx] x [0, step_h] cut out.  Boundary sides: left, right, top, bottom, step."""
    nodes, tris, _ = rect_tri_nodes(nx, ny, lx, ly, jitter, seed)
    cx = (nodes[tris[:, 0], 0] + nodes[tris[:, 1], 0] + nodes[tris[:, 2], 0]) / 3.0
    cy = (nodes[tris[:, 0], 1] + nodes[tris[:, 1], 1] + nodes[tris[:, 2], 1]) / 3.0
    keep = ~((cx > step_x) & (cy < step_h))
    tris = tris[keep]
    used = np.zeros(nodes.shape[0], dtype=bool)
    used[tris.ravel()] = True
    remap = np.cumsum(used) - 1
    nodes = nodes[used]
    tris = remap[tris].astype(np.int32)
    m = build_mesh(nodes, tris)
    be = m.boundary_edges()
    mx = 0.5 * (nodes[m.edge_n1[be], 0] + nodes[m.edge_n2[be], 0])
    my = 0.5 * (nodes[m.edge_n1[be], 1] + nodes[m.edge_n2[be], 1])
    tol = 1e-9 * max(lx, ly)
    pairs = np.stack([m.edge_n1[be], m.edge_n2[be]], axis=1)
    sides = {}
    left = mx < tol
    right = mx > lx - tol
    top = my > ly - tol
    bottom = (my < tol) & ~left & ~right
    step = ~(left | right | top | bottom)
    for name, mask in (("left", left), ("right", right), ("top", top), ("bottom", bottom), ("step", step)):
        sides[name] = pairs[mask]
    return nodes, tris, sides


def shuffle_cells(tris: np.ndarray, seed: int = 4321) -> np.ndarray:
    """Seeded random permutation of the cell order (exercises genuinely unstructured numbering)."""
    rng = np.random.default_rng(seed)
    return tris[rng.permutation(tris.shape[0])]


# --------------------------------------------------------------------------------------------
# Salome UNV writer (the subset MeshReaderSalomeUnv.cpp:267-448 parses)
# --------------------------------------------------------------------------------------------

def write_unv(path: str, nodes: np.ndarray, tris: np.ndarray,
              cell_groups: dict[str, np.ndarray], edge_groups: dict[str, np.ndarray]) -> None:
    """Write nodes/triangles/boundary edges/groups as UNV blocks 2411, 2412, 2467.

    Labels: boundary-edge elements first (1..B), then triangles (B+1..B+nc), in the given order.
    The file deliberately ends on the closing ``-1`` WITHOUT a newline: the reference's read loop
    (``:31-34``) would otherwise parse an empty block and dereference ``sl[0]``.
    """
    out = []
    out.append("    -1\n  2411\n")
    for i, (px, py) in enumerate(nodes):
        out.append(f"{i + 1:10d}{1:10d}{1:10d}{11:10d}\n")
        out.append(f"  {float(px)!r}  {float(py)!r}  0.0\n")
    out.append("    -1\n    -1\n  2412\n")
    label = 0
    edge_label = {}
    for name, pairs in edge_groups.items():
        labs = []
        for a, b in np.asarray(pairs):
            label += 1
            out.append(f"{label:10d}{11:10d}{2:10d}{1:10d}{7:10d}{2:10d}\n")
            out.append(f"{0:10d}{1:10d}{1:10d}\n")
            out.append(f"{int(a) + 1:10d}{int(b) + 1:10d}\n")
            labs.append(label)
        edge_label[name] = labs
    first_cell_label = label + 1
    for t in tris:
        label += 1
        out.append(f"{label:10d}{41:10d}{2:10d}{1:10d}{7:10d}{3:10d}\n")
        out.append(f"{int(t[0]) + 1:10d}{int(t[1]) + 1:10d}{int(t[2]) + 1:10d}\n")
    out.append("    -1\n    -1\n  2467\n")
    gid = 0

    def group(name, labels):
        nonlocal gid
        gid += 1
        n = len(labels)
        out.append(f"{gid:10d}{0:10d}{0:10d}{0:10d}{0:10d}{0:10d}{0:10d}{n:10d}\n")
        out.append(f"{name}\n")
        for k in range(0, n - 1, 2):
            out.append(f"{8:10d}{labels[k]:10d}{0:10d}{0:10d}{8:10d}{labels[k + 1]:10d}{0:10d}{0:10d}\n")
        if n % 2 == 1:
            out.append(f"{8:10d}{labels[-1]:10d}{0:10d}{0:10d}\n")

    for name, cells in cell_groups.items():
        group(name, [first_cell_label + int(c) for c in np.asarray(cells)])
    for name, labs in edge_label.items():
        group(name, labs)
    out.append("    -1")
    with open(path, "w") as f:
        f.write("".join(out))
