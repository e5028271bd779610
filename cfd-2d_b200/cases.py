"""Synthetic cases for parity and throughput (BASELINE.json configs).

Every case is a structured-triangulated outline (optionally jittered / cell-shuffled so the
numbering is genuinely unstructured) plus a task description.  Cell size is h = 1 m on purpose:
the reference's time step is ``CFL * S / (|u|+c)`` with S the cell AREA (fvm_tvd.cpp:223), so the
effective Courant number is ``CFL * h / 2`` -- with h = 1 m and CFL = 0.3 waves cross ~0.15 cell
per step and 100 steps give a visibly evolved, stable flow.
"""
from __future__ import annotations

import dataclasses

import numpy as np

from . import mesh as _mesh
from . import task as _task


@dataclasses.dataclass
class Case:
    name: str
    nodes: np.ndarray
    tris: np.ndarray
    cell_groups: dict      # region name -> cell ids
    edge_groups: dict      # boundary name -> (m,2) node pairs
    task: _task.Task
    mesh: _mesh.Mesh

    def initial_state(self):
        """Piecewise-constant region state, as FVM_TVD::init forms it (fvm_tvd.cpp:199-204)."""
        nc = self.mesh.nc
        ro = np.empty(nc); ru = np.empty(nc); rv = np.empty(nc); re = np.empty(nc)
        for ir, r in enumerate(self.task.regions):
            s = _task.region_state(self.task, r)
            sel = self.mesh.cell_region == ir
            ro[sel], ru[sel], rv[sel], re[sel] = s
        return ro, ru, rv, re

    def smooth_state(self, amp=0.05, u0=50.0, v0=0.0, T0=300.0, p0=1.0e5, sigma_frac=0.08, tiles=1, extent=None):
        """Gaussian pressure bump on a uniform stream (the reference terminates on it).
        Returned as conservative variables with the reference's own conversions.
        ``tiles`` > 1 (weak-scaling runs): the domain is ``tiles`` slabs side by side in x and every
        slab carries the bump of the one-slab case, so each rank solves statistically the same
        problem as the single-GPU run (the exact Riemann solver's cost depends on the data)."""
        m = self.task.materials[0]
        cx, cy = self.mesh.cell_cx, self.mesh.cell_cy
        if extent is None:
            extent = (self.nodes[:, 0].min(), self.nodes[:, 0].max(), self.nodes[:, 1].min(), self.nodes[:, 1].max())
        x0 = extent[0]                     # extent of the WHOLE domain (this case may be a window of it)
        lx = (extent[1] - x0) / tiles
        ly = extent[3] - extent[2]
        yc = extent[2] + 0.5 * ly
        sig = sigma_frac * max(lx, ly)
        bump = np.zeros_like(cx)
        for k in range(tiles):
            xc = x0 + (k + 0.4) * lx
            bump += np.exp(-((cx - xc) ** 2 + (cy - yc) ** 2) / (sig * sig))
        p = p0 * (1.0 + amp * bump)
        T = np.full_like(p, T0)
        u = u0 * (1.0 + 0.1 * amp * bump)
        v = v0 + 5.0 * amp * bump
        Cv = m.Cp - _task.GR / m.M
        gam = m.Cp / Cv
        r = p * m.M / (T * _task.GR)
        e = p / (r * (gam - 1))
        return r, r * u, r * v, r * (e + 0.5 * (u * u + v * v))

    def write(self, workdir: str, xml: str = "task.xml"):
        import os
        os.makedirs(workdir, exist_ok=True)
        _mesh.write_unv(os.path.join(workdir, self.task.mesh_name), self.nodes, self.tris,
                        self.cell_groups, self.edge_groups)
        _task.write_task_xml(os.path.join(workdir, xml), self.task)


def bind(m: _mesh.Mesh, t: _task.Task, cell_groups: dict, edge_groups: dict) -> None:
    """Region -> material and boundary-name -> BC binding (fvm_tvd.cpp:132-174, :199-204)."""
    m.cell_region = np.full(m.nc, -1, dtype=np.int32)
    for ir, r in enumerate(t.regions):
        if r.name in cell_groups:
            m.cell_region[np.asarray(cell_groups[r.name])] = ir
    if (m.cell_region < 0).any():
        raise ValueError("ERROR: unknown cell name: a cell belongs to no region")
    mat_of_region = np.array([r.mat_id for r in t.regions], dtype=np.int32)
    m.cell_mat = mat_of_region[m.cell_region]
    m.edge_bc = np.full(m.ne, -1, dtype=np.int32)
    for ib, b in enumerate(t.boundaries):
        if b.name in edge_groups and len(edge_groups[b.name]):
            ids = m.edge_key_lookup(np.asarray(edge_groups[b.name]))
            if (ids < 0).any():
                raise ValueError(f"boundary group {b.name}: edge not in mesh")
            m.edge_bc[ids] = ib
    be = m.edge_c2 < 0
    if (m.edge_bc[be] < 0).any():
        raise ValueError("ERROR (boundary condition): unknown edge type of a boundary edge")


def _finish(name, nodes, tris, cell_groups, edge_groups, t) -> Case:
    m = _mesh.build_mesh(nodes, tris)
    bind(m, t, cell_groups, edge_groups)
    return Case(name=name, nodes=nodes, tris=tris, cell_groups=cell_groups, edge_groups=edge_groups,
                task=t, mesh=m)


def strip(nx=200, ny=50, jump="weak", jitter=0.0, shuffle=False, h=1.0) -> Case:
    """C1 / BASELINE config 1: strip, regions left|right, all walls.
    jump='weak': P=1e5|9e4, T=348.4|340 (runs through the reference's 2nd-order Godunov path);
    jump='sod' : 10:1 (only for the first-order LF variant V1)."""
    nodes, tris, sides = _mesh.rect_tri_nodes(nx, ny, nx * h, ny * h, jitter)
    if shuffle:
        tris = _mesh.shuffle_cells(tris)
    cx = (nodes[tris[:, 0], 0] + nodes[tris[:, 1], 0] + nodes[tris[:, 2], 0]) / 3.0
    left = np.nonzero(cx < 0.5 * nx * h)[0]
    right = np.nonzero(cx >= 0.5 * nx * h)[0]
    t = _task.Task(TAU=1.0e10)
    if jump == "weak":
        t.regions = [_task.Region("left", 0, 0.0, 0.0, 348.4, 1.0e5), _task.Region("right", 0, 0.0, 0.0, 340.0, 9.0e4)]
    elif jump == "sod":
        t.regions = [_task.Region("left", 0, 0.0, 0.0, 348.4, 1.0e5), _task.Region("right", 0, 0.0, 0.0, 278.7, 1.0e4)]
    else:
        raise ValueError(jump)
    t.boundaries = [_task.BoundCond("walls", _task.BOUND_WALL_SLIP)]
    walls = np.concatenate([sides["bottom"], sides["right"], sides["top"], sides["left"]])
    return _finish(f"strip{nx}x{ny}", nodes, tris, {"left": left, "right": right}, {"walls": walls}, t)


def channel(nx=200, ny=100, jitter=0.0, shuffle=False, h=1.0, two_materials=False, x0=0.0) -> Case:
    """C2/C3/C5-style channel: inlet left, outlet right, slip wall bottom, no-slip wall top.
    One region (uniform stream); smooth initial data are injected with ``Case.smooth_state``.
    ``x0``: x of the left side (a window of a longer channel, see decomp.slab_rank_mesh)."""
    nodes, tris, sides = _mesh.rect_tri_nodes(nx, ny, nx * h, ny * h, jitter, x0=x0)
    if shuffle:
        tris = _mesh.shuffle_cells(tris)
    t = _task.Task(TAU=1.0e10)
    cells = np.arange(tris.shape[0])
    if two_materials:
        t.materials = [_task.Material("air", 0.02898, 1004.5), _task.Material("gas2", 0.0400, 900.0)]
        cy = (nodes[tris[:, 0], 1] + nodes[tris[:, 1], 1] + nodes[tris[:, 2], 1]) / 3.0
        lo = np.nonzero(cy < 0.5 * ny * h)[0]
        hi = np.nonzero(cy >= 0.5 * ny * h)[0]
        t.regions = [_task.Region("lower", 0, 50.0, 0.0, 300.0, 1.0e5), _task.Region("upper", 1, 50.0, 0.0, 300.0, 1.0e5)]
        groups = {"lower": lo, "upper": hi}
    else:
        t.regions = [_task.Region("flow", 0, 50.0, 0.0, 300.0, 1.0e5)]
        groups = {"flow": cells}
    t.boundaries = [
        _task.BoundCond("inlet", _task.BOUND_INLET, 50.0, 0.0, 300.0, 1.0e5),
        _task.BoundCond("outlet", _task.BOUND_OUTLET),
        _task.BoundCond("wall_bottom", _task.BOUND_WALL_SLIP),
        _task.BoundCond("wall_top", _task.BOUND_WALL_NO_SLIP),
    ]
    eg = {"inlet": sides["left"], "outlet": sides["right"], "wall_bottom": sides["bottom"], "wall_top": sides["top"]}
    return _finish(f"channel{nx}x{ny}", nodes, tris, groups, eg, t)


def forward_step(nx=300, ny=100, jitter=0.0, h=1.0) -> Case:
    """BASELINE config 2 outline: channel 3:1 with a step of height 0.2 Ly starting at 0.2 Lx."""
    lx, ly = nx * h, ny * h
    nodes, tris, sides = _mesh.step_channel(nx, ny, lx, ly, 0.2 * lx, 0.2 * ly, jitter)
    t = _task.Task(TAU=1.0e10)
    t.regions = [_task.Region("flow", 0, 50.0, 0.0, 300.0, 1.0e5)]
    t.boundaries = [
        _task.BoundCond("inlet", _task.BOUND_INLET, 50.0, 0.0, 300.0, 1.0e5),
        _task.BoundCond("outlet", _task.BOUND_OUTLET),
        _task.BoundCond("walls", _task.BOUND_WALL_SLIP),
    ]
    walls = np.concatenate([sides["bottom"], sides["top"], sides["step"]])
    eg = {"inlet": sides["left"], "outlet": sides["right"], "walls": walls}
    return _finish(f"step{nx}x{ny}", nodes, tris, {"flow": np.arange(tris.shape[0])}, eg, t)
