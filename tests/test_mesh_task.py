"""CPU: host logic -- mesh builder vs the reference UNV reader (golden + live when oracle/_ref is
present), UNV writer/reader round trip, task.xml round trip, case binding."""
import os
import tempfile

import numpy as np
import pytest

import parity_cases as pc
from cfd2d_b200 import cases, mesh as M, task as T

G = os.path.join(os.path.dirname(__file__), "golden")
MESH_KEYS = ["nodes", "cell_nodes", "cell_edges", "cell_neigh", "cell_S", "cell_cx", "cell_cy", "cell_mat",
             "edge_n1", "edge_n2", "edge_c1", "edge_c2", "edge_nx", "edge_ny", "edge_l", "edge_gp", "edge_bc"]


def test_build_mesh_matches_reference_reader_golden():
    c, _, _ = pc.build("strip_weak_v0")
    g = np.load(os.path.join(G, "strip_weak_v0.npz"))
    for k in MESH_KEYS:
        assert np.array_equal(getattr(c.mesh, k), g["mesh_" + k]), k


def test_mesh_invariants():
    c = cases.channel(20, 10, jitter=0.2, shuffle=True)
    m = c.mesh
    assert m.nc == 400 and m.ne == 3 * 200 + 20 + 10
    assert (np.diff(m.cell_edges, axis=1) > 0).all()                 # ascending edge ids = summation order
    assert np.isclose(m.cell_S.sum(), 20.0 * 10.0)
    # closed cells: sum of outward n*l vanishes
    sgn = np.where(m.edge_c1[m.cell_edges] == np.arange(m.nc)[:, None], 1.0, -1.0)
    sx = (sgn * m.edge_nx[m.cell_edges] * m.edge_l[m.cell_edges]).sum(axis=1)
    sy = (sgn * m.edge_ny[m.cell_edges] * m.edge_l[m.cell_edges]).sum(axis=1)
    assert np.abs(sx).max() < 1e-12 and np.abs(sy).max() < 1e-12
    assert ((m.edge_c2 < 0) == (m.edge_bc >= 0)).all()
    assert (m.edge_c1[m.edge_c2 >= 0] < m.edge_c2[m.edge_c2 >= 0]).all()  # created by the lower cell


def test_unbound_boundary_edge_raises():
    nodes, tris, sides = M.rect_tri_nodes(4, 3, 4.0, 3.0)
    t = T.Task(regions=[T.Region("flow")], boundaries=[T.BoundCond("walls")])
    m = M.build_mesh(nodes, tris)
    with pytest.raises(ValueError):
        cases.bind(m, t, {"flow": np.arange(m.nc)}, {"walls": sides["left"]})


def test_task_xml_round_trip():
    c = cases.channel(4, 2, two_materials=True)
    d = tempfile.mkdtemp()
    p = os.path.join(d, "task.xml")
    T.write_task_xml(p, c.task)
    t2 = T.read_task_xml(p)
    assert t2 == c.task


def test_unknown_boundary_type_raises():
    with pytest.raises(ValueError):
        T.BoundCond("x", "BOUND_MAGIC").kind


def test_unv_round_trip():
    from cfd2d_b200 import unv
    c = cases.forward_step(12, 6, jitter=0.1)
    d = tempfile.mkdtemp()
    c.write(d)
    nodes, tris, cg, eg = unv.read_unv(os.path.join(d, c.task.mesh_name))
    assert np.array_equal(nodes, c.nodes) and np.array_equal(tris, c.tris)
    assert set(cg) == set(c.cell_groups) and set(eg) == set(c.edge_groups)
    for k in cg:
        assert np.array_equal(np.sort(cg[k]), np.sort(c.cell_groups[k]))
    m = M.build_mesh(nodes, tris)
    cases.bind(m, c.task, cg, eg)
    assert np.array_equal(m.edge_bc, c.mesh.edge_bc) and np.array_equal(m.cell_mat, c.mesh.cell_mat)


@pytest.mark.skipif(not os.path.exists("/root/reference/src/methods/fvm_tvd.cpp"), reason="reference tree not present")
def test_build_mesh_matches_live_reference_reader():
    from oracle import refharness as R
    if not R.available("v0"):
        pytest.skip("oracle/_ref not built")
    c = cases.forward_step(18, 8, jitter=0.2)
    d = tempfile.mkdtemp()
    c.write(d)
    s = R.RefSolver(d)
    rm = s.mesh()
    for k in MESH_KEYS:
        assert np.array_equal(getattr(c.mesh, k), rm[k]), k
    st = s.state()
    for a, b in zip(c.initial_state(), st[:4]):
        assert np.array_equal(a, b)
