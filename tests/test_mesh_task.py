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


# ------------------------------------------------------------------------------------------------
# native UNV reader (csrc/unv_reader.cpp) == the Python parser, which the tests
# above hold equal to the reference's MeshReaderSalomeUnv
# ------------------------------------------------------------------------------------------------
def _same_unv(a, b):
    assert np.array_equal(a[0], b[0]) and a[0].dtype == b[0].dtype            # node coordinates: same doubles
    assert np.array_equal(a[1], b[1])
    assert list(a[2]) == list(b[2]) and list(a[3]) == list(b[3])             # group names, in std::map order
    for k in a[2]:
        assert np.array_equal(a[2][k], b[2][k])
    for k in a[3]:
        assert np.array_equal(a[3][k], b[3][k])


@pytest.mark.parametrize("make", [
    lambda: cases.strip(12, 5, jitter=0.2, shuffle=True),
    lambda: cases.channel(9, 7, jitter=0.3, two_materials=True),          # odd group sizes -> the 4-int tail line
    lambda: cases.forward_step(30, 10, jitter=0.15),
])
def test_native_unv_reader_equals_python_parser(make, tmp_path):
    from cfd2d_b200 import unv
    c = make()
    c.write(str(tmp_path))
    p = os.path.join(str(tmp_path), c.task.mesh_name)
    _same_unv(unv.read_unv(p, native=True), unv.read_unv(p, native=False))
    # the file as the generator writes it ends WITHOUT a newline; with one, and with CRLF, too
    txt = open(p).read()
    open(p, "w").write(txt + "\n")
    _same_unv(unv.read_unv(p, native=True), unv.read_unv(p, native=False))


def test_native_unv_reader_errors(tmp_path):
    from cfd2d_b200 import unv
    with pytest.raises(ValueError):
        unv.read_unv(os.path.join(str(tmp_path), "missing.unv"), native=True)
    c = cases.strip(4, 2)
    c.write(str(tmp_path))
    p = os.path.join(str(tmp_path), c.task.mesh_name)
    txt = open(p).read().replace("        41         2", "        44         2", 1)   # a quad element
    open(p, "w").write(txt)
    with pytest.raises(ValueError, match="Unknown element type '44'"):
        unv.read_unv(p, native=True)
    with pytest.raises(ValueError, match="Unknown element type '44'"):
        unv.read_unv(p, native=False)


def test_native_unv_reader_survives_hostile_input(tmp_path):
    """no exception may cross the C ABI and no input may drive an unbounded allocation:
    negative 2467 entity count, absurd 2412 element label, a 2411 coordinate line holding one number
    (the reference's sscanf stays on the line; strtod would eat the next record's label)."""
    from cfd2d_b200 import unv
    D = "    -1\n"
    cases_ = {
        "neg_count": D + "  2467\n" + "         1         0         0         0         0         0         0        -5\nGRP\n" + D,
        "huge_label": D + "  2412\n" + " 2000000000        41         2         1         7         3\n         1         2         3\n" + D,
        "one_coord": D + "  2411\n" + "         1         1         1        11\n   1.0\n         2         1         1        11\n   2.0   3.0   0.0\n" + D,
    }
    for name, txt in cases_.items():
        f = tmp_path / (name + ".unv")
        f.write_text(txt)
        with pytest.raises(ValueError):
            unv.read_unv_native(str(f))
