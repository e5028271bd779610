"""Reader for the Salome-UNV subset the reference parses (``src/mesh/MeshReaderSalomeUnv.cpp:267-448``,
SURVEY.md Appendix C): blocks 2411 (nodes), 2412 (fe_id 11 boundary edges / 41 triangles) and 2467
(named groups).  Returns exactly what ``mesh.build_mesh`` + ``cases.bind`` need."""
from __future__ import annotations

import numpy as np


def _blocks(path):
    with open(path) as f:
        lines = f.read().split("\n")
    cur, inside = [], False
    for ln in lines:
        pos = ln.find("-1")
        is_delim = pos != -1 and pos == len(ln) - 2      # MeshReaderSalomeUnv.cpp:277-283
        if is_delim:
            if inside:
                yield cur
                cur, inside = [], False
            else:
                inside = True
            continue
        if inside:
            cur.append(ln)
    if inside and cur:
        yield cur


def read_unv(path, native=None):
    """(nodes, tris, cell_groups, edge_groups).  By default the file is parsed by the native reader
    of the C-ABI library (``cfd2d_unv_read``, csrc/unv_reader.cpp: one linear pass, ~20x faster);
    ``native=False`` (or ``CFD2D_UNV_PYTHON=1``) uses the pure-Python parser below, which the tests
    hold the native one equal to."""
    import os
    if native is None:
        native = os.environ.get("CFD2D_UNV_PYTHON", "0") != "1"
    if native:
        return read_unv_native(path)
    return read_unv_python(path)


def read_unv_native(path):
    import ctypes as C
    from . import fvm
    lib = fvm.load_library()
    H = C.c_void_p
    i64, i32 = C.c_int64, C.c_int32
    lib.cfd2d_unv_read.argtypes = [C.c_char_p, C.POINTER(H)]
    lib.cfd2d_unv_counts.argtypes = [H, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i32)]
    lib.cfd2d_unv_copy.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.cfd2d_unv_group_name.argtypes = [H, C.c_int]
    lib.cfd2d_unv_group_name.restype = C.c_char_p
    lib.cfd2d_unv_group_counts.argtypes = [H, C.c_int, C.POINTER(i64), C.POINTER(i64)]
    lib.cfd2d_unv_group_copy.argtypes = [H, C.c_int, C.c_void_p, C.c_void_p]
    lib.cfd2d_unv_free.argtypes = [H]
    lib.cfd2d_unv_free.restype = None
    h = H()
    rc = lib.cfd2d_unv_read(str(path).encode(), C.byref(h))
    if rc != 0:
        msg = lib.cfd2d_fvm_last_error(None)
        raise ValueError(msg.decode() if msg else f"cfd2d_unv_read failed ({rc})")
    try:
        nn, nc, nbe, ng = i64(), i64(), i64(), i32()
        lib.cfd2d_unv_counts(h, C.byref(nn), C.byref(nc), C.byref(nbe), C.byref(ng))
        nodes = np.empty((nn.value, 2), np.float64)
        tris = np.empty((nc.value, 3), np.int32)
        lib.cfd2d_unv_copy(h, nodes.ctypes.data, tris.ctypes.data, None)
        cell_groups, edge_groups = {}, {}
        for g in range(ng.value):
            name = lib.cfd2d_unv_group_name(h, g).decode()
            a, b = i64(), i64()
            lib.cfd2d_unv_group_counts(h, g, C.byref(a), C.byref(b))
            cells = np.empty(a.value, np.int64)
            edges = np.empty((b.value, 2), np.int32)
            lib.cfd2d_unv_group_copy(h, g, cells.ctypes.data if a.value else None, edges.ctypes.data if b.value else None)
            if a.value:
                cell_groups[name] = cells
            if b.value:
                edge_groups[name] = edges
    finally:
        lib.cfd2d_unv_free(h)
    return nodes, tris, cell_groups, edge_groups


def read_unv_python(path):
    nodes, tris, edge_elems = [], [], []
    elem = {}          # label-1 -> ("cell"|"edge", index)
    groups = {}
    for b in _blocks(path):
        if not b:
            continue
        kind = int(b[0].split()[0])
        it = iter(b[1:])
        if kind == 2411:
            for rec in it:
                if not rec.strip():
                    continue
                xyz = next(it).split()
                nodes.append((float(xyz[0]), float(xyz[1])))
        elif kind == 2412:
            for rec in it:
                if not rec.strip():
                    continue
                t = rec.split()
                label, fe = int(t[0]) - 1, int(t[1])
                if fe == 11:
                    next(it)
                    n = next(it).split()
                    elem[label] = ("edge", len(edge_elems))
                    edge_elems.append((int(n[0]) - 1, int(n[1]) - 1))
                elif fe == 41:
                    n = next(it).split()
                    elem[label] = ("cell", len(tris))
                    tris.append((int(n[0]) - 1, int(n[1]) - 1, int(n[2]) - 1))
                else:
                    raise ValueError(f"Unknown element type '{fe}'.")   # Exception::TYPE_MESH_UNV_UNKNOWN_ELEMENT
        elif kind == 2467:
            for rec in it:
                if not rec.strip():
                    continue
                n = int(rec.split()[7])
                name = next(it).split()[0]
                labs = []
                for _ in range(n // 2):
                    t = next(it).split()
                    labs += [int(t[1]) - 1, int(t[5]) - 1]
                if n % 2 == 1:
                    labs.append(int(next(it).split()[1]) - 1)
                groups[name] = labs
    cell_groups, edge_groups = {}, {}
    for name in sorted(groups):                 # std::map iteration order (bnd_map)
        cells = [elem[l][1] for l in groups[name] if l in elem and elem[l][0] == "cell"]
        edges = [edge_elems[elem[l][1]] for l in groups[name] if l in elem and elem[l][0] == "edge"]
        if cells:
            cell_groups[name] = np.array(cells, dtype=np.int64)
        if edges:
            edge_groups[name] = np.array(edges, dtype=np.int32).reshape(-1, 2)
    return (np.array(nodes, dtype=np.float64).reshape(-1, 2), np.array(tris, dtype=np.int32).reshape(-1, 3),
            cell_groups, edge_groups)
