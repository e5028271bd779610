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


def read_unv(path):
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
