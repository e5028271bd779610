"""Legacy-ASCII VTK writer, byte-compatible with ``FVM_TVD::save`` (reference
``src/methods/fvm_tvd.cpp:501-600``), including its quirks: ``i+1 % 8 == 0`` parses as
``i + (1 % 8) == 0`` so most arrays are written on ONE line; "MachNumber" is (u^2+v^2)/c; the field is
spelled "Velosity"; Total_pressure alone uses ``%f`` and a real 8-per-line wrap."""
from __future__ import annotations

import numpy as np


def _row(vals, fmt="%25.16f "):
    return "".join(fmt % v for v in vals) + "\n"


def write_vtk(path, nodes, tris, prim, ctau, gam):
    nc = tris.shape[0]
    r, p, T, u, v, cz = (prim[k] for k in ("r", "p", "T", "u", "v", "cz"))
    out = ["# vtk DataFile Version 2.0\n", "GASDIN data file\n", "ASCII\n", "DATASET UNSTRUCTURED_GRID\n",
           "POINTS %d float\n" % nodes.shape[0]]
    out.append("".join("%f %f %f  " % (x, y, 0.0) for x, y in nodes) + "\n")
    out.append("CELLS %d %d\n" % (nc, 4 * nc))
    out.append("".join("3 %d %d %d\n" % (a, b, c) for a, b, c in tris))
    out.append("\nCELL_TYPES %d\n" % nc)
    out.append("5\n" * nc)
    out.append("\nCELL_DATA %d\nSCALARS Density float 1\nLOOKUP_TABLE default\n" % nc)
    out.append(_row(r))
    out.append("SCALARS Pressure float 1\nLOOKUP_TABLE default\n")
    out.append(_row(p))
    out.append("SCALARS Temperature float 1\nLOOKUP_TABLE default\n")
    out.append(_row(T))
    out.append("SCALARS MachNumber float 1\nLOOKUP_TABLE default\n")
    out.append(_row((u * u + v * v) / cz))
    out.append("VECTORS Velosity float\n")
    out.append("".join("%25.16f %25.16f %25.16f " % (a, b, 0.0) for a, b in zip(u, v)) + "\n")
    out.append("SCALARS Total_pressure float 1\nLOOKUP_TABLE default\n")
    agam = gam - 1.0
    M2 = (u * u + v * v) / (gam * p / r)
    tp = p * np.power(1.0 + 0.5 * M2 * agam, gam / (gam - 1.0))
    s = []
    for i, x in enumerate(tp):
        s.append("%f " % x)
        if (i + 1) % 8 == 0 or i + 1 == nc:
            s.append("\n")
    out.append("".join(s))
    out.append("SCALARS TAU float 1\nLOOKUP_TABLE default\n")
    out.append(_row(ctau))
    with open(path, "w") as f:
        f.write("".join(out))
