#!/usr/bin/env python
"""Sweep the device edge order of k_flux (CFD2D_EDGE_TILE x CFD2D_EDGE_SORT) at the bench workload."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cfd2d_b200 import cases, fvm
c = cases.channel(2000, 1000)
st = c.smooth_state()
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
for flux, order in ((0, 2), (1, 2)):
    for sort in ("1", "2"):
        for tile in ("16384", "65536"):
            os.environ["CFD2D_EDGE_SORT"] = sort; os.environ["CFD2D_EDGE_TILE"] = tile
            s = fvm.Solver(c.mesh, c.task, flux, order)
            s.set_stream(stream.cuda_stream); s.set_state(*st); s.calc_time_step(); s.step(3)
            p = s.profile(5)
            print(json.dumps({"flux": flux, "order": order, "edge_sort": int(sort), "edge_tile": int(tile),
                              "flux_ms": p["flux"][0] / p["flux"][1], "grad_ms": p["grad"][0] / max(1, p["grad"][1]),
                              "update_ms": (p["update1"][0] + p["update2"][0]) / 10}), flush=True)
            s.close()
