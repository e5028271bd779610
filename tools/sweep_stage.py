#!/usr/bin/env python
"""Sweep the tile size (CFD2D_TILE) and block size (CFD2D_NT) of the fused stage kernel at the bench
workload; prints one JSON line per configuration (ms per RK2 step, device-resident)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=2000)
    ap.add_argument("--ny", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--tiles", default="256,384,512,768")
    ap.add_argument("--nts", default="128,256,384,512")
    ap.add_argument("--variants", default="0:2,1:2")
    a = ap.parse_args()
    import torch
    from cfd2d_b200 import cases, fvm
    c = cases.channel(a.nx, a.ny)
    st = c.smooth_state()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for v in a.variants.split(","):
        flux, order = (int(x) for x in v.split(":"))
        for tc in a.tiles.split(","):
            for nt in a.nts.split(","):
                os.environ["CFD2D_FUSED"] = "1"
                os.environ["CFD2D_TILE"] = tc
                os.environ["CFD2D_NT"] = nt
                try:
                    s = fvm.Solver(c.mesh, c.task, flux, order)
                except fvm.CFDError as ex:
                    print(json.dumps({"flux": flux, "order": order, "tile": int(tc), "nt": int(nt), "error": str(ex)}), flush=True)
                    continue
                s.set_stream(stream.cuda_stream)
                s.set_state(*st)
                s.calc_time_step()
                s.step(3)
                torch.cuda.synchronize()
                e0.record(stream); s.step_async(a.steps); e1.record(stream); s.sync()
                ms = e0.elapsed_time(e1) / a.steps
                print(json.dumps({"flux": flux, "order": order, "tile": int(tc), "nt": int(nt), "ms_per_step": ms,
                                  "gcups": c.mesh.nc * 2 / ms / 1e6, "plan": s.plan_summary}), flush=True)
                s.close()


if __name__ == "__main__":
    main()
