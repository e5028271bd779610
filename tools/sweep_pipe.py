#!/usr/bin/env python
"""Sweep the step layouts at the bench workload (4 M cells): three sweeps (exact / reduced-instruction
Riemann solver), k_stage, and k_stage_pipe over tile size (CFD2D_PIPE_TILE), block size (CFD2D_PIPE_NT)
and resident CTAs per SM (CFD2D_PIPE_CTAS).  One JSON line per configuration: ms per RK2 step
(device-resident, CUDA events on the launching stream), per-kernel times, 320 B/cell roofline fraction."""
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
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--pipe", default="128:256:2,128:384:2,128:512:2,256:512:1,256:768:1,256:1024:1,192:768:1",
                    help="TC:NT:MINB list (tile cells : threads per CTA : launch-bounds min CTAs per SM)")
    ap.add_argument("--variants", default="0:2,1:2,1:1,0:1")
    ap.add_argument("--skip-base", action="store_true")
    a = ap.parse_args()
    import torch
    from cfd2d_b200 import cases, fvm
    peak = 6553.3
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    c = cases.channel(a.nx, a.ny)
    st = c.smooth_state()
    nc = c.mesh.nc
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def run(tag, flux, order, layout, exact=False, env=None):
        for k, v in (env or {}).items():
            os.environ[k] = str(v)
        try:
            s = fvm.Solver(c.mesh, c.task, flux, order)
            s.use_exact_riemann(exact)
            s.use_fused(layout)
            s.set_stream(stream.cuda_stream)
            s.set_state(*st)
            s.calc_time_step()
            s.step(10)
            torch.cuda.synchronize()
            e0.record(stream); s.step_async(a.steps); e1.record(stream); s.sync()
            ms = e0.elapsed_time(e1) / a.steps
            p = s.profile(4)
            per = {k: v[0] / v[1] for k, v in p.items() if v[1]}
            out = {"tag": tag, "flux": flux, "order": order, "layout": layout, "exact_riemann": exact, "env": env or {},
                   "ms_per_step": ms, "G_cell_updates_s": nc * 2.0 / (ms * 1e-3) / 1e9,
                   "roofline_frac_320B": 320.0 * nc * 2.0 / (ms * 1e-3) / 1e9 / peak, "per_kernel_ms": per,
                   "plan": s.plan_summary if layout else ""}
            s.close()
        except Exception as ex:  # keep sweeping
            out = {"tag": tag, "flux": flux, "order": order, "layout": layout, "env": env or {}, "error": repr(ex)}
        print(json.dumps(out), flush=True)

    variants = [tuple(int(x) for x in v.split(":")) for v in a.variants.split(",")]
    for flux, order in variants:
        if not a.skip_base:
            run("sweeps", flux, order, 0)
            if flux == 0:
                run("sweeps_exact_riemann", flux, order, 0, exact=True)
            if (flux, order) == (1, 1):
                run("sweeps_no_lf1cell", flux, order, 0, env={"CFD2D_LF1_CELL": 0})
                os.environ["CFD2D_LF1_CELL"] = "1"
        for cfg in a.pipe.split(","):
            tc, nt, minb = (int(x) for x in cfg.split(":"))
            run("pipe", flux, order, 2, env={"CFD2D_PIPE_TILE": tc, "CFD2D_PIPE_NT": nt, "CFD2D_PIPE_MINB": minb})
        if flux == 0 and order == 2:
            run("pipe_exact_riemann", flux, order, 2, exact=True, env={"CFD2D_PIPE_TILE": 128, "CFD2D_PIPE_NT": 256, "CFD2D_PIPE_MINB": 2})


if __name__ == "__main__":
    main()
