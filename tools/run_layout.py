#!/usr/bin/env python
"""Run a few RK2 steps of one step layout at the bench workload (for ncu / compute-sanitizer captures).
  python tools/run_layout.py --flux lax --order 2 --layout 2 --steps 3 [--nx 2000 --ny 1000]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--flux", default="godunov", choices=["godunov", "lax"])
    ap.add_argument("--order", type=int, default=2)
    ap.add_argument("--layout", type=int, default=0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--nx", type=int, default=2000)
    ap.add_argument("--ny", type=int, default=1000)
    ap.add_argument("--exact-riemann", action="store_true")
    ap.add_argument("--p-max", type=float, default=None, help="lower the pressure limit (trips the limiter / remediation)")
    ap.add_argument("--no-graph", action="store_true")
    a = ap.parse_args()
    from cfd2d_b200 import cases, fvm
    c = cases.channel(a.nx, a.ny)
    if a.p_max is not None:
        c.task.p_max = a.p_max
    st = c.smooth_state()
    s = fvm.Solver(c.mesh, c.task, 0 if a.flux == "godunov" else 1, a.order)
    s.use_exact_riemann(a.exact_riemann)
    s.use_fused(a.layout)
    if a.no_graph:
        s.use_graph(False)
    s.set_state(*st)
    s.calc_time_step()
    s.step(a.steps)
    got = s.get_state()
    print("layout", a.layout, "flux", a.flux, "order", a.order, "steps", a.steps, "plan:", s.plan_summary,
          "| checksum", float(got[0].sum()), "flagged", int((got[5] != 0).sum()))
    s.close()


if __name__ == "__main__":
    main()
