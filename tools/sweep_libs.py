#!/usr/bin/env python
"""Kernel-variant sweep: times the three-sweep layout with differently compiled libraries
(CFD2D_LIB=<path>), one subprocess per library.  python tools/sweep_libs.py lib1.so lib2.so ...
(SWEEP_CASES="flux:order,..." selects the schemes, default "0:2,1:2")"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, os, sys
sys.path.insert(0, %r)
import torch
from cfd2d_b200 import cases, fvm
c = cases.channel(2000, 1000); st = c.smooth_state(); nc = c.mesh.nc
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
cases_ = [tuple(int(v) for v in x.split(':')) for x in os.environ.get('SWEEP_CASES', '0:2,1:2').split(',')]
for flux, order in cases_:
    s = fvm.Solver(c.mesh, c.task, flux, order)
    s.set_stream(stream.cuda_stream); s.set_state(*st); s.calc_time_step(); s.step(10)
    torch.cuda.synchronize(); e0.record(stream); s.step_async(40); e1.record(stream); s.sync()
    ms = e0.elapsed_time(e1) / 40
    p = s.profile(4)
    print(json.dumps({"lib": os.path.basename(os.environ.get("CFD2D_LIB", "default")), "flux": flux, "order": order, "ms_per_step": ms,
                      "per_kernel_ms": {k: round(v[0] / v[1], 4) for k, v in p.items() if v[1]}}), flush=True)
    s.close()
''' % ROOT
for lib in sys.argv[1:]:
    env = dict(os.environ, CFD2D_LIB=os.path.abspath(lib))
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    sys.stdout.write(r.stdout)
    if r.returncode:
        sys.stdout.write(json.dumps({"lib": lib, "error": r.stderr[-400:]}) + "\n")
    sys.stdout.flush()
