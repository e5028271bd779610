"""Generate tests/golden/*.npz by running the REAL reference (oracle/_ref, compiled from
/root/reference by oracle/Makefile) on the deterministic inputs of tests/parity_cases.py.
Run in the build container only:  python oracle/make_golden.py
TEST INFRASTRUCTURE ONLY."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import parity_cases as pc  # noqa: E402
from oracle import refharness as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in pc.CASES:
        c, spec, st = pc.build(name)
        d = tempfile.mkdtemp()
        c.write(d)
        s = R.RefSolver(d, variant=spec["variant"])
        s.set_state(*st)
        tau = s.calc_time_step()
        grad = s.calc_grad()
        flux = s.edge_fluxes()
        out = dict(tau=tau, grad0=grad, flux0=flux)
        if name == "strip_weak_v0":   # one full reference-reader mesh, to pin cfd2d_b200.mesh.build_mesh
            for k, v in s.mesh().items():
                out["mesh_" + k] = v
        half = spec["nsteps"] // 2
        s.run(half)
        mid = s.state()
        s.run(spec["nsteps"] - half)
        fin = s.state()
        for tag, stt in (("mid", mid), ("fin", fin)):
            for k, v in zip(("ro", "ru", "rv", "re", "ctau", "flag"), stt):
                out[f"{tag}_{k}"] = v
        out["nsteps"] = spec["nsteps"]
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, "nc", s.nc, "flagged mid/fin", int((mid[5] & 2 > 0).sum()), int((fin[5] & 2 > 0).sum()),
              "touched", int((fin[5] != 0).sum()), "tau", tau)
    a = pc.kat_rim_inputs()
    # the reference's Newton loop has no cap: check with the capped port first that every
    # input terminates before handing it to the real rim_orig
    from oracle import port as P
    _, it = P.rim_orig(a, max_newton=200)
    assert (it >= 0).all(), "kat_rim_inputs contains a non-terminating state: %s" % np.nonzero(it < 0)[0][:10]
    np.savez_compressed(os.path.join(OUT, "kat_rim_orig.npz"), out=R.rim_orig(a))
    f = pc.kat_flux_inputs()
    np.savez_compressed(os.path.join(OUT, "kat_calc_flux.npz"), godunov=R.calc_flux(f, variant="v0"), lax=R.calc_flux(f, variant="v1"))
    rng = np.random.default_rng(7)
    io = np.abs(rng.standard_normal((500, 8))) * [1, 1e5, 2e5, 3e5, 100, 100, 300, 300] + [0.1, 1e3, 1e3, 1e3, 0, 0, 1, 100]
    np.savez_compressed(os.path.join(OUT, "kat_urs.npz"), inp=io, m0=R.urs(io, 0.02898, 1004.5, 0), m1=R.urs(io, 0.02898, 1004.5, 1),
                        m2=R.urs(io, 0.02898, 1004.5, 2))
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
