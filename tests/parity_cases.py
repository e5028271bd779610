"""The parity cases shared by oracle/make_golden.py (which runs the REAL reference on them and
commits the outputs under tests/golden/) and by the tests (which rebuild the same deterministic
inputs and compare the oracle port / the CUDA path with those outputs)."""
import numpy as np

from cfd2d_b200 import cases

# name -> (builder, reference variant, flux, order, initial state, nsteps, options)
#   flux: 0 Godunov, 1 Lax-Friedrichs; options: limits override / steady
CASES = {
    # reference as implemented: 2nd-order unlimited reconstruction + exact Riemann solver
    "strip_weak_v0": dict(make=lambda: cases.strip(40, 10, jitter=0.2, shuffle=True), variant="v0", flux=0, order=2,
                          init="region", nsteps=50),
    "channel_2mat_v0": dict(make=lambda: cases.channel(32, 16, jitter=0.2, shuffle=True, two_materials=True),
                            variant="v0", flux=0, order=2, init="smooth", nsteps=50),
    "step_v0": dict(make=lambda: cases.forward_step(30, 10, jitter=0.15), variant="v0", flux=0, order=2,
                    init="smooth", nsteps=50),
    # BASELINE config 1: Sod 10:1, first-order Lax-Friedrichs
    "sod_lf1_v1": dict(make=lambda: cases.strip(40, 10, jump="sod", jitter=0.2), variant="v1", flux=1, order=1,
                       init="region", nsteps=60),
    # 2nd-order + LF
    "channel_lf2_v2": dict(make=lambda: cases.channel(32, 16, jitter=0.2, shuffle=True), variant="v2", flux=1, order=2,
                           init="smooth", nsteps=50),
    # limit flags + remediateLimCells: p max BELOW the initial peak => a cluster of adjacent cells trips at once (in-place sweep order matters)
    "channel_limits_v0": dict(make=lambda: cases.channel(32, 16, jitter=0.1), variant="v0", flux=0, order=2,
                              init="smooth", nsteps=60, limits=[1e-6, 1e6, 1e-3, 1.03e5, 1e6]),
    # steady mode: local time step recomputed every step
    "channel_steady_v0": dict(make=lambda: cases.channel(24, 12, jitter=0.2), variant="v0", flux=0, order=2,
                              init="smooth", nsteps=40, steady=1),
}


def build(name):
    spec = CASES[name]
    c = spec["make"]()
    if "limits" in spec:
        (c.task.ro_min, c.task.ro_max, c.task.p_min, c.task.p_max, c.task.u_max) = spec["limits"]
    if spec.get("steady"):
        c.task.steady = 1
    st = c.smooth_state() if spec["init"] == "smooth" else c.initial_state()
    return c, spec, st


def kat_rim_inputs(n=4000, seed=2024):
    """Seeded left/right states: weak jumps (the smooth-flow regime), strong jumps, sonic and
    supersonic cases, near-vacuum."""
    rng = np.random.default_rng(seed)
    a = np.empty((n, 8))
    a[:, 0] = rng.uniform(0.1, 2, n); a[:, 1] = rng.uniform(1e4, 2e5, n)
    a[:, 2] = rng.uniform(-300, 300, n); a[:, 3] = rng.uniform(-100, 100, n)
    a[:, 4] = rng.uniform(0.1, 2, n); a[:, 5] = rng.uniform(1e4, 2e5, n)
    a[:, 6] = rng.uniform(-300, 300, n); a[:, 7] = rng.uniform(-100, 100, n)
    q = n // 4
    # weak jumps around a common state
    a[:q, 4] = a[:q, 0] * (1 + 1e-3 * rng.standard_normal(q)); a[:q, 5] = a[:q, 1] * (1 + 1e-3 * rng.standard_normal(q))
    a[:q, 6] = a[:q, 2] + 0.1 * rng.standard_normal(q)
    # supersonic both ways
    a[q:q + q // 2, 2] = rng.uniform(800, 1500, q // 2); a[q:q + q // 2, 6] = a[q:q + q // 2, 2] + rng.uniform(-50, 50, q // 2)
    a[q + q // 2:2 * q, 2] = -rng.uniform(800, 1500, q - q // 2); a[q + q // 2:2 * q, 6] = a[q + q // 2:2 * q, 2] + rng.uniform(-50, 50, q - q // 2)
    # strong receding flow -> vacuum branch
    a[2 * q:2 * q + 50, 2] = -9000; a[2 * q:2 * q + 50, 6] = 9000
    return a


def kat_flux_inputs(n=2000, seed=2025):
    rng = np.random.default_rng(seed)
    a = np.empty((n, 12))
    for o in (0, 5):
        a[:, o + 0] = rng.uniform(0.5, 1.5, n); a[:, o + 1] = rng.uniform(5e4, 1.5e5, n)
        a[:, o + 2] = rng.uniform(-200, 200, n); a[:, o + 3] = rng.uniform(-200, 200, n)
        a[:, o + 4] = a[:, o + 1] / (a[:, o + 0] * 0.4) + 0.5 * (a[:, o + 2] ** 2 + a[:, o + 3] ** 2)
    th = rng.uniform(0, 2 * np.pi, n)
    a[:, 10] = np.cos(th); a[:, 11] = np.sin(th)
    return a


def err_norm(a, b):
    """max_c|a-b| / max_c|b| per conservative variable; the momentum
    components share a scale."""
    ro, ru, rv, re = a
    rro, rru, rrv, rre = b
    sm = max(np.abs(rru).max(), np.abs(rrv).max(), 1e-300)
    return (np.abs(ro - rro).max() / np.abs(rro).max(), np.abs(ru - rru).max() / sm,
            np.abs(rv - rrv).max() / sm, np.abs(re - rre).max() / np.abs(rre).max())
