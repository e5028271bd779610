"""GPU (-m gpu): parity at the sizes BASELINE.json names, against the REAL reference run beside the
GPU on the box's host cores (oracle/_ref/libcfd2d_ref_v*.so = zhrv/cfd-2d's FVM_TVD compiled by
oracle/Makefile; it ships prebuilt with the snapshot -- /root/reference itself is not needed here).

  configs[0]  Sod 10:1 strip, 200x50x2 = 20 k cells, first-order Lax-Friedrichs   -> bit-exact (v1)
  configs[1]  forward-facing step outline, 315 k cells, 2nd order + exact Godunov  -> <= 1e-12 (v0)
  configs[1]  1 M-cell channel, 2nd order + Lax-Friedrichs                         -> bit-exact (v2)

The reference reads the same UNV + task.xml the case writes; the GPU side gets the flattened mesh of
cfd2d_b200.mesh (pinned equal to the reference reader's by tests/test_mesh_task.py)."""
import os
import tempfile

import numpy as np
import pytest

import parity_cases as pc
from cfd2d_b200 import cases, fvm
from oracle import refharness as R

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _ref(case, variant, state):
    if not R.available(variant):
        pytest.skip(f"oracle/_ref/libcfd2d_ref_{variant}.so not built (run __graft_entry__.build() where /root/reference exists)")
    d = tempfile.mkdtemp(prefix="cfd2d_ref_")
    case.write(d)
    s = R.RefSolver(d, variant=variant)
    s.set_state(*state)
    return s


def test_config0_sod_strip_20k_lf1_bit_exact():
    c = cases.strip(200, 50, jump="sod")                 # 20 000 cells, regions 10:1, all walls
    st = c.initial_state()
    r = _ref(c, "v1", st)
    g = fvm.Solver(c.mesh, c.task, fvm.FLUX_LAX, 1)
    g.set_state(*st)
    assert g.calc_time_step() == r.calc_time_step()
    for _ in range(2):
        g.step(100)
        r.run(100)
        got, ref = g.get_state(), r.state()
        for k in range(6):                               # ro, ru, rv, re, cTau, flag
            assert np.array_equal(got[k], ref[k]), k
    assert np.abs(got[1]).max() > 1.0                    # the shock tube really runs
    g.close()


def test_config1_forward_step_315k_godunov2_vs_reference():
    c = cases.forward_step(750, 250, jitter=0.15)
    assert c.mesh.nc >= 250000
    st = c.smooth_state()
    r = _ref(c, "v0", st)
    g = fvm.Solver(c.mesh, c.task)                       # the reference's live scheme
    g.set_state(*st)
    assert g.calc_time_step() == r.calc_time_step()
    assert np.array_equal(g.calc_grad(), r.calc_grad())  # no exp/log before the flux: bit-exact
    g.step(20)
    r.run(20)
    got, ref = g.get_state(), r.state()
    e = pc.err_norm(got[:4], ref[:4])
    print("315k-cell forward step, 20 steps, Godunov order 2: rel. Linf per variable =", e)
    assert max(e) < TOL, e
    assert np.array_equal(got[5], ref[5])
    g.close()


def test_config1_channel_1m_lf2_bit_exact():
    c = cases.channel(1000, 500)                         # 1 000 000 cells
    st = c.smooth_state()
    r = _ref(c, "v2", st)
    g = fvm.Solver(c.mesh, c.task, fvm.FLUX_LAX, 2)
    g.set_state(*st)
    assert g.calc_time_step() == r.calc_time_step()
    g.step(5)
    r.run(5)
    got, ref = g.get_state(), r.state()
    for k in range(6):
        assert np.array_equal(got[k], ref[k]), k
    # the tile-fused layouts (1: k_stage, 2: pipelined k_stage_pipe) must give the same bits at this size
    for mode in (1, 2):
        g.set_state(*st)
        g.use_fused(mode)
        g.step(5)
        got2 = g.get_state()
        for k in range(6):
            assert np.array_equal(got2[k], ref[k]), (mode, k)
    g.close()
