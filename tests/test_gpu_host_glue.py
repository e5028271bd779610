"""GPU (-m gpu): the real drop-in -- the reference's unchanged C++ host (TinyXML task.xml, UNV mesh
reader, CFDBoundary, FVM_TVD::init/save) driving the CUDA library through the FVM_TVD_CUDA glue
(cfd-2d_b200/host), compared with the reference's CPU FVM_TVD run by the SAME binary on the SAME
task.xml + UNV mesh.  The comparison is on the reference's own output artefact: res_%010d.vtk
(FVM_TVD::save prints %25.16f)."""
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from cfd2d_b200 import cases, task as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "cfd-2d_b200", "host", "_build", "cfd2d_cuda")


def read_vtk_cell_data(path):
    txt = open(path).read()
    out = {}
    for m in re.finditer(r"(SCALARS|VECTORS) (\w+) float[^\n]*\n(?:LOOKUP_TABLE default\n)?", txt):
        start = m.end()
        nxt = re.search(r"\n(SCALARS|VECTORS) ", txt[start:])
        body = txt[start: start + nxt.start()] if nxt else txt[start:]
        out[m.group(2)] = np.array(body.split(), dtype=np.float64)
    return out


def run_driver(case, method, workdir, extra_gpu=None):
    case.task.method = method
    case.write(workdir)
    if extra_gpu:
        import xml.etree.ElementTree as ET
        p = os.path.join(workdir, "task.xml")
        tree = ET.parse(p)
        ET.SubElement(tree.getroot(), "gpu", **extra_gpu)
        tree.write(p)
    r = subprocess.run([BIN, "task.xml"], cwd=workdir, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(BIN), reason="host driver not built (needs the reference tree at build time)")
def test_dropin_driver_matches_reference_cpu_method():
    c = cases.strip(60, 16, jitter=0.2, shuffle=True)
    c.task.STEP_MAX = 40
    c.task.FILE_OUTPUT_STEP = 20
    c.task.LOG_OUTPUT_STEP = 10
    d_gpu, d_cpu = tempfile.mkdtemp(), tempfile.mkdtemp()
    out_gpu = run_driver(c, "FVM_TVD_CUDA", d_gpu)
    out_cpu = run_driver(c, "FVM_TVD", d_cpu)
    assert "FVM_TVD_CUDA:" in out_gpu and "sm_100a" in out_gpu
    # same log lines (time step, step counter, simulated time)
    def pick(s):
        return [ln for ln in s.splitlines() if ln.startswith("time step:") or ln.startswith("step:")]
    assert pick(out_gpu) == pick(out_cpu)
    for step in (0, 20, 40):
        a = read_vtk_cell_data(os.path.join(d_gpu, "res_%010d.vtk" % step))
        b = read_vtk_cell_data(os.path.join(d_cpu, "res_%010d.vtk" % step))
        assert set(a) == set(b) and "Density" in a and "Velosity" in a
        for k in b:
            scale = max(np.abs(b[k]).max(), 1e-300)
            assert np.abs(a[k] - b[k]).max() / scale < 1e-12, (step, k)
    # step 0 is written by the reference init itself: byte-identical files
    assert open(os.path.join(d_gpu, "res_0000000000.vtk")).read() == open(os.path.join(d_cpu, "res_0000000000.vtk")).read()


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(BIN), reason="host driver not built")
def test_dropin_driver_lax_variant_is_bit_exact_vs_oracle():
    """<gpu flux="LAX" order="1"/>: BASELINE config 1 (Sod 10:1, first-order Lax-Friedrichs)."""
    from oracle import port as P
    c = cases.strip(60, 16, jump="sod", jitter=0.2)
    c.task.STEP_MAX = 30
    c.task.FILE_OUTPUT_STEP = 30
    d = tempfile.mkdtemp()
    run_driver(c, "FVM_TVD_CUDA", d, extra_gpu=dict(device="0", flux="LAX", order="1"))
    got = read_vtk_cell_data(os.path.join(d, "res_%010d.vtk" % 30))
    o = P.OracleSolver(c.mesh, c.task, 1, 1)
    o.set_state(*c.initial_state())
    o.calc_time_step()
    o.step(30)
    pr = o.get_primitive()
    txt = np.array([float("%25.16f" % x) for x in pr["r"]])
    assert np.array_equal(got["Density"], txt)


@pytest.mark.skipif(not os.path.exists(BIN), reason="host driver not built")
def test_driver_cpu_method_runs_without_gpu():
    """not-gpu sanity: the same binary runs the reference CPU method (no CUDA call on that path)."""
    c = cases.strip(12, 4)
    c.task.STEP_MAX = 3
    d = tempfile.mkdtemp()
    out = run_driver(c, "FVM_TVD", d)
    assert "time step:" in out and os.path.exists(os.path.join(d, "res_0000000000.vtk"))


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(BIN), reason="host driver not built")
def test_python_method_mirror_writes_the_same_vtk_as_the_cpp_dropin():
    """cfd2d_b200.fvm.FVM_TVD (init/run/done over ctypes) vs the C++ glue: same files."""
    from cfd2d_b200 import fvm
    c = cases.channel(40, 20, jitter=0.2, shuffle=True)
    c.task.STEP_MAX = 12
    c.task.FILE_OUTPUT_STEP = 6
    c.task.LOG_OUTPUT_STEP = 4
    d_cpp, d_py = tempfile.mkdtemp(), tempfile.mkdtemp()
    run_driver(c, "FVM_TVD_CUDA", d_cpp)
    c.write(d_py)
    m = fvm.FVM_TVD(workdir=d_py)
    m.init("task.xml")
    m.run()
    m.done()
    for step in (0, 6, 12):
        a = open(os.path.join(d_cpp, "res_%010d.vtk" % step)).read()
        b = open(os.path.join(d_py, "res_%010d.vtk" % step)).read()
        assert a == b, step


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(BIN), reason="host driver not built (needs the reference tree at build time)")
@pytest.mark.parametrize("nranks", [2, 4])
def test_dropin_driver_multi_rank_equals_single_rank(nranks):
    """Multi-GPU THROUGH the reference's Method boundary: N processes of the same cfd2d_cuda binary (one per
    GPU, tools/launch_ranks.py), each running the reference's init on the global mesh, recomputing Decomp's
    partition (bundled METIS) and renumbering in the glue, cross-checking it against the mesh/mesh.NNNN.proc
    files the reference's own DECOMP method wrote, exchanging halos over NCCL, gathering on rank 0 which
    writes res_*.vtk with the reference's writer.  The files must be BYTE-identical to the 1-GPU run's.
    Needs >= nranks GPUs (skipped otherwise: NCCL refuses two ranks on one device)."""
    if _ngpu() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    import ctypes
    import sys
    import xml.etree.ElementTree as ET
    c = cases.channel(64, 32, jitter=0.2, shuffle=True)
    c.task.p_max = 1.01e5                      # the limiter trips around the bump: remediation across the cut
    c.task.STEP_MAX = 30
    c.task.FILE_OUTPUT_STEP = 10
    c.task.LOG_OUTPUT_STEP = 10
    d1, dn = tempfile.mkdtemp(), tempfile.mkdtemp()
    run_driver(c, "FVM_TVD_CUDA", d1)
    c.task.method = "FVM_TVD_CUDA"
    c.write(dn)
    # the reference's own DECOMP method -> mesh/mesh.NNNN.proc (oracle/_ref ships prebuilt)
    dlib = os.path.join(ROOT, "oracle", "_ref", "libcfd2d_ref_decomp.so")
    if os.path.exists(dlib):
        tree = ET.parse(os.path.join(dn, "task.xml"))
        dec = ET.SubElement(tree.getroot(), "decomp")
        ET.SubElement(dec, "processors", value=str(nranks))
        tree.write(os.path.join(dn, "task_decomp.xml"))
        os.makedirs(os.path.join(dn, "mesh"), exist_ok=True)
        cwd = os.getcwd()
        try:
            assert ctypes.CDLL(dlib, mode=os.RTLD_LOCAL).ref_decomp_run(dn.encode(), b"task_decomp.xml") == 0
        finally:
            os.chdir(cwd)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_ranks.py"), str(nranks), BIN, "task.xml"],
                       cwd=dn, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"rank 0/{nranks}" in r.stdout
    if os.path.exists(dlib):
        assert "partition maps equal mesh/mesh.0000.proc" in r.stdout
    for step in (0, 10, 20, 30):
        a = open(os.path.join(dn, "res_%010d.vtk" % step)).read()
        b = open(os.path.join(d1, "res_%010d.vtk" % step)).read()
        assert a == b, step
    fin = read_vtk_cell_data(os.path.join(dn, "res_%010d.vtk" % 30))
    assert np.isfinite(fin["Density"]).all()
