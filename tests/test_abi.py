"""CPU: the C-ABI shared library loads and exports every symbol include/cfd2d_fvm.h declares; the
product path fails loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from cfd2d_b200 import cases, fvm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    h = open(os.path.join(ROOT, "include", "cfd2d_fvm.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(cfd2d_[a-z0-9_]+)\s*\(", h)))


def test_header_symbols_are_exported():
    if not os.path.exists(fvm.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(fvm.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cfd2d_fvm.h but not exported"
    assert set(fvm.EXPORTS) <= set(names)


def test_version_string():
    lib = fvm.load_library()
    assert b"sm_100a" in lib.cfd2d_version()


def test_null_handle_queries_are_safe():
    """the query entry points take a NULL handle without touching the device"""
    lib = fvm.load_library()
    assert lib.cfd2d_fvm_halo_transport(None) == 0          # no handle: no halo, no transport
    assert lib.cfd2d_fvm_launch_count(None) == 0
    assert lib.cfd2d_fvm_plan_summary(None) in (b"", None)


def _gpu():
    import torch
    return torch.cuda.is_available()


@pytest.mark.skipif(_gpu(), reason="only meaningful without a GPU")
def test_no_cpu_fallback_without_gpu():
    c = cases.strip(4, 2)
    with pytest.raises(fvm.CFDError) as e:
        fvm.Solver(c.mesh, c.task)
    assert e.value.code == -2          # CFD2D_ENODEV


def test_create_rejects_bad_arguments():
    c = cases.strip(4, 2)
    with pytest.raises(fvm.CFDError) as e:
        fvm.Solver(c.mesh, c.task, order=3)
    assert e.value.code == -1
    m = c.mesh
    bad = np.array(m.edge_bc, copy=True)
    bad[m.edge_c2 < 0] = -1
    import dataclasses
    m2 = dataclasses.replace(m, edge_bc=bad)
    with pytest.raises(fvm.CFDError) as e:
        fvm.Solver(m2, c.task)
    assert e.value.code == -5          # CFD2D_EBC, fvm_tvd.cpp:706-710


def test_create_rejects_unsorted_cell_edges():
    """The slot order of cell_edges is the summation order (ascending edge id, as the reference's
    readers produce it): a caller with another order is refused instead of silently getting other bits.
    The check runs before any device work, so it is testable without a GPU."""
    import dataclasses
    c = cases.strip(4, 2)
    ce = np.array(c.mesh.cell_edges, copy=True)
    ce[0] = ce[0][::-1]
    with pytest.raises(fvm.CFDError) as e:
        fvm.Solver(dataclasses.replace(c.mesh, cell_edges=ce), c.task)
    assert e.value.code == -1 and "ascending" in str(e.value)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the
    contract's keys, the reference's own single-thread rate as `value`, zero copy bytes in `e2e`."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "cell-updates/s" and d["higher_is_better"] is True
    assert d["value"] > 1e5 and d["dtype"] == "f64"
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
