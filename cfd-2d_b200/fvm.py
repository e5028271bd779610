"""ctypes binding of the C-ABI (include/cfd2d_fvm.h) and the host-side mirror of the reference's
``FVM_TVD`` method object (``init`` / ``run`` / ``done`` / ``save``,
reference ``src/methods/fvm_tvd.{h,cpp}``, interface ``src/methods/method.h:6-12``).

Everything numerical happens in ``csrc/libcfd2d_b200.so`` (hand-written sm_100a CUDA).  There is
no CPU fallback: if the library is missing or no GPU is usable, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import mesh as _mesh
from . import task as _task

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CFD2D_LIB") or os.path.join(HERE, "csrc", "libcfd2d_b200.so")   # env: kernel-variant sweeps only

FLUX_GODUNOV, FLUX_LAX = 0, 1
K_NAMES = ["grad", "flux", "update1", "update2", "remediate", "timestep", "halo", "stage1", "stage2"]
NKERNELS = len(K_NAMES)

ERRORS = {0: "OK", -1: "EINVAL", -2: "ENODEV", -3: "ECUDA", -4: "ENEWTON", -5: "EBC", -6: "ENCCL"}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_up = C.POINTER(C.c_uint32)


class CMesh(C.Structure):
    _fields_ = [("nc", C.c_int32), ("nc_ex", C.c_int32), ("ne", C.c_int32),
                ("cell_S", _dp), ("cell_cx", _dp), ("cell_cy", _dp), ("cell_mat", _ip), ("cell_edges", _ip),
                ("edge_c1", _ip), ("edge_c2", _ip), ("edge_nx", _dp), ("edge_ny", _dp), ("edge_l", _dp),
                ("edge_gp", _dp), ("edge_bc", _ip)]


class CPhys(C.Structure):
    _fields_ = [("nmat", C.c_int32), ("mat_M", _dp), ("mat_Cp", _dp), ("nbc", C.c_int32),
                ("bc_kind", _ip), ("bc_par", _dp), ("limits", C.c_double * 5)]


class CCtrl(C.Structure):
    _fields_ = [("CFL", C.c_double), ("TAU", C.c_double), ("steady", C.c_int32), ("flux", C.c_int32),
                ("order", C.c_int32), ("max_newton", C.c_int32)]


class CHalo(C.Structure):
    _fields_ = [("rank", C.c_int32), ("nranks", C.c_int32), ("recv_count", _ip), ("send_count", _ip),
                ("send_ind", _ip), ("nccl_unique_id", C.c_void_p), ("cell_gid", _ip)]


class CFDError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__(f"cfd2d error {ERRORS.get(code, code)}: {msg}")
        self.code = code


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Packed:
    """Holds the numpy arrays alive behind a (CMesh, CPhys, CCtrl[, CHalo]) quadruple."""

    def __init__(self, m: _mesh.Mesh | dict, t: _task.Task, flux=FLUX_GODUNOV, order=2, max_newton=0,
                 nc_owned: int | None = None):
        g = m if isinstance(m, dict) else {k: getattr(m, k) for k in
                                            ("cell_S", "cell_cx", "cell_cy", "cell_mat", "cell_edges", "edge_c1", "edge_c2",
                                             "edge_nx", "edge_ny", "edge_l", "edge_gp", "edge_bc")}
        self.a = dict(
            cell_S=_f64(g["cell_S"]), cell_cx=_f64(g["cell_cx"]), cell_cy=_f64(g["cell_cy"]),
            cell_mat=_i32(g["cell_mat"]), cell_edges=_i32(g["cell_edges"]),
            edge_c1=_i32(g["edge_c1"]), edge_c2=_i32(g["edge_c2"]), edge_nx=_f64(g["edge_nx"]),
            edge_ny=_f64(g["edge_ny"]), edge_l=_f64(g["edge_l"]), edge_gp=_f64(g["edge_gp"]),
            edge_bc=_i32(g["edge_bc"]))
        a = self.a
        nc_ex = a["cell_S"].shape[0]
        nc = nc_ex if nc_owned is None else int(nc_owned)
        self.mesh = CMesh(nc, nc_ex, a["edge_c1"].shape[0],
                          a["cell_S"].ctypes.data_as(_dp), a["cell_cx"].ctypes.data_as(_dp),
                          a["cell_cy"].ctypes.data_as(_dp), a["cell_mat"].ctypes.data_as(_ip),
                          a["cell_edges"].ctypes.data_as(_ip), a["edge_c1"].ctypes.data_as(_ip),
                          a["edge_c2"].ctypes.data_as(_ip), a["edge_nx"].ctypes.data_as(_dp),
                          a["edge_ny"].ctypes.data_as(_dp), a["edge_l"].ctypes.data_as(_dp),
                          a["edge_gp"].ctypes.data_as(_dp), a["edge_bc"].ctypes.data_as(_ip))
        self.mat_M = _f64([mm.M for mm in t.materials])
        self.mat_Cp = _f64([mm.Cp for mm in t.materials])
        nb = len(t.boundaries)
        self.bc_kind = _i32([b.kind for b in t.boundaries] or [0])
        self.bc_par = _f64([b.par for b in t.boundaries] or [[0, 0, 0, 0]])
        self.phys = CPhys(len(t.materials), self.mat_M.ctypes.data_as(_dp), self.mat_Cp.ctypes.data_as(_dp),
                          nb, self.bc_kind.ctypes.data_as(_ip), self.bc_par.ctypes.data_as(_dp),
                          (C.c_double * 5)(*t.limits))
        self.ctrl = CCtrl(t.CFL, t.TAU, int(t.steady), int(flux), int(order), int(max_newton))
        self.nc, self.nc_ex, self.ne = nc, nc_ex, self.mesh.ne


_LIB = None


def load_library() -> C.CDLL:
    """Load libcfd2d_b200.so (built by __graft_entry__.build()).  Fails loudly when absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise CFDError(-2, f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    H = C.c_void_p
    lib.cfd2d_fvm_create.argtypes = [C.POINTER(CMesh), C.POINTER(CPhys), C.POINTER(CCtrl), C.POINTER(CHalo), C.c_int, C.POINTER(H)]
    lib.cfd2d_fvm_destroy.argtypes = [H]
    lib.cfd2d_fvm_destroy.restype = None
    lib.cfd2d_fvm_set_state.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.cfd2d_fvm_calc_time_step.argtypes = [H, _dp]
    lib.cfd2d_fvm_step.argtypes = [H, C.c_int]
    lib.cfd2d_fvm_step_async.argtypes = [H, C.c_int]
    lib.cfd2d_fvm_sync.argtypes = [H]
    lib.cfd2d_fvm_get_state.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.cfd2d_fvm_get_primitive.argtypes = [H] + [C.c_void_p] * 6
    lib.cfd2d_fvm_tau.argtypes = [H]
    lib.cfd2d_fvm_tau.restype = C.c_double
    lib.cfd2d_fvm_time.argtypes = [H]
    lib.cfd2d_fvm_time.restype = C.c_double
    lib.cfd2d_fvm_calc_grad.argtypes = [H, _dp]
    lib.cfd2d_fvm_edge_fluxes.argtypes = [H, _dp]
    lib.cfd2d_kat_rim_orig.argtypes = [C.c_int, C.c_int, _dp, C.c_double, C.c_int, _dp, _ip]
    lib.cfd2d_kat_calc_flux.argtypes = [C.c_int, C.c_int, _dp, C.c_double, C.c_int, _dp]
    lib.cfd2d_kat_rim_orig_fast.argtypes = [C.c_int, C.c_int, _dp, C.c_int, _dp, _ip]
    lib.cfd2d_fvm_use_exact_riemann.argtypes = [H, C.c_int]
    lib.cfd2d_fvm_snapshot_begin.argtypes = [H]
    lib.cfd2d_fvm_snapshot_end.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.cfd2d_kat_urs.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, _dp]
    lib.cfd2d_fvm_profile.argtypes = [H, C.c_int, _dp, C.POINTER(C.c_int64)]
    lib.cfd2d_fvm_launch_count.argtypes = [H]
    lib.cfd2d_fvm_launch_count.restype = C.c_int64
    lib.cfd2d_fvm_halo_transport.argtypes = [H]
    lib.cfd2d_fvm_halo_transport.restype = C.c_int
    lib.cfd2d_fvm_set_stream.argtypes = [H, C.c_void_p]
    lib.cfd2d_fvm_use_graph.argtypes = [H, C.c_int]
    lib.cfd2d_fvm_use_fused.argtypes = [H, C.c_int]
    lib.cfd2d_fvm_plan_summary.argtypes = [H]
    lib.cfd2d_fvm_plan_summary.restype = C.c_char_p
    lib.cfd2d_tiling_plan.argtypes = [C.POINTER(CMesh), C.c_int, C.c_int, _ip, C.POINTER(C.c_int64)]
    lib.cfd2d_pipe_plan.argtypes = [C.POINTER(CMesh), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64)]
    lib.cfd2d_fvm_last_error.argtypes = [H]
    lib.cfd2d_fvm_last_error.restype = C.c_char_p
    lib.cfd2d_version.restype = C.c_char_p
    _LIB = lib
    return lib


EXPORTS = [
    "cfd2d_fvm_create", "cfd2d_fvm_destroy", "cfd2d_fvm_set_state", "cfd2d_fvm_calc_time_step", "cfd2d_fvm_step",
    "cfd2d_fvm_step_async", "cfd2d_fvm_sync", "cfd2d_fvm_get_state", "cfd2d_fvm_get_primitive", "cfd2d_fvm_tau",
    "cfd2d_fvm_time", "cfd2d_fvm_calc_grad", "cfd2d_fvm_edge_fluxes", "cfd2d_kat_rim_orig", "cfd2d_kat_calc_flux", "cfd2d_kat_urs",
    "cfd2d_kat_rim_orig_fast", "cfd2d_fvm_use_exact_riemann", "cfd2d_fvm_snapshot_begin", "cfd2d_fvm_snapshot_end", "cfd2d_fvm_gather_state",
    "cfd2d_fvm_profile", "cfd2d_fvm_launch_count", "cfd2d_fvm_halo_transport", "cfd2d_fvm_set_stream", "cfd2d_fvm_use_graph",
    "cfd2d_fvm_use_fused", "cfd2d_fvm_plan_summary", "cfd2d_tiling_plan", "cfd2d_pipe_plan",
    "cfd2d_unv_read", "cfd2d_unv_counts", "cfd2d_unv_copy", "cfd2d_unv_group_name", "cfd2d_unv_group_counts",
    "cfd2d_unv_group_copy", "cfd2d_unv_free",
    "cfd2d_fvm_last_error", "cfd2d_version",
]


class Solver:
    """Thin object wrapper over one cfd2d_fvm handle (one GPU)."""

    def __init__(self, m, t: _task.Task, flux=FLUX_GODUNOV, order=2, max_newton=0, device=0,
                 nc_owned=None, halo: dict | None = None):
        self.lib = load_library()
        self.pk = Packed(m, t, flux, order, max_newton, nc_owned)
        self.nc, self.ne = self.pk.nc, self.pk.ne
        self.h = C.c_void_p()
        halo_ref = None
        if halo is not None:
            self._halo_arrays = dict(recv_count=_i32(halo["recv_count"]), send_count=_i32(halo["send_count"]),
                                     send_ind=_i32(halo["send_ind"] if len(halo["send_ind"]) else [0]))
            self._nccl_id = C.create_string_buffer(bytes(halo["nccl_unique_id"]), 128)
            gid = halo.get("cell_gid")
            if gid is not None:
                self._halo_arrays["cell_gid"] = _i32(gid)
            self._halo = CHalo(int(halo["rank"]), int(halo["nranks"]),
                               self._halo_arrays["recv_count"].ctypes.data_as(_ip),
                               self._halo_arrays["send_count"].ctypes.data_as(_ip),
                               self._halo_arrays["send_ind"].ctypes.data_as(_ip),
                               C.cast(self._nccl_id, C.c_void_p),
                               self._halo_arrays["cell_gid"].ctypes.data_as(_ip) if gid is not None else None)
            halo_ref = C.byref(self._halo)
        rc = self.lib.cfd2d_fvm_create(C.byref(self.pk.mesh), C.byref(self.pk.phys), C.byref(self.pk.ctrl),
                                       halo_ref, int(device), C.byref(self.h))
        if rc != 0:
            msg = self.lib.cfd2d_fvm_last_error(None)
            self.h = None
            raise CFDError(rc, msg.decode() if msg else "")

    def _chk(self, rc):
        if rc != 0:
            msg = self.lib.cfd2d_fvm_last_error(self.h)
            raise CFDError(rc, msg.decode() if msg else "")

    @staticmethod
    def _ptr(a):
        if a is None:
            return None
        if isinstance(a, int):
            return C.c_void_p(a)
        return C.c_void_p(a.ctypes.data)

    def set_state(self, ro, ru, rv, re, flag=None):
        """Arrays are numpy float64 (or raw host addresses, e.g. pinned torch tensors' data_ptr())."""
        args = [x if isinstance(x, int) else _f64(x) for x in (ro, ru, rv, re)]
        fl = None if flag is None else (flag if isinstance(flag, int) else np.ascontiguousarray(flag, np.uint32))
        self._keep = (args, fl)
        self._chk(self.lib.cfd2d_fvm_set_state(self.h, *[self._ptr(x) for x in args], self._ptr(fl)))

    def calc_time_step(self) -> float:
        tau = C.c_double()
        self._chk(self.lib.cfd2d_fvm_calc_time_step(self.h, C.byref(tau)))
        return tau.value

    def step(self, nsteps=1):
        self._chk(self.lib.cfd2d_fvm_step(self.h, int(nsteps)))

    def step_async(self, nsteps=1):
        self._chk(self.lib.cfd2d_fvm_step_async(self.h, int(nsteps)))

    def sync(self):
        self._chk(self.lib.cfd2d_fvm_sync(self.h))

    def get_state(self, out=None, want_tau=True, want_flag=True):
        n = self.nc
        if out is None:
            ro, ru, rv, re = (np.empty(n) for _ in range(4))
        else:
            ro, ru, rv, re = out
        ct = np.empty(n) if want_tau else None
        fl = np.empty(n, np.uint32) if want_flag else None
        self._chk(self.lib.cfd2d_fvm_get_state(self.h, self._ptr(ro), self._ptr(ru), self._ptr(rv), self._ptr(re),
                                               self._ptr(ct), self._ptr(fl)))
        return ro, ru, rv, re, ct, fl

    def get_primitive(self):
        n = self.nc
        arrs = [np.empty(n) for _ in range(6)]
        self._chk(self.lib.cfd2d_fvm_get_primitive(self.h, *[self._ptr(x) for x in arrs]))
        return dict(zip(("r", "p", "T", "u", "v", "cz"), arrs))

    @property
    def tau(self) -> float:
        return self.lib.cfd2d_fvm_tau(self.h)

    @property
    def time(self) -> float:
        return self.lib.cfd2d_fvm_time(self.h)

    def calc_grad(self):
        g = np.empty((self.nc, 8))
        self._chk(self.lib.cfd2d_fvm_calc_grad(self.h, g.ctypes.data_as(_dp)))
        return g

    def edge_fluxes(self):
        f = np.empty((self.ne, 4))
        self._chk(self.lib.cfd2d_fvm_edge_fluxes(self.h, f.ctypes.data_as(_dp)))
        return f

    def profile(self, nsteps=5):
        ms = np.zeros(NKERNELS)
        cnt = np.zeros(NKERNELS, np.int64)
        self._chk(self.lib.cfd2d_fvm_profile(self.h, int(nsteps), ms.ctypes.data_as(_dp),
                                             cnt.ctypes.data_as(C.POINTER(C.c_int64))))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(K_NAMES)}

    @property
    def launch_count(self) -> int:
        return int(self.lib.cfd2d_fvm_launch_count(self.h))

    @property
    def halo_transport(self) -> str:
        """how Method::exchange moves halo records on this handle"""
        return {0: "none", 1: "nccl", 2: "peer-store"}[int(self.lib.cfd2d_fvm_halo_transport(self.h))]

    def set_stream(self, stream_ptr: int):
        self._chk(self.lib.cfd2d_fvm_set_stream(self.h, C.c_void_p(stream_ptr)))

    def use_graph(self, on: bool):
        self._chk(self.lib.cfd2d_fvm_use_graph(self.h, 1 if on else 0))

    def use_fused(self, on):
        """Step layout: 0 / False = three sweeps per stage, 1 / True = one tile-fused kernel per RK stage
        (k_stage), 2 = the persistent, bulk-copy-fed tile kernel (k_stage_pipe).  All give the same bits."""
        self._chk(self.lib.cfd2d_fvm_use_fused(self.h, int(on)))

    def snapshot_begin(self):
        """Start an asynchronous copy of the current state to the host (FVM_TVD::save pipeline)."""
        self._chk(self.lib.cfd2d_fvm_snapshot_begin(self.h))

    def snapshot_end(self):
        n = self.nc
        ro, ru, rv, re, ct = (np.empty(n) for _ in range(5))
        fl = np.empty(n, np.uint32)
        self._chk(self.lib.cfd2d_fvm_snapshot_end(self.h, *[C.c_void_p(x.ctypes.data) for x in (ro, ru, rv, re, ct, fl)]))
        return ro, ru, rv, re, ct, fl

    def use_exact_riemann(self, on: bool):
        """Godunov handles: True = rim_orig evaluated in the reference's operation order (rim_orig_dev),
        False (default) = the reduced-instruction solver (csrc/fvm_riemann_fast.cuh)."""
        self._chk(self.lib.cfd2d_fvm_use_exact_riemann(self.h, 1 if on else 0))

    @property
    def plan_summary(self) -> str:
        return self.lib.cfd2d_fvm_plan_summary(self.h).decode()

    def close(self):
        if getattr(self, "h", None):
            self.lib.cfd2d_fvm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def tiling_plan(m, t: _task.Task, tile_cells=512, hilbert=True, nc_owned=None):
    """Host-only: the cell renumbering and tile-plan statistics create() would use (CPU test hook)."""
    lib = load_library()
    pk = Packed(m, t, nc_owned=nc_owned)
    perm = np.empty(pk.nc_ex, np.int32)
    stats = np.zeros(8, np.int64)
    rc = lib.cfd2d_tiling_plan(C.byref(pk.mesh), int(tile_cells), 1 if hilbert else 0, perm.ctypes.data_as(_ip),
                               stats.ctypes.data_as(C.POINTER(C.c_int64)))
    if rc != 0:
        msg = lib.cfd2d_fvm_last_error(None)
        raise CFDError(rc, msg.decode() if msg else "")
    keys = ("ntiles", "nl_max", "ne_max", "sum_ng", "sum_ne", "sum_ring", "interior", "boundary")
    return perm, dict(zip(keys, (int(x) for x in stats)))


def pipe_plan(m, t: _task.Task, tile_cells=128, dir_bins=True, hilbert=True, nc_owned=None):
    """Host-only: statistics of the per-tile blobs of the pipelined tile kernel (layout 2); every table
    is re-derived from the blob bytes inside the hook (CPU test hook)."""
    lib = load_library()
    pk = Packed(m, t, nc_owned=nc_owned)
    stats = np.zeros(12, np.int64)
    rc = lib.cfd2d_pipe_plan(C.byref(pk.mesh), int(tile_cells), 1 if dir_bins else 0, 1 if hilbert else 0,
                             stats.ctypes.data_as(C.POINTER(C.c_int64)))
    if rc != 0:
        msg = lib.cfd2d_fvm_last_error(None)
        raise CFDError(rc, msg.decode() if msg else "")
    keys = ("ntiles", "nl2_max", "ne_max", "blob_max", "sum_ne", "sum_ring1", "sum_ring2", "blob_bytes", "interior", "boundary",
            "nring_max", "nl_max")
    return dict(zip(keys, (int(x) for x in stats)))


def kat_rim_orig(in8, gam=1.4, max_newton=0, device=0, fast=False):
    """rim_orig on the device: fast=False the statement that keeps the reference's operation order,
    fast=True the reduced-instruction solver the Godunov kernels use by default (g = 1.4 only)."""
    lib = load_library()
    a = _f64(in8)
    out = np.empty((a.shape[0], 5))
    it = np.empty(a.shape[0], np.int32)
    if fast:
        rc = lib.cfd2d_kat_rim_orig_fast(device, a.shape[0], a.ctypes.data_as(_dp), int(max_newton),
                                         out.ctypes.data_as(_dp), it.ctypes.data_as(_ip))
    else:
        rc = lib.cfd2d_kat_rim_orig(device, a.shape[0], a.ctypes.data_as(_dp), float(gam), int(max_newton),
                                    out.ctypes.data_as(_dp), it.ctypes.data_as(_ip))
    if rc not in (0, -4):
        raise CFDError(rc, "kat_rim_orig")
    return out, it


def kat_calc_flux(in12, gam=1.4, flux=FLUX_GODUNOV, device=0):
    lib = load_library()
    a = _f64(in12)
    out = np.empty((a.shape[0], 4))
    rc = lib.cfd2d_kat_calc_flux(device, a.shape[0], a.ctypes.data_as(_dp), float(gam), int(flux), out.ctypes.data_as(_dp))
    if rc != 0:
        raise CFDError(rc, "kat_calc_flux")
    return out


def kat_urs(io8, M, Cp, mode, device=0):
    """Material::URS (global.cpp:9-30) on the device; io8[n][8] = r,p,e,E,u,v,cz,T (a copy is returned)."""
    lib = load_library()
    a = np.array(io8, dtype=np.float64, copy=True, order="C")
    rc = lib.cfd2d_kat_urs(device, a.shape[0], float(M), float(Cp), int(mode), a.ctypes.data_as(_dp))
    if rc != 0:
        raise CFDError(rc, "kat_urs")
    return a


# ------------------------------------------------------------------------------------------------
# Host mirror of the reference's method object
# ------------------------------------------------------------------------------------------------

class FVM_TVD:
    """Mirror of ``class FVM_TVD : public Method`` (reference src/methods/fvm_tvd.h:7-113): same
    life cycle -- ``init(xmlFileName)``, ``run()``, ``done()`` -- same task.xml, same UNV mesh, same
    ``res_%010d.vtk`` output and log lines; the loop body runs on the GPU through the C-ABI.

    Extra, optional task.xml element the reference ignores: ``<gpu device=".." flux="GODUNOV|LAX"
    order="1|2"/>``.
    """

    def __init__(self, device=0, flux=FLUX_GODUNOV, order=2, workdir="."):
        self.device, self.flux, self.order, self.workdir = device, flux, order, workdir
        self.solver = None
        self.log_lines = []

    def _log(self, s):
        self.log_lines.append(s)
        print(s, end="")

    def init(self, xml_file_name: str):
        from . import unv as _unv
        path = os.path.join(self.workdir, xml_file_name)
        self.task = t = _task.read_task_xml(path)
        if t.mesh_type != "salome_unv":
            # the Berkeley-Triangle reader yields zero Gauss points per edge, i.e. no
            # flux at all; this path supports the live format only.
            raise CFDError(-1, f"mesh filesType '{t.mesh_type}' not supported on the GPU path (use salome_unv)")
        nodes, tris, cell_groups, edge_groups = _unv.read_unv(os.path.join(self.workdir, t.mesh_name))
        self.nodes, self.tris = nodes, tris
        self.mesh = m = _mesh.build_mesh(nodes, tris)
        from .cases import bind
        bind(m, t, cell_groups, edge_groups)
        self.solver = Solver(m, t, self.flux, self.order, device=self.device)
        nc = m.nc
        ro = np.empty(nc); ru = np.empty(nc); rv = np.empty(nc); re = np.empty(nc)
        for ir, r in enumerate(t.regions):
            s = _task.region_state(t, r)
            sel = m.cell_region == ir
            ro[sel], ru[sel], rv[sel], re[sel] = s
        self.solver.set_state(ro, ru, rv, re)
        tau = self.solver.calc_time_step()
        if not t.steady:
            self._log("time step: %25.16E\n" % tau)
        self.save(0)

    def run(self):
        t = self.task
        tt, step = 0.0, 0
        while tt < t.TMAX and step < t.STEP_MAX:
            nxt = min(t.STEP_MAX,
                      (step // t.FILE_OUTPUT_STEP + 1) * t.FILE_OUTPUT_STEP,
                      (step // t.LOG_OUTPUT_STEP + 1) * t.LOG_OUTPUT_STEP)
            n = nxt - step
            if not t.steady:
                tau = self.solver.tau
                if tau > 0:
                    import math
                    n = max(1, min(n, int(math.ceil((t.TMAX - tt) / tau))))
            self.solver.step(n)
            step += n
            tt = self.solver.time
            if step % t.FILE_OUTPUT_STEP == 0:
                self.save(step)
            if step % t.LOG_OUTPUT_STEP == 0:
                self._log("step: %d\t\ttime step: %.16f\n" % (step, tt))

    def save(self, step: int):
        from . import vtk as _vtk
        prim = self.solver.get_primitive()
        _, _, _, _, ctau, _ = self.solver.get_state(want_flag=False)
        gam = np.array([mm.gamma for mm in self.task.materials])[self.mesh.cell_mat]
        _vtk.write_vtk(os.path.join(self.workdir, "res_%010d.vtk" % step), self.nodes, self.tris, prim, ctau, gam)
        print("File 'res_%010d.vtk' saved..." % step)

    def done(self):
        if self.solver:
            self.solver.close()
            self.solver = None
