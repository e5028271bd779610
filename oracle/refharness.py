"""ctypes window on oracle/_ref/libcfd2d_ref_{v0,v1,v2}.so -- TEST INFRASTRUCTURE ONLY.

The .so files are the REAL reference (zhrv/cfd-2d FVM_TVD) compiled by oracle/Makefile from
/root/reference plus oracle/ref_harness.cpp.  Only tests/, oracle/make_golden.py,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS: dict[str, C.CDLL] = {}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_up = C.POINTER(C.c_uint)


def lib_path(variant: str = "v0") -> str:
    return os.path.join(HERE, "_ref", f"libcfd2d_ref_{variant}.so")


def available(variant: str = "v0") -> bool:
    return os.path.exists(lib_path(variant))


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def load(variant: str = "v0") -> C.CDLL:
    if variant in _LIBS:
        return _LIBS[variant]
    lib = C.CDLL(lib_path(variant), mode=os.RTLD_LOCAL)
    lib.ref_open.restype = C.c_void_p
    lib.ref_open.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    lib.ref_close.argtypes = [C.c_void_p]
    lib.ref_counts.argtypes = [C.c_void_p, _ip, _ip, _ip, _ip, _ip]
    lib.ref_get_mesh.argtypes = [C.c_void_p] + [C.c_void_p] * 17
    lib.ref_get_phys.argtypes = [C.c_void_p, _dp, _dp, _ip, _dp, _dp, _dp, _dp, _ip]
    lib.ref_set_state.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
    lib.ref_set_flags.argtypes = [C.c_void_p, _up]
    lib.ref_set_control.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int]
    lib.ref_set_limits.argtypes = [C.c_void_p, _dp]
    lib.ref_calc_time_step.restype = C.c_double
    lib.ref_calc_time_step.argtypes = [C.c_void_p, C.c_int]
    lib.ref_run.restype = C.c_double
    lib.ref_run.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.ref_get_state.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp, _up]
    lib.ref_calc_grad.argtypes = [C.c_void_p, _dp]
    lib.ref_edge_fluxes.argtypes = [C.c_void_p, _dp]
    lib.ref_cons_to_par.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.ref_boundary_cond.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
    lib.ref_rim_orig.argtypes = [C.c_int, _dp, C.c_double, _dp]
    lib.ref_calc_flux.argtypes = [C.c_int, _dp, C.c_double, _dp]
    lib.ref_urs.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, _dp]
    lib.ref_variant.restype = C.c_char_p
    _LIBS[variant] = lib
    return lib


class RefSolver:
    """One reference FVM_TVD instance, initialised from <workdir>/<xml> (+ the UNV it names)."""

    def __init__(self, workdir: str, xml: str = "task.xml", variant: str = "v0", quiet: bool = True):
        self.lib = load(variant)
        self.variant = variant
        self.quiet = 1 if quiet else 0
        cwd = os.getcwd()
        try:
            self.h = self.lib.ref_open(os.path.abspath(workdir).encode(), xml.encode(), self.quiet)
        finally:
            os.chdir(cwd)
        if not self.h:
            raise RuntimeError("ref_open failed")
        c = [C.c_int() for _ in range(5)]
        self.lib.ref_counts(self.h, *[C.byref(x) for x in c])
        self.nc, self.ne, self.nn, self.nmat, self.nbc = [x.value for x in c]

    def mesh(self) -> dict:
        nc, ne, nn = self.nc, self.ne, self.nn
        m = dict(
            nodes=np.empty((nn, 2)), cell_nodes=np.empty((nc, 3), np.int32),
            cell_edges=np.empty((nc, 3), np.int32), cell_neigh=np.empty((nc, 3), np.int32),
            cell_S=np.empty(nc), cell_cx=np.empty(nc), cell_cy=np.empty(nc), cell_mat=np.empty(nc, np.int32),
            edge_n1=np.empty(ne, np.int32), edge_n2=np.empty(ne, np.int32),
            edge_c1=np.empty(ne, np.int32), edge_c2=np.empty(ne, np.int32),
            edge_nx=np.empty(ne), edge_ny=np.empty(ne), edge_l=np.empty(ne),
            edge_gp=np.empty((ne, 4)), edge_bc=np.empty(ne, np.int32))
        self.lib.ref_get_mesh(self.h, *[v.ctypes.data_as(C.c_void_p) for v in m.values()])
        return m

    def phys(self) -> dict:
        mat_M = np.empty(self.nmat); mat_Cp = np.empty(self.nmat)
        bc_kind = np.empty(max(self.nbc, 1), np.int32); bc_par = np.empty((max(self.nbc, 1), 4))
        lim = np.empty(5)
        cfl = C.c_double(); tau = C.c_double(); steady = C.c_int()
        self.lib.ref_get_phys(self.h, _d(mat_M), _d(mat_Cp), _i(bc_kind), _d(bc_par), _d(lim),
                              C.byref(cfl), C.byref(tau), C.byref(steady))
        return dict(mat_M=mat_M, mat_Cp=mat_Cp, bc_kind=bc_kind[:self.nbc], bc_par=bc_par[:self.nbc],
                    limits=lim, CFL=cfl.value, TAU=tau.value, steady=steady.value)

    def set_state(self, ro, ru, rv, re):
        a = [np.ascontiguousarray(x, np.float64) for x in (ro, ru, rv, re)]
        self.lib.ref_set_state(self.h, *[_d(x) for x in a])

    def set_flags(self, flag):
        f = np.ascontiguousarray(flag, np.uint32)
        self.lib.ref_set_flags(self.h, f.ctypes.data_as(_up))

    def set_control(self, tau, cfl, steady):
        self.lib.ref_set_control(self.h, float(tau), float(cfl), int(steady))

    def set_limits(self, limits5):
        l = np.ascontiguousarray(limits5, np.float64)
        self.lib.ref_set_limits(self.h, _d(l))

    def calc_time_step(self) -> float:
        return self.lib.ref_calc_time_step(self.h, self.quiet)

    def run(self, nsteps: int) -> float:
        return self.lib.ref_run(self.h, int(nsteps), self.quiet)

    def state(self):
        n = self.nc
        ro, ru, rv, re, ct = (np.empty(n) for _ in range(5))
        fl = np.empty(n, np.uint32)
        self.lib.ref_get_state(self.h, _d(ro), _d(ru), _d(rv), _d(re), _d(ct), fl.ctypes.data_as(_up))
        return ro, ru, rv, re, ct, fl

    def calc_grad(self):
        g = np.empty((self.nc, 8))
        self.lib.ref_calc_grad(self.h, _d(g))
        return g

    def edge_fluxes(self):
        f = np.empty((self.ne, 4))
        self.lib.ref_edge_fluxes(self.h, _d(f))
        return f

    def boundary_cond(self, iedge: int, pL8):
        a = np.ascontiguousarray(pL8, np.float64)
        out = np.empty(8)
        self.lib.ref_boundary_cond(self.h, int(iedge), _d(a), _d(out))
        return out

    def cons_to_par(self, icell: int):
        out = np.empty(8)
        self.lib.ref_cons_to_par(self.h, int(icell), _d(out))
        return out

    def close(self):
        if self.h:
            self.lib.ref_close(self.h)
            self.h = None


def rim_orig(in8, gam=1.4, variant="v0"):
    a = np.ascontiguousarray(in8, np.float64)
    out = np.empty((a.shape[0], 5))
    load(variant).ref_rim_orig(a.shape[0], _d(a), float(gam), _d(out))
    return out


def calc_flux(in12, gam=1.4, variant="v0"):
    a = np.ascontiguousarray(in12, np.float64)
    out = np.empty((a.shape[0], 4))
    load(variant).ref_calc_flux(a.shape[0], _d(a), float(gam), _d(out))
    return out


def urs(io8, M, Cp, mode, variant="v0"):
    a = np.array(io8, dtype=np.float64, copy=True, order="C")
    load(variant).ref_urs(a.shape[0], float(M), float(Cp), int(mode), _d(a))
    return a
