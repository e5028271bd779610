"""ctypes window on oracle/_build/libfvm_oracle.so (oracle/fvm_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, oracle/make_golden.py, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg may import this.  It reuses the product's ctypes struct packing so checker and
checked consume identical bytes.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libfvm_oracle.so")
_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def build(force=False):
    src = os.path.join(HERE, "fvm_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    return LIB


def load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    from cfd2d_b200.fvm import CMesh, CPhys, CCtrl
    lib = C.CDLL(LIB, mode=os.RTLD_LOCAL)
    H = C.c_void_p
    lib.fvm_oracle_create.restype = H
    lib.fvm_oracle_create.argtypes = [C.POINTER(CMesh), C.POINTER(CPhys), C.POINTER(CCtrl)]
    lib.fvm_oracle_destroy.argtypes = [H]
    lib.fvm_oracle_set_state.argtypes = [H, _dp, _dp, _dp, _dp, C.c_void_p]
    lib.fvm_oracle_calc_time_step.argtypes = [H]
    lib.fvm_oracle_calc_time_step.restype = C.c_double
    lib.fvm_oracle_step.argtypes = [H, C.c_int]
    lib.fvm_oracle_get_state.argtypes = [H, _dp, _dp, _dp, _dp, C.c_void_p, C.c_void_p]
    lib.fvm_oracle_calc_grad.argtypes = [H, _dp]
    lib.fvm_oracle_edge_fluxes.argtypes = [H, _dp]
    lib.fvm_oracle_get_primitive.argtypes = [H] + [C.c_void_p] * 6
    lib.fvm_oracle_newton_iters.argtypes = [H]
    lib.fvm_oracle_newton_iters.restype = C.c_longlong
    lib.fvm_oracle_riemann_calls.argtypes = [H]
    lib.fvm_oracle_riemann_calls.restype = C.c_longlong
    lib.fvm_oracle_rim_orig.argtypes = [C.c_int, _dp, C.c_double, C.c_int, _dp, _ip]
    lib.fvm_oracle_calc_flux.argtypes = [C.c_int, _dp, C.c_double, C.c_int, _dp]
    lib.fvm_oracle_urs.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, _dp]
    _lib = lib
    return lib


class OracleSolver:
    def __init__(self, mesh, task, flux=0, order=2, max_newton=0):
        from cfd2d_b200.fvm import Packed
        self.lib = load()
        self.pk = Packed(mesh, task, flux, order, max_newton)
        self.nc, self.ne = self.pk.nc, self.pk.ne
        self.h = self.lib.fvm_oracle_create(C.byref(self.pk.mesh), C.byref(self.pk.phys), C.byref(self.pk.ctrl))

    def set_state(self, ro, ru, rv, re, flag=None):
        a = [np.ascontiguousarray(x, np.float64) for x in (ro, ru, rv, re)]
        fl = None if flag is None else np.ascontiguousarray(flag, np.uint32)
        self.lib.fvm_oracle_set_state(self.h, *[x.ctypes.data_as(_dp) for x in a],
                                      None if fl is None else C.c_void_p(fl.ctypes.data))

    def calc_time_step(self):
        return self.lib.fvm_oracle_calc_time_step(self.h)

    def step(self, n=1):
        return self.lib.fvm_oracle_step(self.h, int(n))

    def get_state(self):
        n = self.nc
        ro, ru, rv, re, ct = (np.empty(n) for _ in range(5))
        fl = np.empty(n, np.uint32)
        self.lib.fvm_oracle_get_state(self.h, *[x.ctypes.data_as(_dp) for x in (ro, ru, rv, re)],
                                      C.c_void_p(ct.ctypes.data), C.c_void_p(fl.ctypes.data))
        return ro, ru, rv, re, ct, fl

    def calc_grad(self):
        g = np.empty((self.nc, 8))
        self.lib.fvm_oracle_calc_grad(self.h, g.ctypes.data_as(_dp))
        return g

    def edge_fluxes(self):
        f = np.empty((self.ne, 4))
        self.lib.fvm_oracle_edge_fluxes(self.h, f.ctypes.data_as(_dp))
        return f

    def get_primitive(self):
        arrs = [np.empty(self.nc) for _ in range(6)]
        self.lib.fvm_oracle_get_primitive(self.h, *[C.c_void_p(x.ctypes.data) for x in arrs])
        return dict(zip(("r", "p", "T", "u", "v", "cz"), arrs))

    @property
    def newton_iters(self):
        return self.lib.fvm_oracle_newton_iters(self.h)

    @property
    def riemann_calls(self):
        return self.lib.fvm_oracle_riemann_calls(self.h)

    def close(self):
        if self.h:
            self.lib.fvm_oracle_destroy(self.h)
            self.h = None


def rim_orig(in8, gam=1.4, max_newton=0):
    a = np.ascontiguousarray(in8, np.float64)
    out = np.empty((a.shape[0], 5))
    it = np.empty(a.shape[0], np.int32)
    load().fvm_oracle_rim_orig(a.shape[0], a.ctypes.data_as(_dp), float(gam), int(max_newton),
                               out.ctypes.data_as(_dp), it.ctypes.data_as(_ip))
    return out, it


def calc_flux(in12, gam=1.4, flux=0):
    a = np.ascontiguousarray(in12, np.float64)
    out = np.empty((a.shape[0], 4))
    load().fvm_oracle_calc_flux(a.shape[0], a.ctypes.data_as(_dp), float(gam), int(flux), out.ctypes.data_as(_dp))
    return out


def urs(io8, M, Cp, mode):
    a = np.array(io8, dtype=np.float64, copy=True, order="C")
    load().fvm_oracle_urs(a.shape[0], float(M), float(Cp), int(mode), a.ctypes.data_as(_dp))
    return a
