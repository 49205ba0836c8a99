"""Mirror of the reference's plugin loader `SimPlugin` (sim/cases.py:13-66) bound to this library's
`c_plbm_*` exports -- the sim/lbm.h seam.  Same method names and argument meaning:

    sim = SimPlugin()                 # reference: SimPlugin("./libslbm.so")
    sim.init((nx, ny), dt, p, u)      # p is PRESSURE, u has shape (2, ny, nx): Fortran u(nx,ny,2)
    sim.step(omega)                   # collide -> push-stream -> periodic fold
    rho, u = sim.vars()
    sim.free()
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .capi import LIB_PATH, PlbmError, lib


class SimPlugin:
    def __init__(self, path: str = LIB_PATH, name: str = "plbm"):
        # the reference derives the symbol prefix from the file name: lib<name>.so -> c_<name>_*
        self.name = name
        self._init = getattr(lib, f"c_{name}_init")
        self._step = getattr(lib, f"c_{name}_step")
        self._step_n = getattr(lib, f"c_{name}_step_n", None)
        self._vars = getattr(lib, f"c_{name}_vars")
        self._free = getattr(lib, f"c_{name}_free")
        self._norm = getattr(lib, f"c_{name}_norm")
        self.ptr = None
        self.shape = None

    def init(self, grid_size, dt, rho, u, sigma=None, params=None):
        nx, ny = grid_size
        rho = np.ascontiguousarray(rho, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        assert rho.size == nx * ny and u.size == 2 * nx * ny
        sp = None if sigma is None else np.ascontiguousarray(sigma, dtype=np.float64).ctypes.data_as(C.c_void_p)
        self.ptr = self._init(nx, ny, float(dt), rho.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p), sp, None)
        if not self.ptr:
            msg = lib.plbm_last_error()
            raise PlbmError(msg.decode() if msg else "c_plbm_init failed")
        self.shape = (nx, ny)

    def step(self, omega, n=1):
        if n == 1 or self._step_n is None:
            for _ in range(int(n)):
                self._step(self.ptr, float(omega))
        else:
            self._step_n(self.ptr, float(omega), int(n))

    def vars(self):
        nx, ny = self.shape
        rho = np.empty((ny, nx), dtype=np.float64)
        u = np.empty((2, ny, nx), dtype=np.float64)
        self._vars(self.ptr, rho.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p))
        return rho, u

    def norm(self, u, ua):
        u = np.ascontiguousarray(u, dtype=np.float64)
        ua = np.ascontiguousarray(ua, dtype=np.float64)
        ny, nx = u.shape[-2:]
        return float(self._norm(nx, ny * (u.size // (nx * ny)), u.ctypes.data_as(C.c_void_p), ua.ctypes.data_as(C.c_void_p)))

    def free(self):
        if self.ptr:
            self._free(self.ptr)
            self.ptr = None
