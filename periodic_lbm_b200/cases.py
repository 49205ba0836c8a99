"""Flow cases of the reference (`src/benchmarks/`), host side, and the driver recipes built on them.

  TaylorGreen   <-> type taylor_green_t   (src/benchmarks/taylor_green.f90:12-84)
  VortexCase    <-> type vortex_case_t    (src/benchmarks/barotropic_vortex_case.F90:17-83)

Evaluation happens in libplbm_b200.so's host helpers (same libm and expression order as the
Fortran intrinsics); parameter derivations follow the drivers (app/main_taylor_green.f90:44-89,
app/main_vortex.f90:48-97) in the working precision.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import check, lib


def _prec(dtype):
    return capi.F64 if np.dtype(dtype) == np.float64 else capi.F32


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def pi(dtype=np.float64):
    """pi = 4*atan(1._wp) (src/benchmarks/taylor_green.f90:23)"""
    T = np.dtype(dtype).type
    return T(4) * np.arctan(T(1))


class TaylorGreen:
    def __init__(self, nx, ny, kx, ky, umax, nu, dtype=np.float64):
        T = np.dtype(dtype).type
        self.dtype = T
        self.nx, self.ny = nx, ny
        self.kx, self.ky, self.umax, self.nu = T(kx), T(ky), T(umax), T(nu)
        self.td = T(lib.plbm_case_tg_decay_time(_prec(T), float(self.kx), float(self.ky), float(self.nu)))

    def decay_time(self):
        return self.td

    def eval(self, t, x_offset=0, nx_local=None, out=None):
        """-> p, ux, uy as (nx, ny) arrays (Fortran (ny,nx)).  x_offset/nx_local select the lines of
        one slab of the global grid; `out` may supply three preallocated (e.g. pinned) arrays."""
        nxl = self.nx if nx_local is None else nx_local
        p, ux, uy = out if out is not None else (np.empty((nxl, self.ny), dtype=self.dtype) for _ in range(3))
        check(lib.plbm_case_taylor_green_slab(_prec(self.dtype), nxl, int(x_offset), self.ny, float(self.kx), float(self.ky),
                                              float(self.umax), float(self.td), float(self.dtype(t)), _ptr(p), _ptr(ux), _ptr(uy)),
              "case_taylor_green")
        return p, ux, uy


class VortexCase:
    def __init__(self, U0, xc, yc, Rc, eps, rho0=1.0, csqr=None, dtype=np.float64):
        T = np.dtype(dtype).type
        self.dtype = T
        self.U0, self.xc, self.yc, self.Rc, self.eps, self.rho0 = T(U0), T(xc), T(yc), T(Rc), T(eps), T(rho0)
        self.csqr = T(1) / T(3) if csqr is None else T(csqr)

    def eval(self, nx, ny):
        rho, ux, uy = (np.empty((nx, ny), dtype=self.dtype) for _ in range(3))
        check(lib.plbm_case_vortex(_prec(self.dtype), nx, ny, float(self.U0), float(self.xc), float(self.yc), float(self.Rc),
                                   float(self.eps), float(self.rho0), float(self.csqr), _ptr(rho), _ptr(ux), _ptr(uy)),
              "case_vortex")
        return rho, ux, uy


def taylor_green_params(n, dt=None, dt_over_tau=None, dtype=np.float64):
    """app/main_taylor_green.f90:44-89: umax = 0.01/sqrt(3), nu = umax*n/100, tau = 3 nu,
    k = 2 pi / n, tmax = ln2 * td, nsteps = int(1.1 tmax/dt).  dt is the CLI argument; the golden
    sweeps of graphs/fvm_*_64.txt use dt = (dt/tau)*tau instead (the commented variant, :60-62)."""
    T = np.dtype(dtype).type
    umax = T(0.01) / np.sqrt(T(3))
    nu = (umax * T(n)) / T(100)
    tau = T(3) * nu
    dt = T(dt) if dt is not None else T(dt_over_tau) * tau
    k = T(2) * pi(T) / T(n)
    tg = TaylorGreen(n, n, k, k, umax, nu, dtype=T)
    tmax = np.log(T(2)) * tg.decay_time()
    return dict(umax=umax, nu=nu, tau=tau, dt=dt, kx=k, ky=k, case=tg, tmax=tmax, nsteps=int(T(1.1) * tmax / dt), magic=T(1) / T(4))


def steps_until(tmax, dt, nsteps_cap, dtype=np.float64):
    """Stopping rule of the drivers (app/main_taylor_green.f90:98-119): t is accumulated by repeated
    t = t + dt in working precision; stop at the first t >= tmax.  -> (steps, t)."""
    T = np.dtype(dtype).type
    t, step, remaining = T(0), 0, int(nsteps_cap)
    while remaining > 0:
        k = min(remaining, 1 << 16)
        acc = np.cumsum(np.concatenate(([t], np.full(k, dt, dtype=T))), dtype=T)[1:]  # sequential recurrence
        hit = np.nonzero(acc >= tmax)[0]
        if hit.size:
            return step + int(hit[0]) + 1, acc[hit[0]]
        step, remaining, t = step + k, remaining - k, acc[-1]
    return step, t


def vortex_params(n, dtype=np.float64):
    """app/main_vortex.f90:48-97: U0 = 0.1/sqrt(3), kappa = 0.2/sqrt(3), nu = 1e-5, xc=yc=n/2, Rc=n/10."""
    T = np.dtype(dtype).type
    U0 = T(0.1) / np.sqrt(T(3))
    kappa = T(0.2) / np.sqrt(T(3))
    nu = T(0.00001)
    case = VortexCase(U0, T(n) / T(2), T(n) / T(2), T(n) / T(10), kappa, dtype=T)
    tmax = (T(n) / U0) * T(4)
    return dict(U0=U0, kappa=kappa, nu=nu, case=case, tmax=tmax, magic=T(1) / T(4))
