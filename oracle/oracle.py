"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package (periodic_lbm_b200) never does.

The oracle is the plain-C restatement in oracle/plbm_oracle.c (reference file:line
citations live there).  `Oracle(precision)` exposes the kernels on numpy arrays in the
reference layout: PDFs f[q, x, y_padded] (C order == Fortran f(ld,nx,0:8)), macroscopic
fields [x, y] (C order == Fortran (ny,nx)).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(force: bool = False) -> None:
    """Compile liboracle.so / liboracle_omp.so with the committed Makefile."""
    need = force or not all(
        os.path.exists(os.path.join(_HERE, n)) for n in ("liboracle.so", "liboracle_omp.so")
    )
    if not need:
        src_m = max(os.path.getmtime(os.path.join(_HERE, n)) for n in ("plbm_oracle.c", "plbm_oracle_impl.h"))
        so_m = min(os.path.getmtime(os.path.join(_HERE, n)) for n in ("liboracle.so", "liboracle_omp.so"))
        need = src_m > so_m
    if need:
        subprocess.run(["make", "-C", _HERE, "-B", "all"], check=True, capture_output=True)


def padded_ld(ny: int) -> int:
    """Leading dimension of f: ny rounded up to a multiple of 16 (src/fvm_bardow.F90:144-147)."""
    return (ny + 15) // 16 * 16


class Oracle:
    SCHEME_LBM, SCHEME_FVM_BARDOW, SCHEME_DUGKS, SCHEME_DUGKS_OFF, SCHEME_FDM_BARDOW, SCHEME_FDM_SOFONEA = 0, 1, 2, 3, 4, 5
    BGK, TRT, RR, BGK_SPLIT, TRT_SPLIT, BGK_IMPROVED = 0, 1, 2, 3, 4, 5

    def __init__(self, precision: str = "f64", omp: bool = False):
        build()
        assert precision in ("f64", "f32")
        self.sfx = "_" + precision
        self.dtype = np.float64 if precision == "f64" else np.float32
        self.ctype = C.c_double if precision == "f64" else C.c_float
        self.lib = C.CDLL(os.path.join(_HERE, "liboracle_omp.so" if omp else "liboracle.so"))
        self._sig()

    # -- plumbing --------------------------------------------------------
    def _fn(self, name, restype, argtypes):
        fn = getattr(self.lib, name + self.sfx)
        fn.restype = restype
        fn.argtypes = argtypes
        return fn

    def _sig(self):
        R, I, P = self.ctype, C.c_int, C.c_void_p
        self._equilibrium = self._fn("orc_equilibrium", None, [R, R, R, P])
        self._set_properties = self._fn("orc_set_properties", None, [R, R, R, I, P])
        self._set_pdf = self._fn("orc_set_pdf_to_equilibrium", None, [I, I, I, P, P, P, P])
        self._macros = self._fn("orc_update_macros", None, [I, I, I, P, P, P, P])
        self._stream = self._fn("orc_lbm_stream", None, [I, I, I, P, P])
        self._bgk = self._fn("orc_collide_bgk", None, [I, I, I, P, R])
        self._kbgk = self._fn("orc_kernel_bgk", None, [I, I, I, P, R])
        self._trt = self._fn("orc_collide_trt", None, [I, I, I, P, R, R])
        self._rr = self._fn("orc_collide_rr", None, [I, I, I, P, R])
        self._trt_split = self._fn("orc_collide_trt_split", None, [I, I, I, P, R, R])
        self._bgk_improved = self._fn("orc_collide_bgk_improved", None, [I, I, I, P, R])
        self._lambda_d = self._fn("orc_lambda_d", R, [R, R])
        self._magic = self._fn("orc_magic_number", R, [R, R])
        self._fvm = self._fn("orc_stream_fvm_bardow", None, [I, I, I, P, P, R])
        self._fdm_bardow = self._fn("orc_stream_fdm_bardow", None, [I, I, I, P, P, R])
        self._fdm_bardow_stencil = self._fn("orc_stream_fdm_bardow_stencil", None, [I, I, I, P, P, R, I])
        self._fdm_sofonea = self._fn("orc_stream_fdm_sofonea", None, [I, I, I, P, P, R])
        self._dcollide = self._fn("orc_dugks_collide", None, [I, I, I, P, P, R, R, R, I])
        self._dstream = self._fn("orc_dugks_stream", None, [I, I, I, P, P, R, R, I])
        self._v2 = self._fn("orc_vorticity_2nd", None, [I, I, P, P, P])
        self._v4 = self._fn("orc_vorticity_4th", None, [I, I, P, P, P])
        self._tgtd = self._fn("orc_tg_decay_time", R, [R, R, R])
        self._tg = self._fn("orc_taylor_green_eval", None, [I, I, R, R, R, R, R, P, P, P])
        self._vortex = self._fn("orc_vortex_eval", None, [I, I, R, R, R, R, R, R, R, P, P, P])
        self._l2 = self._fn("orc_l2_norm", R, [I, I, P, P, P, P, P])
        self._run = self._fn("orc_run", None, [I, I, I, P, P, P, I, I, R, R, R, R, C.c_long])
        self._sim_eqinit = self._fn("orc_sim_eqinit", None, [I, I, P, P, P, P])
        self._sim_step = self._fn("orc_sim_collide_and_stream", None, [I, I, P, P, R])
        self._sim_bc = self._fn("orc_sim_periodic_bc_push", None, [I, I, P])
        self._sim_macros = self._fn("orc_sim_macros", None, [I, I, P, P, P, P])
        self._lw_stream = self._fn("orc_lw_stream", None, [I, I, P, P, R])
        self._lw_collision = self._fn("orc_lw_collision", None, [I, I, P, R])
        self._lw_bc = self._fn("orc_lw_bc", None, [I, I, P])
        # lw4 / lw6 (halo width H = order / 2)
        self._simh_eqinit = self._fn("orc_simh_eqinit", None, [I, I, I, P, P, P, P])
        self._simh_macros = self._fn("orc_simh_macros", None, [I, I, I, P, P, P, P])
        self._lwh_stream = self._fn("orc_lwh_stream", None, [I, I, I, P, P, R])
        self._lwh_collision = self._fn("orc_lwh_collision", None, [I, I, I, P, R])
        self._lwh_bc = self._fn("orc_lwh_bc", None, [I, I, I, P])
        # Heun finite-volume plugin (sim/sim_fvm.F90); the caller swaps f1 <-> fc after a step
        self._simfvm_step = self._fn("orc_simfvm_step", None, [I, I, P, P, P, R, R])
        self.lib.orc_num_threads.restype = C.c_int
        self.lib.orc_set_num_threads.argtypes = [C.c_int]

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.c_void_p)

    def _chk(self, a, shape=None):
        assert a.dtype == self.dtype and a.flags.c_contiguous, (a.dtype, a.flags)
        if shape is not None:
            assert tuple(a.shape) == tuple(shape), (a.shape, shape)
        return a

    def num_threads(self) -> int:
        return int(self.lib.orc_num_threads())

    def set_num_threads(self, n: int) -> None:
        self.lib.orc_set_num_threads(int(n))

    # -- allocation helpers ---------------------------------------------
    def alloc_f(self, nx, ny, fill=np.nan):
        """One lattice f[q,x,ld]; padding rows poisoned with `fill` (never read by any kernel)."""
        return np.full((9, nx, padded_ld(ny)), fill, dtype=self.dtype)

    def alloc_m(self, nx, ny):
        return np.zeros((nx, ny), dtype=self.dtype)

    # -- kernels ----------------------------------------------------------
    def equilibrium(self, rho, ux, uy):
        out = np.zeros(9, dtype=self.dtype)
        self._equilibrium(rho, ux, uy, self._p(out))
        return out

    def set_properties(self, nu, dt, magic=None):
        """-> dict(tau, omega, trt_magic, csqr) in working precision."""
        out = np.zeros(4, dtype=self.dtype)
        self._set_properties(nu, dt, 0.0 if magic is None else magic, 0 if magic is None else 1, self._p(out))
        return dict(tau=out[0], omega=out[1], trt_magic=out[2], csqr=out[3], nu=self.dtype(nu), dt=self.dtype(dt))

    def set_pdf_to_equilibrium(self, rho, ux, uy, f):
        nx, ny = rho.shape
        self._chk(f, (9, nx, padded_ld(ny)))
        self._set_pdf(nx, ny, f.shape[2], self._p(self._chk(rho)), self._p(self._chk(ux)), self._p(self._chk(uy)), self._p(f))

    def update_macros(self, f, ny):
        nx, ld = f.shape[1], f.shape[2]
        rho, ux, uy = self.alloc_m(nx, ny), self.alloc_m(nx, ny), self.alloc_m(nx, ny)
        self._macros(nx, ny, ld, self._p(self._chk(f)), self._p(rho), self._p(ux), self._p(uy))
        return rho, ux, uy

    def lbm_stream(self, fsrc, fdst, ny):
        self._stream(fsrc.shape[1], ny, fsrc.shape[2], self._p(self._chk(fsrc)), self._p(self._chk(fdst)))

    def collide_bgk(self, f, ny, omega):
        self._bgk(f.shape[1], ny, f.shape[2], self._p(self._chk(f)), omega)

    def kernel_bgk(self, f, ny, omega):
        self._kbgk(f.shape[1], ny, f.shape[2], self._p(self._chk(f)), omega)

    def lambda_d(self, omega, magic):
        return self.dtype(self._lambda_d(omega, magic))

    def magic_number(self, le, ld):
        return self.dtype(self._magic(le, ld))

    def collide_trt(self, f, ny, omega, magic):
        self._trt(f.shape[1], ny, f.shape[2], self._p(self._chk(f)), omega, self._lambda_d(omega, magic))

    def collide_trt_split(self, f, ny, omega, magic):
        self._trt_split(f.shape[1], ny, f.shape[2], self._p(self._chk(f)), omega, self._lambda_d(omega, magic))

    def collide_bgk_improved(self, f, ny, omega):
        self._bgk_improved(f.shape[1], ny, f.shape[2], self._p(self._chk(f)), omega)

    def collide_rr(self, f, ny, omega):
        self._rr(f.shape[1], ny, f.shape[2], self._p(self._chk(f)), omega)

    def stream_fvm_bardow(self, fold, fnew, ny, dt):
        self._fvm(fold.shape[1], ny, fold.shape[2], self._p(self._chk(fold)), self._p(self._chk(fnew)), dt)

    FDM_STENCILS = {"default": 0, "wls": 1, "wls_gauss_v1": 2, "wls_gauss_v2": 3, "iso": 4}

    def stream_fdm_bardow(self, fold, fnew, ny, dt, stencil=0):
        """stencil: 0 default build, 1 -DFDM_WLS, 2 -DFDM_WLS_GAUSS_V1, 3 -DFDM_WLS_GAUSS_V2, 4 -DFDM_ISO"""
        stencil = self.FDM_STENCILS.get(stencil, stencil)
        if stencil == 0:
            self._fdm_bardow(fold.shape[1], ny, fold.shape[2], self._p(self._chk(fold)), self._p(self._chk(fnew)), dt)
        else:
            self._fdm_bardow_stencil(fold.shape[1], ny, fold.shape[2], self._p(self._chk(fold)), self._p(self._chk(fnew)), dt, int(stencil))

    def stream_fdm_sofonea(self, fold, fnew, ny, dt):
        self._fdm_sofonea(fold.shape[1], ny, fold.shape[2], self._p(self._chk(fold)), self._p(self._chk(fnew)), dt)

    def dugks_collide(self, fold, fnew, ny, omega, tau, dt, dugks=True):
        self._dcollide(fold.shape[1], ny, fold.shape[2], self._p(self._chk(fold)), self._p(self._chk(fnew)), omega, tau, dt, int(dugks))

    def dugks_stream(self, fold, fnew, ny, tau, dt, dugks=True):
        self._dstream(fold.shape[1], ny, fold.shape[2], self._p(self._chk(fold)), self._p(self._chk(fnew)), tau, dt, int(dugks))

    def vorticity(self, ux, uy, order=2):
        nx, ny = ux.shape
        om = self.alloc_m(nx, ny)
        (self._v2 if order == 2 else self._v4)(nx, ny, self._p(self._chk(ux)), self._p(self._chk(uy)), self._p(om))
        return om

    def tg_decay_time(self, kx, ky, nu):
        return self.dtype(self._tgtd(kx, ky, nu))

    def taylor_green_eval(self, nx, ny, kx, ky, umax, td, t):
        p, ux, uy = self.alloc_m(nx, ny), self.alloc_m(nx, ny), self.alloc_m(nx, ny)
        self._tg(nx, ny, kx, ky, umax, td, t, self._p(p), self._p(ux), self._p(uy))
        return p, ux, uy

    def vortex_eval(self, nx, ny, U0, xc, yc, Rc, eps, rho0=1.0, csqr=None):
        csqr = self.dtype(1.0) / self.dtype(3.0) if csqr is None else csqr
        rho, ux, uy = self.alloc_m(nx, ny), self.alloc_m(nx, ny), self.alloc_m(nx, ny)
        self._vortex(nx, ny, U0, xc, yc, Rc, eps, rho0, csqr, self._p(rho), self._p(ux), self._p(uy))
        return rho, ux, uy

    def l2_norm(self, ux, uy, uxa, uya):
        nx, ny = ux.shape
        scratch = np.zeros(2 * nx * ny, dtype=self.dtype)
        return self.dtype(self._l2(nx, ny, self._p(self._chk(ux)), self._p(self._chk(uy)), self._p(self._chk(uxa)), self._p(self._chk(uya)), self._p(scratch)))


class OracleGrid:
    """Mirror of the reference `lattice_grid` (src/fvm_bardow.F90:37-69) driven by the oracle."""

    def __init__(self, nx, ny, precision="f64", omp=False):
        self.o = Oracle(precision, omp=omp)
        self.nx, self.ny, self.ld = nx, ny, padded_ld(ny)
        # f[0] is lattice #1, f[1] is lattice #2 (1-based indices as in the reference)
        self.f = [self.o.alloc_f(nx, ny), self.o.alloc_f(nx, ny)]
        self.idx = np.array([2, 1], dtype=np.int32)  # [iold, inew]: alloc_grid sets inew=1, iold=2
        self.rho, self.ux, self.uy = (self.o.alloc_m(nx, ny) for _ in range(3))
        self.props = None

    iold = property(lambda s: int(s.idx[0]))
    inew = property(lambda s: int(s.idx[1]))

    def lattice(self, which):  # which: 1-based
        return self.f[which - 1]

    def set_properties(self, nu, dt, magic=None):
        self.props = self.o.set_properties(nu, dt, magic)
        return self.props

    def set_pdf_to_equilibrium(self):
        self.o.set_pdf_to_equilibrium(self.rho, self.ux, self.uy, self.lattice(self.iold))

    def run(self, scheme, collision, nsteps):
        p = self.props
        self.o._run(self.nx, self.ny, self.ld, Oracle._p(self.f[0]), Oracle._p(self.f[1]), Oracle._p(self.idx),
                    scheme, collision, p["omega"], p["tau"], p["dt"], p["trt_magic"], int(nsteps))

    def update_macros(self, lagged=True):
        """lagged=True reads lattice `inew` exactly like the reference."""
        src = self.lattice(self.inew if lagged else self.iold)
        self.rho, self.ux, self.uy = self.o.update_macros(src, self.ny)
        return self.rho, self.ux, self.uy


def taylor_green_setup(o: Oracle, n: int, dt=None, dt_over_tau=None, magic=0.25):
    """Parameter derivation of app/main_taylor_green.f90:44-89 in working precision."""
    T = o.dtype
    umax = T(0.01) / np.sqrt(T(3.0))
    nu = (umax * T(n)) / T(100.0)
    tau = T(3.0) * nu
    if dt is None:
        dt = T(dt_over_tau) * tau
    dt = T(dt)
    pi = T(4.0) * np.arctan(T(1.0))
    kx = T(2) * pi / T(n)
    ky = T(2) * pi / T(n)
    td = o.tg_decay_time(kx, ky, nu)
    tmax = np.log(T(2.0)) * td
    nsteps_cap = int(T(1.1) * tmax / dt)
    return dict(umax=umax, nu=nu, tau=tau, dt=dt, kx=kx, ky=ky, td=td, tmax=tmax, nsteps_cap=nsteps_cap, magic=magic)


def taylor_green_steps_to_tmax(o: Oracle, s) -> tuple[int, float]:
    """Stopping rule of app/main_taylor_green.f90:98-119: t accumulated by repeated t+=dt."""
    T = o.dtype
    t, dt, tmax = T(0.0), T(s["dt"]), T(s["tmax"])
    step = 0
    # vectorised accumulation is not bit-equal to the scalar loop; do it in chunks with cumsum on
    # the exact scalar recurrence (cumsum over a constant IS the sequential recurrence in numpy).
    remaining = s["nsteps_cap"]
    while remaining > 0:
        k = min(remaining, 1 << 16)
        acc = np.cumsum(np.concatenate(([t], np.full(k, dt, dtype=T))), dtype=T)[1:]
        hit = np.nonzero(acc >= tmax)[0]
        if hit.size:
            return step + int(hit[0]) + 1, acc[hit[0]]
        step += k
        remaining -= k
        t = acc[-1]
    return step, t


def taylor_green_l2_run(n, scheme, collision, dt=None, dt_over_tau=None, precision="f64", lagged=True, omp=False):
    """Whole driver pipeline of app/main_taylor_green.f90 on an n x n grid; returns (L2, steps, t, grid)."""
    g = OracleGrid(n, n, precision, omp=omp)
    o = g.o
    s = taylor_green_setup(o, n, dt=dt, dt_over_tau=dt_over_tau)
    g.set_properties(s["nu"], s["dt"], magic=s["magic"])
    p, ux, uy = o.taylor_green_eval(n, n, s["kx"], s["ky"], s["umax"], s["td"], o.dtype(0.0))
    g.rho = p / g.props["csqr"] + o.dtype(1.0)  # app/main_taylor_green.f90:145
    g.ux, g.uy = ux, uy
    g.set_pdf_to_equilibrium()
    steps, t = taylor_green_steps_to_tmax(o, s)
    g.run(scheme, collision, steps)
    g.update_macros(lagged=lagged)
    _, uxa, uya = o.taylor_green_eval(n, n, s["kx"], s["ky"], s["umax"], s["td"], t)
    return float(o.l2_norm(g.ux, g.uy, uxa, uya)), steps, float(t), g
