"""Host-side mirror of the reference's Fortran module interfaces, over the C ABI.

Names, argument meaning and call sequences follow the reference modules so a driver written
against them reads the same here:

  fvm_bardow            lattice_grid, alloc_grid, dealloc_grid, set_properties,
                        set_pdf_to_equilibrium, perform_step, update_macros, stream_fvm_bardow
                        (src/fvm_bardow.F90:13-33)
  periodic_lbm          perform_lbm_step, lbm_stream           (src/periodic_lbm.f90:9-11)
  collision_bgk/trt/regularized   collide_bgk, collide_trt, collide_rr, lambda_d, magic_number
  periodic_dugks        perform_dugks_step, dugks_collide, dugks_stream
  vorticity             vorticity_2nd, vorticity_4th           (src/vorticity.f90)

Arrays: numpy, C order.  A macroscopic field of the Fortran shape (ny,nx) is a numpy array of
shape (nx, ny) (same memory); a PDF lattice f(ld,nx,0:8) is numpy (9, nx, ld).

The PDFs never leave the GPU; `grid.rho/ux/uy` are the host views the drivers read and write
(app/main_taylor_green.f90:139-147, 105-109).  All compute goes through libplbm_b200.so; there is
no Python or CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import check, lib

__all__ = [
    "LatticeGrid", "alloc_grid", "dealloc_grid", "set_properties", "set_pdf_to_equilibrium",
    "perform_step", "perform_lbm_step", "perform_triple_step", "perform_dugks_step", "update_macros",
    "lbm_stream", "stream_fvm_bardow", "stream_fdm_bardow", "stream_fdm_sofonea", "collide_bgk", "collide_trt", "collide_rr", "collide_bgk_split",
    "collide_trt_split", "collide_bgk_improved",
    "dugks_collide", "dugks_stream", "vorticity_2nd", "vorticity_4th", "lambda_d", "magic_number",
    "cx", "cy", "csqr",
]

# src/fvm_bardow.F90:87-88, 94
cx = np.array([0, 1, 0, -1, 0, 1, -1, -1, 1])
cy = np.array([0, 0, 1, 0, -1, 1, 1, -1, -1])
csqr = 1.0 / 3.0

_PREC = {"f64": capi.F64, "f32": capi.F32, np.float64: capi.F64, np.float32: capi.F32}


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class LatticeGrid:
    """Mirror of `type lattice_grid` (src/fvm_bardow.F90:37-69)."""

    def __init__(self):
        self._h = None
        self.nx = self.ny = self.ld = self.nf = 0
        self.dtype = np.float64
        self.rho = self.ux = self.uy = None  # host views (Fortran (ny,nx))
        self.collision = None                # procedure pointers
        self.streaming = None
        self.logger = None
        self.filename = None
        self.foldername = None
        self.dugks = True                    # -DDUGKS branch of periodic_dugks
        self.lagged_macros = True            # update_macros reads lattice `inew` like the reference (F3)

    # scalar components kept coherent with the device handle
    def _props(self):
        out = (C.c_double * 6)()
        check(lib.plbm_get_properties(self._h, out), "get_properties")
        return list(out)

    nu = property(lambda s: s.dtype(s._props()[0]))
    dt = property(lambda s: s.dtype(s._props()[1]))
    tau = property(lambda s: s.dtype(s._props()[2]))
    trt_magic = property(lambda s: s.dtype(s._props()[4]))
    csqr = property(lambda s: s.dtype(s._props()[5]))

    @property
    def omega(self):
        return self.dtype(self._props()[3])

    @omega.setter
    def omega(self, value):
        check(lib.plbm_set_omega(self._h, float(value)), "set_omega")

    def _indices(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        check(lib.plbm_get_indices(self._h, C.byref(a), C.byref(b), C.byref(c)), "get_indices")
        return a.value, b.value, c.value

    iold = property(lambda s: s._indices()[0])
    inew = property(lambda s: s._indices()[1])
    imid = property(lambda s: s._indices()[2])

    # raw PDF access (tests / checkpoints); `which` is 1-based like grid%iold
    def download_f(self, which):
        f = np.empty((9, self.nx, self.ld), dtype=self.dtype)
        check(lib.plbm_download_f(self._h, which, _ptr(f)), "download_f")
        return f

    def upload_f(self, which, f):
        f = np.ascontiguousarray(f, dtype=self.dtype)
        assert f.shape == (9, self.nx, self.ld), f.shape
        check(lib.plbm_upload_f(self._h, which, _ptr(f)), "upload_f")

    def lattice_hash(self, which):
        """64-bit checksum of lattice `which` computed on the device (plbm_lattice_hash)."""
        out = C.c_ulonglong()
        check(lib.plbm_lattice_hash(self._h, int(which), C.byref(out)), "lattice_hash")
        return int(out.value)

    def synchronize(self):
        check(lib.plbm_synchronize(self._h), "synchronize")

    def set_stream(self, cuda_stream_ptr):
        check(lib.plbm_set_stream(self._h, C.c_void_p(cuda_stream_ptr)), "set_stream")

    def set_variant(self, variant):
        check(lib.plbm_set_variant(self._h, int(variant)), "set_variant")

    def set_step_deferral(self, max_pending):
        """Deferred stepping: single perform_lbm_step calls are counted and run batched (two steps per pass over
        HBM) once `max_pending` have accumulated or anything else touches the grid; 0 = eager (default)."""
        check(lib.plbm_set_step_deferral(self._h, int(max_pending)), "set_step_deferral")

    FDM_STENCILS = {"default": 0, "wls": 1, "wls_gauss_v1": 2, "wls_gauss_v2": 3, "iso": 4}

    def pair_kernel(self):
        """Kernel a perform_lbm_step call of >= 3 steps uses on this grid: 'k_lbm' (one step per launch),
        'k_lbm2' or 'k_lbm2_bulk' (two steps per launch)."""
        k = lib.plbm_lbm_pair_kernel(self._h)
        if k < 0:
            check(k, "lbm_pair_kernel")
        return ("k_lbm", "k_lbm2", "k_lbm2_bulk")[k]

    def steps_per_pass(self, collision=None):
        """Time steps one pass over HBM advances in a many-step perform_lbm_step call: 3 (k_lbmn_bulk), 2 (k_lbm2 / k_lbm2_bulk)
        or 1 (k_lbm).  `collision`: a collide_* procedure (default: grid.collision)."""
        coll = collision if collision is not None else self.collision
        k = lib.plbm_lbm_steps_per_pass(self._h, _COLLISION_ID[coll] if callable(coll) else int(coll))
        if k < 0:
            check(k, "lbm_steps_per_pass")
        return k

    def triple_kernel(self, collision=None):
        """Kernel of the three-step launches of a collision (default: grid.collision) that read no halo lines: 'k_lbmn_bulk' or
        'k_lbm3_ws' (the default except for the two-relaxation-time collisions; PLBM_TRIPLE_WS = 0 / 1 forces one)."""
        coll = collision if collision is not None else self.collision
        k = lib.plbm_lbm_triple_kernel(self._h, _COLLISION_ID[coll] if callable(coll) else int(coll))
        if k < 0:
            check(k, "lbm_triple_kernel")
        return ("k_lbmn_bulk", "k_lbm3_ws")[k]

    def closing_triple(self, collision=None):
        """True if a many-step perform_lbm_step call on this grid closes with a fused triple that also stores state n-1 (a third,
        hidden lattice buffer is available), False if it closes with a single-step launch.  Bit-identical either way."""
        coll = collision if collision is not None else self.collision
        k = lib.plbm_lbm_closing_triple(self._h, _COLLISION_ID[coll] if callable(coll) else int(coll))
        if k < 0:
            check(k, "lbm_closing_triple")
        return bool(k)

    def set_fdm_stencil(self, stencil):
        """Derivative stencil of stream_fdm_bardow; the reference picks it at compile time with -DFDM_WLS,
        -DFDM_WLS_GAUSS_V1, -DFDM_WLS_GAUSS_V2 or -DFDM_ISO (src/fvm_bardow.F90:591-660)."""
        check(lib.plbm_set_fdm_stencil(self._h, int(self.FDM_STENCILS.get(stencil, stencil))), "set_fdm_stencil")

    def diagnostics(self):
        """max/min |u|, sum(rho), kinetic energy of the device-resident macroscopic fields."""
        out = (C.c_double * capi.DIAG_COUNT)()
        check(lib.plbm_diagnostics(self._h, out), "diagnostics")
        return dict(max_speed=out[0], min_speed=out[1], sum_rho=out[2], kinetic_energy=out[3])

    def l2_sums(self, uxa, uya):
        """(sum |u - ua|^2, sum |ua|^2) of the device ux,uy vs host analytic fields; under a slab decomposition the
        sums run over the GLOBAL grid (each rank passes its own lines of the analytic fields)."""
        uxa = np.ascontiguousarray(uxa, dtype=self.dtype)
        uya = np.ascontiguousarray(uya, dtype=self.dtype)
        out = (C.c_double * 2)()
        check(lib.plbm_l2_sums(self._h, _ptr(uxa), _ptr(uya), out), "l2_sums")
        return float(out[0]), float(out[1])

    def l2_error(self, uxa, uya):
        """calc_L2_norm (app/main_taylor_green.f90:174-212) of the device ux,uy vs host analytic fields."""
        uxa = np.ascontiguousarray(uxa, dtype=self.dtype)
        uya = np.ascontiguousarray(uya, dtype=self.dtype)
        out = (C.c_double * 2)()
        check(lib.plbm_l2_sums(self._h, _ptr(uxa), _ptr(uya), out), "l2_sums")
        return float(np.sqrt(out[0] / out[1]))

    def __del__(self):
        try:
            dealloc_grid(self)
        except Exception:
            pass


def alloc_grid(nx, ny, nf=2, precision="f64", device=None, log=False) -> LatticeGrid:
    """alloc_grid(grid,nx,ny,nf,log) (src/fvm_bardow.F90:129-202).  `log` is accepted for signature
    parity; the log file of the reference is host I/O and is not opened here."""
    g = LatticeGrid()
    h = C.c_void_p()
    prec = _PREC[precision]
    if device is None:
        check(lib.plbm_alloc_grid(C.byref(h), nx, ny, nf, prec), "alloc_grid")
    else:
        check(lib.plbm_alloc_grid_on(C.byref(h), nx, ny, nf, prec, int(device)), "alloc_grid")
    g._h = h
    g.nx, g.ny, g.nf = nx, ny, nf
    g.ld = (ny + 15) // 16 * 16
    g.dtype = np.float64 if prec == capi.F64 else np.float32
    g.rho = np.zeros((nx, ny), dtype=g.dtype)
    g.ux = np.zeros((nx, ny), dtype=g.dtype)
    g.uy = np.zeros((nx, ny), dtype=g.dtype)
    return g


def dealloc_grid(grid: LatticeGrid) -> None:
    if grid._h is not None:
        h, grid._h = grid._h, None
        check(lib.plbm_dealloc_grid(h), "dealloc_grid")


def set_properties(grid, nu, dt, magic=None) -> None:
    check(lib.plbm_set_properties(grid._h, float(nu), float(dt), 0.0 if magic is None else float(magic),
                                  0 if magic is None else 1), "set_properties")


def set_pdf_to_equilibrium(grid) -> None:
    """f(:,:,:,iold) = equilibrium(rho,ux,uy) from the host views grid.rho/ux/uy."""
    r, u, v = (np.ascontiguousarray(a, dtype=grid.dtype) for a in (grid.rho, grid.ux, grid.uy))
    check(lib.plbm_set_pdf_to_equilibrium(grid._h, _ptr(r), _ptr(u), _ptr(v)), "set_pdf_to_equilibrium")


# ---- kernels with the reference's `subroutine name(grid)` interface ------------------------
def lbm_stream(grid):
    check(lib.plbm_lbm_stream(grid._h), "lbm_stream")


def stream_fvm_bardow(grid):
    check(lib.plbm_stream_fvm_bardow(grid._h), "stream_fvm_bardow")


def stream_fdm_bardow(grid):
    """stream_fdm_bardow (src/fvm_bardow.F90:511-685, default build: Lax-Wendroff finite differences)."""
    check(lib.plbm_stream_fdm_bardow(grid._h), "stream_fdm_bardow")


def stream_fdm_sofonea(grid):
    """stream_fdm_sofonea (src/fvm_bardow.F90:688-893): characteristic-wise Lax-Wendroff."""
    check(lib.plbm_stream_fdm_sofonea(grid._h), "stream_fdm_sofonea")


def collide_bgk(grid):
    check(lib.plbm_collide(grid._h, capi.BGK), "collide_bgk")


def collide_trt(grid):
    check(lib.plbm_collide(grid._h, capi.TRT), "collide_trt")


def collide_rr(grid):
    check(lib.plbm_collide(grid._h, capi.RR), "collide_rr")


def collide_trt_split(grid):
    """collide_trt as built with -DSPLIT (trt_split, src/collision_trt.F90:162-290)."""
    check(lib.plbm_collide(grid._h, capi.TRT_SPLIT), "collide_trt_split")


def collide_bgk_improved(grid):
    """collide_bgk_improved (src/collision_bgk_improved.f90:16-107)."""
    check(lib.plbm_collide(grid._h, capi.BGK_IMPROVED), "collide_bgk_improved")


def collide_bgk_split(grid):
    """collide_bgk as built with -DSPLIT (bgk_kernel_cache, src/collision_bgk.F90:84-176)."""
    check(lib.plbm_collide(grid._h, capi.BGK_SPLIT), "collide_bgk_split")


def dugks_collide(grid):
    check(lib.plbm_dugks_collide(grid._h, int(grid.dugks)), "dugks_collide")


def dugks_stream(grid):
    check(lib.plbm_dugks_stream(grid._h, int(grid.dugks)), "dugks_stream")


_COLLISION_ID = {collide_bgk: capi.BGK, collide_trt: capi.TRT, collide_rr: capi.RR, collide_bgk_split: capi.BGK_SPLIT,
                 collide_trt_split: capi.TRT_SPLIT, collide_bgk_improved: capi.BGK_IMPROVED}


_STREAMING_ID = {stream_fvm_bardow: capi.STREAM_FVM_BARDOW, stream_fdm_bardow: capi.STREAM_FDM_BARDOW,
                 stream_fdm_sofonea: capi.STREAM_FDM_SOFONEA}


def _swap(grid):
    check(lib.plbm_swap(grid._h), "swap")


# ---- orchestrators ----------------------------------------------------------------------------
def perform_lbm_step(grid, nsteps=1) -> None:
    """perform_lbm_step (src/periodic_lbm.f90:15-29): streaming(); collision(); swap.  When the two
    procedure pointers are a pair this library fuses (lbm_stream + collide_*), one kernel does the
    whole step; otherwise the procedures are called one after the other like the reference does."""
    cid = _COLLISION_ID.get(grid.collision)
    if grid.streaming is lbm_stream and cid is not None:
        check(lib.plbm_perform_lbm_step(grid._h, cid, int(nsteps)), "perform_lbm_step")
    elif grid.streaming in _STREAMING_ID and cid is not None:
        check(lib.plbm_perform_step(grid._h, _STREAMING_ID[grid.streaming], cid, int(nsteps)), "perform_step")
    else:
        for _ in range(int(nsteps)):
            grid.streaming(grid)
            grid.collision(grid)
            _swap(grid)


def perform_step(grid, nsteps=1) -> None:
    """perform_step (src/fvm_bardow.F90:307-320): identical sequence to perform_lbm_step."""
    perform_lbm_step(grid, nsteps)


def perform_triple_step(grid, nsteps=1) -> None:
    """perform_triple_step (src/fvm_bardow.F90:322-340) on a grid allocated with nf=3: the streamed,
    pre-collision PDFs stay available in lattice `imid` after the index rotation."""
    cid = _COLLISION_ID.get(grid.collision)
    sid = {lbm_stream: capi.STREAM_LBM, **_STREAMING_ID}.get(grid.streaming)
    if cid is None or sid is None:
        raise capi.PlbmError("perform_triple_step: needs lbm_stream|stream_fvm_bardow and collide_bgk|trt|rr")
    check(lib.plbm_perform_triple_step(grid._h, sid, cid, int(nsteps)), "perform_triple_step")


def perform_dugks_step(grid, nsteps=1) -> None:
    """perform_dugks_step (src/periodic_dugks.F90:25-38): collision(); streaming(); swap."""
    if grid.collision in (None, dugks_collide) and grid.streaming in (None, dugks_stream):
        check(lib.plbm_perform_dugks_step(grid._h, int(grid.dugks), int(nsteps)), "perform_dugks_step")
    else:
        for _ in range(int(nsteps)):
            grid.collision(grid)
            grid.streaming(grid)
            _swap(grid)


def update_macros(grid, lagged=None) -> None:
    """update_macros (src/fvm_bardow.F90:343-390) into the host views grid.rho/ux/uy.  By default it
    reads lattice `inew` exactly like the reference (the state before the last step);
    lagged=False reads the current state."""
    lag = grid.lagged_macros if lagged is None else lagged
    check(lib.plbm_update_macros(grid._h, _ptr(grid.rho), _ptr(grid.ux), _ptr(grid.uy), int(lag)), "update_macros")


def _vorticity(order, ux, uy):
    ux = np.ascontiguousarray(ux)
    uy = np.ascontiguousarray(uy, dtype=ux.dtype)
    nx, ny = ux.shape
    g = alloc_grid(nx, ny, precision=ux.dtype.type)
    try:
        om = np.empty_like(ux)
        check(lib.plbm_vorticity_host(g._h, order, _ptr(ux), _ptr(uy), _ptr(om)), "vorticity")
    finally:
        dealloc_grid(g)
    return om


def vorticity_2nd(ux, uy, grid=None):
    """vorticity_2nd(ux,uy,omega) (src/vorticity.f90:13-43).  With `grid`, the device-resident
    velocity of that grid is used and nothing is uploaded."""
    if grid is not None:
        om = np.empty((grid.nx, grid.ny), dtype=grid.dtype)
        check(lib.plbm_vorticity(grid._h, 2, _ptr(om)), "vorticity_2nd")
        return om
    return _vorticity(2, ux, uy)


def vorticity_4th(ux, uy, grid=None):
    """vorticity_4th (src/vorticity.f90:46-87), weights exactly as shipped."""
    if grid is not None:
        om = np.empty((grid.nx, grid.ny), dtype=grid.dtype)
        check(lib.plbm_vorticity(grid._h, 4, _ptr(om)), "vorticity_4th")
        return om
    return _vorticity(4, ux, uy)


# ---- scalar helpers of collision_trt (src/collision_trt.F90:24-34) -----------------------------
def magic_number(le, ld, dtype=np.float64):
    T = dtype
    le, ld = T(le), T(ld)
    return (T(2) - le) * (T(2) - ld) / (T(4) * le * ld)


def lambda_d(omega, x, dtype=np.float64):
    T = dtype
    omega, x = T(omega), T(x)
    return (T(4) - T(2) * omega) / (T(4) * x * omega + T(2) - omega)
