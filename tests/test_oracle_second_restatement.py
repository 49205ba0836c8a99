"""The C oracle against a second, independent numpy restatement of the reference's LBM path
(tests/numpy_restatement.py): bit for bit, fp64 and fp32.  This is the pin for the routines the reference ships no
golden data for (lbm_stream, collide_trt, collide_rr) -- two separate readings of the Fortran agreeing
to the last bit leave the Fortran text itself as the only common source."""
import numpy as np
import pytest

import numpy_restatement as nr
from conftest import random_state
from oracle.oracle import Oracle

SHAPES = [(64, 64), (67, 53), (5, 3), (16, 130)]


def start(prec, nx, ny):
    o = Oracle(prec)
    p = o.set_properties(0.02, 1.0, 0.25)
    f = random_state(o, nx, ny)
    return o, p, f


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("nx,ny", SHAPES)
def test_stream_and_macros(prec, nx, ny):
    o, p, f = start(prec, nx, ny)
    g = o.alloc_f(nx, ny)
    o.lbm_stream(f, g, ny)
    assert np.array_equal(g[:, :, :ny], nr.lbm_stream(f[:, :, :ny]))
    r, u, v = o.update_macros(f, ny)
    r2, u2, v2 = nr.update_macros(f[:, :, :ny])
    assert np.array_equal(r, r2) and np.array_equal(u, u2) and np.array_equal(v, v2)


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("nx,ny", SHAPES)
@pytest.mark.parametrize("omega", [0.6, 1.0, 1.95662121778720])
def test_collisions(prec, nx, ny, omega):
    o, p, f = start(prec, nx, ny)
    T = o.dtype
    for name in ("bgk", "trt", "rr"):
        a = f.copy()
        if name == "bgk":
            o.collide_bgk(a, ny, T(omega))
            b = nr.collide_bgk(f[:, :, :ny], omega)
        elif name == "trt":
            o.collide_trt(a, ny, T(omega), T(0.25))
            b = nr.collide_trt(f[:, :, :ny], omega, 0.25)
        else:
            o.collide_rr(a, ny, T(omega))
            b = nr.collide_rr(f[:, :, :ny], omega)
        assert b.dtype == a.dtype
        assert np.array_equal(a[:, :, :ny], b), (name, np.abs(a[:, :, :ny] - b).max())


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("name", ["bgk", "trt", "rr"])
def test_fifty_steps(prec, name):
    """stream + collide, 50 steps on a 37 x 22 grid: the two restatements stay bit-identical"""
    nx, ny = 37, 22
    o, p, f = start(prec, nx, ny)
    T = o.dtype
    a, b = f.copy(), o.alloc_f(nx, ny)
    n = f[:, :, :ny].copy()
    for _ in range(50):
        o.lbm_stream(a, b, ny)
        n = nr.lbm_stream(n)
        if name == "bgk":
            o.collide_bgk(b, ny, p["omega"])
            n = nr.collide_bgk(n, p["omega"])
        elif name == "trt":
            o.collide_trt(b, ny, p["omega"], p["trt_magic"])
            n = nr.collide_trt(n, p["omega"], p["trt_magic"])
        else:
            o.collide_rr(b, ny, p["omega"])
            n = nr.collide_rr(n, p["omega"])
        a, b = b, a
    assert np.array_equal(a[:, :, :ny], n)
    assert np.isfinite(n).all()


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("nx,ny", SHAPES)
def test_vorticity(prec, nx, ny):
    o = Oracle(prec)
    rng = np.random.default_rng(7)
    ux = rng.standard_normal((nx, ny)).astype(o.dtype)
    uy = rng.standard_normal((nx, ny)).astype(o.dtype)
    assert np.array_equal(o.vorticity(ux, uy, order=2), nr.vorticity_2nd(ux, uy))
    assert np.array_equal(o.vorticity(ux, uy, order=4), nr.vorticity_4th(ux, uy))


def test_sim_plugin_seam():
    """the DDF-shifted path behind c_slbm_* (sim/sim.F90): init, 25 collide + push-stream + halo-fold steps, macros"""
    nx, ny, steps, omega = 48, 40, 25, 1.7
    o = Oracle("f64")
    rng = np.random.default_rng(11)
    p = 1e-3 * rng.standard_normal((ny, nx))
    u = 0.05 * rng.standard_normal((2, ny, nx))
    f1 = np.zeros((9, ny + 2, nx + 2))
    f2 = np.zeros_like(f1)
    P = lambda a: a.ctypes.data  # noqa: E731
    o._sim_eqinit(nx, ny, P(f1), P(p), P(u[0]), P(u[1]))
    n = nr.sim_eqinit(p, u[0], u[1])
    assert np.array_equal(f1[:, 1:ny + 1, 1:nx + 1], n)
    f2[...] = f1
    for _ in range(steps):
        o._sim_step(nx, ny, P(f1), P(f2), omega)
        o._sim_bc(nx, ny, P(f2))
        f1, f2 = f2, f1
        n = nr.sim_step(n, omega)
    assert np.array_equal(f1[:, 1:ny + 1, 1:nx + 1], n)
    rho_w, u_w, v_w = np.zeros((ny, nx)), np.zeros((ny, nx)), np.zeros((ny, nx))
    o._sim_macros(nx, ny, P(f1), P(rho_w), P(u_w), P(v_w))
    r, a, b = nr.sim_macros(n)
    assert np.array_equal(rho_w, r) and np.array_equal(u_w, a) and np.array_equal(v_w, b)


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("n", [64, 5, 37])
def test_bardow_fvm_with_bgk(prec, n):
    """perform_step with stream_fvm_bardow + collide_bgk (what app/main_vortex.f90 runs), 20 steps, dt = 0.3"""
    o = Oracle(prec)
    p = o.set_properties(0.02, 0.3, 0.25)
    f = random_state(o, n, n)
    a, b = f.copy(), o.alloc_f(n, n)
    m = f[:, :, :n].copy()
    for _ in range(20):
        o.stream_fvm_bardow(a, b, n, p["dt"])
        o.collide_bgk(b, n, p["omega"])
        m = nr.collide_bgk(nr.stream_fvm_bardow(m, p["dt"]), p["omega"])
        a, b = b, a
    assert np.array_equal(a[:, :, :n], m)


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("n,ratio", [(64, 5.0), (5, 1.0), (37, 10.0)])
def test_dugks(prec, n, ratio):
    """perform_dugks_step (-DDUGKS) x 20 on an n x n grid with dt = ratio * tau: both lattices, bit for bit"""
    o = Oracle(prec)
    T = o.dtype
    nu = T(0.004)
    tau = T(3) * nu
    p = o.set_properties(nu, T(ratio) * tau, 0.25)
    f = random_state(o, n, n)
    a, b = f.copy(), o.alloc_f(n, n)  # a = iold (ftilde), b = inew
    m = f[:, :, :n].copy()
    for _ in range(20):
        o.dugks_collide(a, b, n, p["omega"], p["tau"], p["dt"], True)
        o.dugks_stream(a, b, n, p["tau"], p["dt"], True)
        a, b = b, a  # swap: iold = the new ftilde, inew = fbar+
        m, fbar = nr.dugks_step(m, p["omega"], p["tau"], p["dt"])
        assert np.array_equal(a[:, :, :n], m)
        assert np.array_equal(b[:, :, :n], fbar)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_properties_and_equilibrium_init(prec):
    o = Oracle(prec)
    T = o.dtype
    for nu, dt, magic in [(0.02, 1.0, 0.25), (0.0036950, 0.05542, None), (1.8919, 1.0, 0.25), (0.47297, 1.0, None)]:
        a, b = o.set_properties(nu, dt, magic), nr.set_properties(T, nu, dt, magic)
        assert all(a[k] == b[k] and type(b[k]) is T for k in ("tau", "omega", "trt_magic", "csqr")), (a, b)
    rng = np.random.default_rng(3)
    nx, ny = 19, 23
    rho = (0.9 + 0.2 * rng.random((nx, ny))).astype(T)
    ux = (0.1 * (rng.random((nx, ny)) - 0.5)).astype(T)
    uy = (0.1 * (rng.random((nx, ny)) - 0.5)).astype(T)
    f = o.alloc_f(nx, ny)
    o.set_pdf_to_equilibrium(rho, ux, uy, f)
    assert np.array_equal(f[:, :, :ny], np.stack(nr.equilibrium(T, rho, ux, uy)))


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("stencil", ["default", "wls", "wls_gauss_v1", "wls_gauss_v2", "iso", "sofonea"])
def test_fdm_streaming_schemes(prec, stencil):
    """stream_fdm_bardow (every cpp build) and stream_fdm_sofonea fused with collide_trt, 10 steps, dt = 0.3"""
    n = 29
    o = Oracle(prec)
    p = o.set_properties(0.02, 0.3, 0.25)
    f = random_state(o, n, n)
    a, b = f.copy(), o.alloc_f(n, n)
    m = f[:, :, :n].copy()
    for _ in range(10):
        if stencil == "sofonea":
            o.stream_fdm_sofonea(a, b, n, p["dt"])
            m = nr.stream_fdm_sofonea(m, p["dt"])
        else:
            o.stream_fdm_bardow(a, b, n, p["dt"], Oracle.FDM_STENCILS[stencil])
            m = nr.stream_fdm_bardow(m, p["dt"], stencil)
        o.collide_trt(b, n, p["omega"], p["trt_magic"])
        m = nr.collide_trt(m, p["omega"], p["trt_magic"])
        a, b = b, a
    assert np.array_equal(a[:, :, :n], m), np.abs(a[:, :, :n] - m).max()


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("omega", [0.6, 1.0, 1.95662121778720])
def test_split_and_improved_collisions(prec, omega):
    """the -DSPLIT builds (bgk_kernel_cache, trt_split) and collide_bgk_improved"""
    nx, ny = 23, 37
    o, p, f = start(prec, nx, ny)
    T = o.dtype
    a = f.copy()
    o.kernel_bgk(a, ny, T(omega))
    assert np.array_equal(a[:, :, :ny], nr.collide_bgk_split(f[:, :, :ny], omega))
    a = f.copy()
    o.collide_trt_split(a, ny, T(omega), T(0.25))
    assert np.array_equal(a[:, :, :ny], nr.collide_trt_split(f[:, :, :ny], omega, 0.25))
    a = f.copy()
    o.collide_bgk_improved(a, ny, T(omega))
    assert np.array_equal(a[:, :, :ny], nr.collide_bgk_improved(f[:, :, :ny], omega))


@pytest.mark.parametrize("dt", [1.0, 0.4])
def test_sim_lw_plugin(dt):
    """the 2nd-order Lax-Wendroff plugin behind c_lw_* (sim/sim_lw.F90): stream, collide, periodic halo; 12 steps"""
    nx, ny, steps, omega = 40, 56, 12, 1.4
    o = Oracle("f64")
    rng = np.random.default_rng(12)
    p = 1e-3 * rng.standard_normal((ny, nx))
    u = 0.05 * rng.standard_normal((2, ny, nx))
    f1 = np.zeros((9, ny + 2, nx + 2))
    f2 = np.zeros_like(f1)
    P = lambda a: a.ctypes.data  # noqa: E731
    o._sim_eqinit(nx, ny, P(f1), P(p), P(u[0]), P(u[1]))
    o._lw_bc(nx, ny, P(f1))
    n = nr.sim_eqinit(p, u[0], u[1])
    for _ in range(steps):
        o._lw_stream(nx, ny, P(f1), P(f2), dt)
        o._lw_collision(nx, ny, P(f2), omega)
        o._lw_bc(nx, ny, P(f2))
        f1, f2 = f2, f1
        n = nr.lw_collision(nr.lw_stream(n, dt), omega)
    assert np.array_equal(f1[:, 1:ny + 1, 1:nx + 1], n)
