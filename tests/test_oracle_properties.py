"""Physics / consistency properties of the oracle kernels that no reference artefact pins (the
restatement is the only pin): conservation laws, agreement between algebraically equal variants, and
exact limits of the streaming schemes."""
import numpy as np
import pytest

from conftest import random_state
from oracle.oracle import Oracle


def moments(o, f, ny):
    rho, ux, uy = o.update_macros(f, ny)
    return rho, rho * ux, rho * uy


@pytest.mark.parametrize("prec,tol", [("f64", 1e-14), ("f32", 2e-6)])
@pytest.mark.parametrize("name", ["collide_bgk", "kernel_bgk", "collide_rr", "collide_bgk_improved", "collide_trt", "collide_trt_split"])
def test_collisions_conserve_mass_and_momentum(prec, tol, name):
    o = Oracle(prec)
    nx, ny = 23, 37
    f = random_state(o, nx, ny)
    f = np.nan_to_num(f, nan=0.0)
    g = f.copy()
    args = (1.3, 0.25) if "trt" in name else (1.3,)
    getattr(o, name)(g, ny, *args)
    if "trt" in name:  # incompressible equilibrium: the conserved momentum is the first moment itself
        j = lambda h: (h[:, :, :ny].sum(0), (h[1] - h[3] + h[5] - h[6] - h[7] + h[8])[:, :ny], (h[2] - h[4] + h[5] + h[6] - h[7] - h[8])[:, :ny])  # noqa: E731
        a, b = j(f), j(g)
    else:
        a, b = moments(o, f, ny), moments(o, g, ny)
    for x, y in zip(a, b):
        assert np.abs(x - y).max() < tol * max(1.0, np.abs(x).max())


def test_split_variants_agree_with_the_default_kernels_to_round_off():
    """-DSPLIT changes the association of a few products, not the mathematics."""
    o = Oracle("f64")
    nx, ny = 19, 33
    f = np.nan_to_num(random_state(o, nx, ny), nan=0.0)
    a, b = f.copy(), f.copy()
    o.collide_bgk(a, ny, 1.7)
    o.kernel_bgk(b, ny, 1.7)
    assert 0 < np.abs(a - b).max() < 1e-15  # different rounding, same maths
    a, b = f.copy(), f.copy()
    o.collide_trt(a, ny, 1.7, 0.25)
    o.collide_trt_split(b, ny, 1.7, 0.25)
    assert np.abs(a - b).max() < 1e-15


def test_trt_with_bgk_magic_relaxes_both_parities_equally():
    """With Lambda = ((2 - w)/(2 w))^2 both TRT rates equal w: the commented default of set_properties."""
    o = Oracle("f64")
    w = 1.3
    magic = ((2.0 - w) / (2.0 * w)) ** 2
    assert abs(float(o.lambda_d(w, magic)) - w) < 1e-14
    assert abs(float(o.magic_number(w, w)) - magic) < 1e-14


@pytest.mark.parametrize("name", ["stream_fdm_bardow", "stream_fdm_sofonea", "stream_fvm_bardow"])
def test_streaming_schemes_conserve_every_population(name):
    o = Oracle("f64")
    nx = ny = 24
    f = np.nan_to_num(random_state(o, nx, ny), nan=0.0)
    g = np.zeros_like(f)
    getattr(o, name)(f, g, ny, 0.37)
    assert np.abs(g[:, :, :ny].sum((1, 2)) - f[:, :, :ny].sum((1, 2))).max() < 1e-12


def test_lax_wendroff_schemes_reduce_to_exact_streaming_at_unit_cfl():
    """dt = 1: the Lax-Wendroff update of an axis population IS the lattice shift (lbm_stream);
    stream_fdm_sofonea also for the diagonals along their own characteristic."""
    o = Oracle("f64")
    nx = ny = 20
    f = np.nan_to_num(random_state(o, nx, ny), nan=0.0)
    exact, lw, sof = (np.zeros_like(f) for _ in range(3))
    o.lbm_stream(f, exact, ny)
    o.stream_fdm_bardow(f, lw, ny, 1.0)
    o.stream_fdm_sofonea(f, sof, ny, 1.0)
    assert np.abs(lw[:5, :, :ny] - exact[:5, :, :ny]).max() < 1e-15
    assert np.abs(sof[:5, :, :ny] - exact[:5, :, :ny]).max() < 1e-15


def test_vorticity_of_the_taylor_green_field():
    """omega = d(uy)/dx - d(ux)/dy of the TG field is 2 k umax cos(kx) cos(ky) up to O(k^2)."""
    o = Oracle("f64")
    n = 128
    k = 2 * np.pi / n
    _, ux, uy = o.taylor_green_eval(n, n, k, k, 0.01, 1.0, 0.0)
    om = o.vorticity(ux, uy, 2)
    x = np.arange(n) + 0.5
    exact = 2 * k * 0.01 * np.cos(k * x)[:, None] * np.cos(k * x)[None, :]
    assert np.abs(om - exact).max() < 1e-3 * np.abs(exact).max()


@pytest.mark.parametrize("order", [4, 6])
def test_higher_order_lax_wendroff_stencils_are_exact_for_polynomials(order):
    """lw4_stream / lw6_stream (sim/sim_lw4.F90:26-115, sim/sim_lw6.F90:26-127): the finite-difference
    stencils differentiate polynomials of total degree <= order exactly, so one streaming step of such a field
    equals its second-order Taylor shift  f - v.grad f + 1/2 v v : grad grad f  (the Lax-Wendroff update)."""
    o = Oracle("f64")
    H = order // 2
    nx, ny, dt = 9, 7, 0.37
    rng = np.random.default_rng(order)
    powers = [(a, b) for a in range(order + 1) for b in range(order + 1 - a)]
    coef = {ab: rng.standard_normal() for ab in powers}
    # coordinates centred on the grid keep the monomials O(1..1e4); the halo is filled with the polynomial itself
    X = (np.arange(1 - H, nx + H + 1) - (nx + 1) / 2.0)[None, :]
    Y = (np.arange(1 - H, ny + H + 1) - (ny + 1) / 2.0)[:, None]

    def poly(da=0, db=0):
        out = np.zeros((ny + 2 * H, nx + 2 * H))
        for (a, b), c in coef.items():
            if a < da or b < db:
                continue
            ca = np.prod([a - i for i in range(da)]) if da else 1.0
            cb = np.prod([b - i for i in range(db)]) if db else 1.0
            out += c * ca * cb * X ** (a - da) * Y ** (b - db)
        return out

    fsrc = np.ascontiguousarray(np.broadcast_to(poly(), (9, ny + 2 * H, nx + 2 * H)))
    fdst = np.zeros_like(fsrc)
    o._lwh_stream(order, nx, ny, fsrc.ctypes.data, fdst.ctypes.data, dt)
    cx = [0, 1, 0, -1, 0, 1, -1, -1, 1]
    cy = [0, 0, 1, 0, -1, 1, 1, -1, -1]
    fx, fy, fxx, fxy, fyy = poly(1, 0), poly(0, 1), poly(2, 0), poly(1, 1), poly(0, 2)
    inner = (slice(H, H + ny), slice(H, H + nx))
    scale = np.abs(fsrc[0][inner]).max()
    for k in range(9):
        vx, vy = dt * cx[k], dt * cy[k]
        want = poly() - vx * fx - vy * fy + 0.5 * vx * vx * fxx + vx * vy * fxy + 0.5 * vy * vy * fyy
        assert np.abs(fdst[k][inner] - want[inner]).max() < 1e-11 * scale, k


@pytest.mark.parametrize("order", [4, 6])
def test_higher_order_lax_wendroff_halo_is_the_periodic_image(order):
    """lw4_bc / lw6_bc fill the H-wide halo (corners included) with the periodic image of the interior."""
    o = Oracle("f64")
    H = order // 2
    nx, ny = 7, 5
    rng = np.random.default_rng(3)
    inner = rng.standard_normal((9, ny, nx))
    f = np.full((9, ny + 2 * H, nx + 2 * H), np.nan)
    f[:, H:H + ny, H:H + nx] = inner
    o._lwh_bc(nx, ny, H, f.ctypes.data)
    jj = (np.arange(-H, ny + H) % ny)[:, None]
    ii = (np.arange(-H, nx + H) % nx)[None, :]
    assert np.array_equal(f, inner[:, jj, ii])


def _simfvm_run(o, nx, ny, p, u, dt, omega, steps):
    """sim_fvm%init + steps x sim_fvm%step (sim/sim_fvm.F90:255-322) on the oracle's haloed arrays."""
    H = 2
    f1 = np.zeros((9, ny + 2 * H, nx + 2 * H))
    f2, fc = np.zeros_like(f1), np.zeros_like(f1)
    P = lambda a: a.ctypes.data  # noqa: E731
    o._simh_eqinit(nx, ny, H, P(f1), P(p), P(u[0]), P(u[1]))
    o._lwh_bc(nx, ny, H, P(f1))
    for _ in range(steps):
        o._simfvm_step(nx, ny, P(f1), P(f2), P(fc), dt, omega)
        f1, fc = fc, f1  # the move_alloc swap
    return f1


def test_heun_finite_volume_plugin_conserves_mass_and_momentum():
    """Central fluxes telescope over the periodic grid and the collision conserves rho and rho*u, so the
    Heun predictor-corrector step of sim/sim_fvm.F90 conserves both."""
    o = Oracle("f64")
    nx, ny, H = 20, 14, 2
    rng = np.random.default_rng(5)
    p = 1e-3 * rng.standard_normal((ny, nx))
    u = 0.05 * rng.standard_normal((2, ny, nx))
    cx = np.array([0, 1, 0, -1, 0, 1, -1, -1, 1.0])[:, None, None]
    f0 = _simfvm_run(o, nx, ny, p, u, 0.3, 1.2, 0)[:, H:H + ny, H:H + nx]
    f5 = _simfvm_run(o, nx, ny, p, u, 0.3, 1.2, 5)[:, H:H + ny, H:H + nx]
    assert abs(f5.sum() - f0.sum()) < 1e-12
    assert abs((cx * f5).sum() - (cx * f0).sum()) < 1e-12
    assert np.abs(f5 - f0).max() > 1e-6  # something did happen


def test_heun_finite_volume_plugin_keeps_a_uniform_state():
    o = Oracle("f64")
    nx, ny, H = 9, 7, 2
    p = np.full((ny, nx), 2e-3)
    u = np.stack([np.full((ny, nx), 0.03), np.full((ny, nx), -0.02)])
    f0 = _simfvm_run(o, nx, ny, p, u, 0.4, 1.5, 0)
    f3 = _simfvm_run(o, nx, ny, p, u, 0.4, 1.5, 3)
    assert np.abs(f3 - f0).max() < 1e-16


@pytest.mark.parametrize("coll", ["bgk", "trt", "rr"])
def test_lbm_taylor_green_decay_matches_the_analytic_solution(coll):
    """lbm_stream + collide_bgk / collide_trt / collide_rr are the rows of SURVEY 8(c) that no golden file of the reference pins.
    A pin that needs no restatement at all: the Taylor-Green vortex of app/main_taylor_green.f90 decays as umax exp(-t / td) with
    td = 1 / (nu (kx^2 + ky^2)) and nu = csqr (tau - dt / 2) -- the streaming directions, the velocity set, the relaxation rates of
    set_properties and all three collisions must be right for the ORACLE to follow it.  48^2, fp64, one decay half-life."""
    from oracle.oracle import OracleGrid, taylor_green_setup, taylor_green_steps_to_tmax

    n = 48
    og = OracleGrid(n, n, "f64")
    s = taylor_green_setup(og.o, n, dt=1.0)
    og.set_properties(s["nu"], s["dt"], magic=0.25)
    pr, ux, uy = og.o.taylor_green_eval(n, n, s["kx"], s["ky"], s["umax"], s["td"], 0.0)
    og.rho, og.ux, og.uy = pr / og.props["csqr"] + 1.0, ux, uy
    og.set_pdf_to_equilibrium()
    steps, t = taylor_green_steps_to_tmax(og.o, s)
    og.run(Oracle.SCHEME_LBM, {"bgk": Oracle.BGK, "trt": Oracle.TRT, "rr": Oracle.RR}[coll], steps)
    rho, u, v = og.update_macros(lagged=False)
    _, uxa, uya = og.o.taylor_green_eval(n, n, s["kx"], s["ky"], s["umax"], s["td"], t)
    speed, want = np.hypot(u, v).max(), np.hypot(uxa, uya).max()  # both sampled at the cell centres
    assert abs(want / float(s["umax"]) - 0.5) < 5e-3  # one half-life: exp(-t / td) = 1/2 up to the last step and the sampling
    assert abs(speed - want) / want < 5e-3, (coll, speed, want)
    assert og.o.l2_norm(u, v, uxa, uya) < 5e-3  # relative L2 error of the velocity field (second order in 2 pi / n)
    assert abs(rho.sum() - n * n) / (n * n) < 1e-11  # mass, up to the round-off of 7,299 steps
    assert abs(u.sum()) < 1e-10 and abs(v.sum()) < 1e-10  # no net momentum
