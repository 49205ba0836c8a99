"""GPU parity: the CUDA library (through its C ABI) against the CPU oracle on the same seeded inputs.

Bar: fp64 AND fp32 results are BIT-IDENTICAL to the oracle (the library is compiled with
-fmad=false and evaluates every node in the reference's operation order), which is stronger than
the north-star tolerance (1e-12 relative fp64, 1e-5 fp32).
"""
import os

import numpy as np
import pytest

from conftest import random_state
from oracle.oracle import Oracle, OracleGrid, padded_ld, taylor_green_setup

pytestmark = pytest.mark.gpu

SIZES = [(64, 64), (67, 53), (16, 130), (5, 3), (128, 256), (2, 2)]
PRECS = ["f64", "f32"]


def make_pair(plbm, nx, ny, prec, nu=0.02, dt=1.0, magic=0.25, seed=None):
    """Oracle grid + device grid holding the same random lattice in `iold`."""
    og = OracleGrid(nx, ny, prec)
    og.set_properties(nu, dt, magic)
    f0 = random_state(og.o, nx, ny) if seed is None else random_state(og.o, nx, ny, seed=seed)
    og.lattice(og.iold)[...] = f0
    og.lattice(og.inew)[...] = 0
    g = plbm.alloc_grid(nx, ny, precision=prec)
    plbm.set_properties(g, nu, dt, magic)
    g.upload_f(g.iold, np.nan_to_num(f0, nan=0.0))
    g.upload_f(g.inew, np.zeros_like(f0))
    return og, g


def collisions(plbm):
    """(device procedure, oracle collision id) for every collision operator of the reference."""
    return ((plbm.collide_bgk, Oracle.BGK), (plbm.collide_trt, Oracle.TRT), (plbm.collide_rr, Oracle.RR),
            (plbm.collide_bgk_split, Oracle.BGK_SPLIT), (plbm.collide_trt_split, Oracle.TRT_SPLIT),
            (plbm.collide_bgk_improved, Oracle.BGK_IMPROVED))


def oracle_collide(o, f, ny, p, ocoll):
    if ocoll == Oracle.BGK:
        o.collide_bgk(f, ny, p["omega"])
    elif ocoll == Oracle.TRT:
        o.collide_trt(f, ny, p["omega"], p["trt_magic"])
    elif ocoll == Oracle.RR:
        o.collide_rr(f, ny, p["omega"])
    elif ocoll == Oracle.BGK_SPLIT:
        o.kernel_bgk(f, ny, p["omega"])
    elif ocoll == Oracle.TRT_SPLIT:
        o.collide_trt_split(f, ny, p["omega"], p["trt_magic"])
    else:
        o.collide_bgk_improved(f, ny, p["omega"])


def assert_same_lattice(g, og, which_g, which_o, ny):
    got = g.download_f(which_g)[:, :, :ny]
    want = og.lattice(which_o)[:, :, :ny]
    assert np.array_equal(got, want), f"max abs diff {np.abs(got - want).max():.3e} in {int((got != want).sum())} values"


@pytest.mark.parametrize("prec", PRECS)
def test_properties_match_oracle(plbm, prec):
    o = Oracle(prec)
    for nu, dt, magic in [(0.02, 1.0, 0.25), (0.0036950, 0.05542, None), (1.8919, 1.0, 0.25)]:
        want = o.set_properties(nu, dt, magic)
        g = plbm.alloc_grid(8, 8, precision=prec)
        plbm.set_properties(g, nu, dt, magic)
        assert g.tau == want["tau"] and g.omega == want["omega"] and g.trt_magic == want["trt_magic"] and g.csqr == want["csqr"]
        plbm.dealloc_grid(g)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("nx,ny", SIZES)
def test_set_pdf_to_equilibrium_and_macros(plbm, nx, ny, prec):
    o = Oracle(prec)
    rng = np.random.default_rng(7)
    rho = (0.9 + 0.2 * rng.random((nx, ny))).astype(o.dtype)
    ux = (0.1 * (rng.random((nx, ny)) - 0.5)).astype(o.dtype)
    uy = (0.1 * (rng.random((nx, ny)) - 0.5)).astype(o.dtype)
    f = o.alloc_f(nx, ny)
    o.set_pdf_to_equilibrium(rho, ux, uy, f)
    g = plbm.alloc_grid(nx, ny, precision=prec)
    g.rho[:], g.ux[:], g.uy[:] = rho, ux, uy
    plbm.set_pdf_to_equilibrium(g)
    assert np.array_equal(g.download_f(g.iold)[:, :, :ny], f[:, :, :ny])
    plbm.update_macros(g, lagged=False)
    r2, u2, v2 = o.update_macros(f, ny)
    assert np.array_equal(g.rho, r2) and np.array_equal(g.ux, u2) and np.array_equal(g.uy, v2)
    plbm.dealloc_grid(g)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("nx,ny", SIZES)
def test_unfused_stream_then_collide_entries(plbm, nx, ny, prec):
    """lbm_stream, collide_bgk/trt/rr as separate calls + swap == oracle, bitwise."""
    for coll, ocoll in collisions(plbm):
        og, g = make_pair(plbm, nx, ny, prec)
        plbm.lbm_stream(g)
        og.o.lbm_stream(og.lattice(og.iold), og.lattice(og.inew), ny)
        assert_same_lattice(g, og, g.inew, og.inew, ny)
        coll(g)
        oracle_collide(og.o, og.lattice(og.inew), ny, og.props, ocoll)
        assert_same_lattice(g, og, g.inew, og.inew, ny)
        plbm.dealloc_grid(g)


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("nx,ny", SIZES)
def test_fused_lbm_steps(plbm, nx, ny, prec, variant):
    """perform_lbm_step (fused kernel), 7 steps, every collision model, every load variant
    (0 direct loads [+ cluster-resident kernel on small grids], 1 shuffle, 2 scalar, 3 TMA-staged tile)."""
    for coll, ocoll in collisions(plbm):
        og, g = make_pair(plbm, nx, ny, prec)
        g.set_variant(variant)
        g.collision, g.streaming = coll, plbm.lbm_stream
        plbm.perform_lbm_step(g, 7)
        og.run(Oracle.SCHEME_LBM, ocoll, 7)
        assert (g.iold, g.inew) == (og.iold, og.inew)
        assert_same_lattice(g, og, g.iold, og.iold, ny)
        assert_same_lattice(g, og, g.inew, og.inew, ny)  # the lagged lattice too
        plbm.update_macros(g)  # lagged, like the reference
        r, u, v = og.update_macros(lagged=True)
        assert np.array_equal(g.rho, r) and np.array_equal(g.ux, u) and np.array_equal(g.uy, v)
        plbm.dealloc_grid(g)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("nx,ny", [(64, 64), (16, 132), (4, 8), (9, 4), (40, 516), (300, 260), (5, 1024), (37, 2048), (70, 16), (8, 32)])
def test_two_step_kernel(plbm, nx, ny, prec):
    """The temporal-blocking kernels (two steps per pass over HBM, plbm_lbm2.cu) on grids the cluster kernel
    would otherwise take: several strips / x segments, ragged last strip, odd step counts, both lattices and
    the lagged macros bit-identical to the oracle.  Variant 5 = the library's default flavour, 6 = raw columns
    by per-thread loads (k_lbm2), 7 = by bulk async copies (k_lbm2_bulk), 8 = 7 issued as the three x ranges
    of the slab schedule (boundary lines, then the interior)."""
    for nsteps, (coll, ocoll) in zip((3, 4, 5, 8, 9, 6), collisions(plbm)):
        og, g0 = make_pair(plbm, nx, ny, prec)
        f0 = g0.download_f(g0.iold)
        plbm.dealloc_grid(g0)
        og.run(Oracle.SCHEME_LBM, ocoll, nsteps)
        r, u, v = og.update_macros(lagged=True)
        for variant in (5, 6, 7, 8):
            g = plbm.alloc_grid(nx, ny, precision=prec)
            plbm.set_properties(g, 0.02, 1.0, 0.25)
            g.upload_f(g.iold, f0)
            g.upload_f(g.inew, np.zeros_like(f0))
            g.set_variant(variant)
            g.collision, g.streaming = coll, plbm.lbm_stream
            plbm.perform_lbm_step(g, nsteps)
            assert (g.iold, g.inew) == (og.iold, og.inew), f"variant {variant}"
            assert_same_lattice(g, og, g.iold, og.iold, ny)
            assert_same_lattice(g, og, g.inew, og.inew, ny)
            plbm.update_macros(g)
            assert np.array_equal(g.rho, r) and np.array_equal(g.ux, u) and np.array_equal(g.uy, v), f"variant {variant}"
            plbm.dealloc_grid(g)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("nx,ny", [(64, 64), (16, 132), (40, 516), (300, 260), (5, 1024), (37, 2048), (70, 16), (8, 32), (4, 48)])
def test_multi_step_kernel_experimental(plbm, nx, ny, prec):
    """k_lbmn_bulk: variant 9 = pairs through its NSTEP = 2 instance, variant 10 = triples (+ pairs for the rest);
    the schedule itself is pinned on the CPU by tests/test_multi_step_schedule.py.  (First run on a B200 in round 2, green:
    profiles/r02_a_pytest_experimental.txt; the name is kept.)"""
    for nsteps, (coll, ocoll) in zip((3, 4, 5, 8, 9, 11, 7, 10, 13), collisions(plbm)[:3] * 2 + collisions(plbm)[3:]):
        og, g0 = make_pair(plbm, nx, ny, prec)
        f0 = g0.download_f(g0.iold)
        plbm.dealloc_grid(g0)
        og.run(Oracle.SCHEME_LBM, ocoll, nsteps)
        r, u, v = og.update_macros(lagged=True)
        for variant in (9, 10):
            g = plbm.alloc_grid(nx, ny, precision=prec)
            plbm.set_properties(g, 0.02, 1.0, 0.25)
            g.upload_f(g.iold, f0)
            g.upload_f(g.inew, np.zeros_like(f0))
            g.set_variant(variant)
            g.collision, g.streaming = coll, plbm.lbm_stream
            plbm.perform_lbm_step(g, nsteps)
            assert (g.iold, g.inew) == (og.iold, og.inew), f"variant {variant}"
            assert_same_lattice(g, og, g.iold, og.iold, ny)
            assert_same_lattice(g, og, g.inew, og.inew, ny)
            plbm.update_macros(g)
            assert np.array_equal(g.rho, r) and np.array_equal(g.ux, u) and np.array_equal(g.uy, v), f"variant {variant}"
            plbm.dealloc_grid(g)


@pytest.mark.parametrize("coll_name", ["bgk", "trt"])
@pytest.mark.parametrize("nsteps", [8, 10, 4])
def test_fp64_default_takes_three_steps_per_pass_and_stays_bit_identical(plbm, nsteps, coll_name):
    """Default stepping (variant 0) of collide_bgk / collide_trt / collide_rr from 512^2 nodes up advances THREE steps per pass
    over HBM (k_lbmn_bulk<3>, csrc/plbm_lbmn.cu lbm_triples_wanted): lattices, indices and lagged macros equal the oracle's bit
    for bit, and the launch count is the triples schedule's (periodic_lbm_b200/slab.py launch_schedule)."""
    nx, ny = 2048, 2048
    og, g = make_pair(plbm, nx, ny, "f64")
    coll, ocoll = {"bgk": (plbm.collide_bgk, Oracle.BGK), "trt": (plbm.collide_trt, Oracle.TRT)}[coll_name]
    g.collision, g.streaming = coll, plbm.lbm_stream
    assert g.steps_per_pass() == 3 and g.steps_per_pass(plbm.collide_rr) == 3
    assert g.steps_per_pass(plbm.collide_bgk_split) == 3 and g.steps_per_pass(plbm.collide_bgk_improved) == 3
    l0 = plbm.launch_count()
    plbm.perform_lbm_step(g, nsteps)
    launches = plbm.launch_count() - l0
    from periodic_lbm_b200.slab import launch_schedule
    # without a third lattice buffer 8: 3+2+2+1, 10: 3+3+3+1, 4: 3+1; with it (the default where the GPU has room) 8: 3+2+3, 10: 3+3+3+1, 4: 3+1
    assert launches == len(launch_schedule(nsteps, pairs=True, triples=True, dual=g.closing_triple())), launches
    og.run(Oracle.SCHEME_LBM, ocoll, nsteps)
    assert (g.iold, g.inew) == (og.iold, og.inew)
    assert_same_lattice(g, og, g.iold, og.iold, ny)
    assert_same_lattice(g, og, g.inew, og.inew, ny)
    plbm.update_macros(g)
    r, u, v = og.update_macros(lagged=True)
    assert np.array_equal(g.rho, r) and np.array_equal(g.ux, u) and np.array_equal(g.uy, v)
    plbm.dealloc_grid(g)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("nx,ny", [(64, 64), (67, 67), (5, 5), (130, 130)])
def test_fvm_bardow_steps(plbm, nx, ny, prec):
    """perform_step with stream_fvm_bardow + collide_bgk (what app/main_vortex.f90 runs), TMA-pipelined
    (variant 0) and plain-load (variant 2) tile kernels, then every other collision operator."""
    for variant in (0, 2):
        og, g = make_pair(plbm, nx, ny, prec, nu=0.02, dt=0.3)
        g.set_variant(variant)
        g.collision, g.streaming = plbm.collide_bgk, plbm.stream_fvm_bardow
        plbm.perform_step(g, 5)
        og.run(Oracle.SCHEME_FVM_BARDOW, Oracle.BGK, 5)
        assert_same_lattice(g, og, g.iold, og.iold, ny)
        plbm.dealloc_grid(g)
    for coll, ocoll in collisions(plbm)[1:]:
        og, g = make_pair(plbm, nx, ny, prec, nu=0.02, dt=0.3)
        g.collision, g.streaming = coll, plbm.stream_fvm_bardow
        plbm.perform_step(g, 3)
        og.run(Oracle.SCHEME_FVM_BARDOW, ocoll, 3)
        assert_same_lattice(g, og, g.iold, og.iold, ny)
        plbm.dealloc_grid(g)
    og, g = make_pair(plbm, nx, ny, prec, nu=0.02, dt=0.3)
    # unfused entry point
    plbm.stream_fvm_bardow(g)
    og.o.stream_fvm_bardow(og.lattice(og.iold), og.lattice(og.inew), ny, og.props["dt"])
    assert_same_lattice(g, og, g.inew, og.inew, ny)
    plbm.dealloc_grid(g)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("scheme", ["fdm_bardow", "fdm_sofonea"])
@pytest.mark.parametrize("nx,ny", [(64, 64), (67, 53), (5, 3), (40, 130)])
def test_fdm_streaming_schemes(plbm, nx, ny, prec, scheme):
    """stream_fdm_bardow / stream_fdm_sofonea (src/fvm_bardow.F90:511-893): unfused entry, fused with
    every collision (bgk/trt/rr in-kernel, the others as a second launch), TMA and plain-load kernels."""
    stream = getattr(plbm, "stream_" + scheme)
    osch = Oracle.SCHEME_FDM_BARDOW if scheme == "fdm_bardow" else Oracle.SCHEME_FDM_SOFONEA
    og, g = make_pair(plbm, nx, ny, prec, nu=0.02, dt=0.3)
    stream(g)
    getattr(og.o, "stream_" + scheme)(og.lattice(og.iold), og.lattice(og.inew), ny, og.props["dt"])
    assert_same_lattice(g, og, g.inew, og.inew, ny)
    plbm.dealloc_grid(g)
    for variant in (0, 2):
        for coll, ocoll in collisions(plbm):
            og, g = make_pair(plbm, nx, ny, prec, nu=0.02, dt=0.3)
            g.set_variant(variant)
            g.collision, g.streaming = coll, stream
            plbm.perform_step(g, 3)
            og.run(osch, ocoll, 3)
            assert (g.iold, g.inew) == (og.iold, og.inew)
            assert_same_lattice(g, og, g.iold, og.iold, ny)
            plbm.dealloc_grid(g)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("stencil", ["wls", "wls_gauss_v1", "wls_gauss_v2", "iso"])
@pytest.mark.parametrize("nx,ny", [(64, 64), (67, 53), (5, 3)])
def test_fdm_bardow_cpp_stencils(plbm, nx, ny, prec, stencil):
    """stream_fdm_bardow as built with -DFDM_WLS / -DFDM_WLS_GAUSS_V1 / -DFDM_WLS_GAUSS_V2 / -DFDM_ISO
    (src/fvm_bardow.F90:591-660): the streaming entry alone, then 4 x perform_step with collide_bgk."""
    og, g = make_pair(plbm, nx, ny, prec, nu=0.02, dt=0.4)
    g.set_fdm_stencil(stencil)
    plbm.stream_fdm_bardow(g)
    og.o.stream_fdm_bardow(og.lattice(og.iold), og.lattice(og.inew), ny, og.props["dt"], stencil)
    assert_same_lattice(g, og, g.inew, og.inew, ny)
    g.collision, g.streaming = plbm.collide_bgk, plbm.stream_fdm_bardow
    plbm.perform_step(g, 4)
    for _ in range(4):
        og.o.stream_fdm_bardow(og.lattice(og.iold), og.lattice(og.inew), ny, og.props["dt"], stencil)
        og.o.collide_bgk(og.lattice(og.inew), ny, og.props["omega"])
        og.idx[:] = og.idx[::-1].copy()  # swap(iold, inew)
    assert (g.iold, g.inew) == (og.iold, og.inew)
    assert_same_lattice(g, og, g.iold, og.iold, ny)
    plbm.dealloc_grid(g)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("scheme", ["lbm", "fvm"])
def test_perform_triple_step(plbm, prec, scheme):
    """perform_triple_step (src/fvm_bardow.F90:322-340): post-collision in iold, pre-collision in imid."""
    nx, ny, steps = 40, 48, 4
    o = Oracle(prec)
    f = [random_state(o, nx, ny), o.alloc_f(nx, ny, fill=0.0), o.alloc_f(nx, ny, fill=0.0)]
    props = o.set_properties(0.02, 0.3 if scheme == "fvm" else 1.0, 0.25)
    g = plbm.alloc_grid(nx, ny, nf=3, precision=prec)
    plbm.set_properties(g, 0.02, 0.3 if scheme == "fvm" else 1.0, 0.25)
    iold, inew, imid = g.iold, g.inew, g.imid
    assert (iold, inew, imid) == (2, 1, 3)
    g.upload_f(iold, np.nan_to_num(f[0], nan=0.0))
    lat = {iold: f[0], inew: f[1], imid: f[2]}
    g.collision = plbm.collide_rr
    g.streaming = plbm.lbm_stream if scheme == "lbm" else plbm.stream_fvm_bardow
    plbm.perform_triple_step(g, steps)
    for _ in range(steps):
        if scheme == "lbm":
            o.lbm_stream(lat[iold], lat[inew], ny)
        else:
            o.stream_fvm_bardow(lat[iold], lat[inew], ny, props["dt"])
        lat[iold][...] = lat[inew]
        o.collide_rr(lat[inew], ny, props["omega"])
        iold, inew, imid = inew, imid, iold
    assert (g.iold, g.inew, g.imid) == (iold, inew, imid)
    assert np.array_equal(g.download_f(g.iold)[:, :, :ny], lat[iold][:, :, :ny])
    assert np.array_equal(g.download_f(g.imid)[:, :, :ny], lat[imid][:, :, :ny])
    plbm.dealloc_grid(g)


@pytest.mark.parametrize("dugks", [True, False])
@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("nx,ny", [(64, 64), (67, 53), (5, 3), (34, 130)])
def test_dugks_steps(plbm, nx, ny, prec, variant, dugks):
    """perform_dugks_step: fused kernel (variant 0) and reference-structured two-pass (variant 1)."""
    og, g = make_pair(plbm, nx, ny, prec, nu=0.02, dt=0.3)
    g.set_variant(variant)
    g.dugks = dugks
    plbm.perform_dugks_step(g, 5)
    og.run(Oracle.SCHEME_DUGKS if dugks else Oracle.SCHEME_DUGKS_OFF, Oracle.BGK, 5)
    assert (g.iold, g.inew) == (og.iold, og.inew)
    assert_same_lattice(g, og, g.iold, og.iold, ny)
    # lattice inew must hold fbar^+ of the last step, exactly what the reference leaves there
    assert_same_lattice(g, og, g.inew, og.inew, ny)
    plbm.update_macros(g)
    r, u, v = og.update_macros(lagged=True)
    assert np.array_equal(g.rho, r) and np.array_equal(g.ux, u) and np.array_equal(g.uy, v)
    # stepping on after the lazy materialisation must still agree
    plbm.perform_dugks_step(g, 2)
    og.run(Oracle.SCHEME_DUGKS if dugks else Oracle.SCHEME_DUGKS_OFF, Oracle.BGK, 2)
    assert_same_lattice(g, og, g.iold, og.iold, ny)
    plbm.dealloc_grid(g)


@pytest.mark.parametrize("prec", PRECS)
def test_dugks_unfused_entries(plbm, prec):
    nx, ny = 40, 48
    og, g = make_pair(plbm, nx, ny, prec, nu=0.02, dt=0.3)
    p = og.props
    plbm.dugks_collide(g)
    og.o.dugks_collide(og.lattice(og.iold), og.lattice(og.inew), ny, p["omega"], p["tau"], p["dt"], True)
    assert_same_lattice(g, og, g.iold, og.iold, ny)
    assert_same_lattice(g, og, g.inew, og.inew, ny)
    plbm.dugks_stream(g)
    og.o.dugks_stream(og.lattice(og.iold), og.lattice(og.inew), ny, p["tau"], p["dt"], True)
    assert_same_lattice(g, og, g.inew, og.inew, ny)
    plbm.dealloc_grid(g)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("order", [2, 4])
@pytest.mark.parametrize("nx,ny", [(64, 64), (67, 53), (5, 3)])
def test_vorticity(plbm, nx, ny, prec, order):
    o = Oracle(prec)
    rng = np.random.default_rng(3)
    ux = rng.standard_normal((nx, ny)).astype(o.dtype)
    uy = rng.standard_normal((nx, ny)).astype(o.dtype)
    want = o.vorticity(ux, uy, order)
    got = (plbm.vorticity_2nd if order == 2 else plbm.vorticity_4th)(ux, uy)
    assert np.array_equal(got, want)


def test_diagnostics_reductions(plbm):
    n = 96
    og = OracleGrid(n, n)
    o = og.o
    s = taylor_green_setup(o, n, dt=1.0)
    p, ux, uy = o.taylor_green_eval(n, n, s["kx"], s["ky"], s["umax"], s["td"], 0.0)
    g = plbm.alloc_grid(n, n)
    plbm.set_properties(g, s["nu"], s["dt"], 0.25)
    g.rho[:], g.ux[:], g.uy[:] = p * 3.0 + 1.0, ux, uy
    plbm.set_pdf_to_equilibrium(g)
    d = g.diagnostics()
    sp = np.hypot(ux, uy)
    # hypot differs by at most an ulp between libm and the CUDA math library
    assert abs(d["max_speed"] - sp.max()) <= 4e-16 * sp.max() and abs(d["min_speed"] - sp.min()) <= 4e-16 * sp.max()
    assert abs(d["sum_rho"] - g.rho.sum()) / g.rho.sum() < 1e-13
    ke = 0.5 * (g.rho * (ux**2 + uy**2)).sum()
    assert abs(d["kinetic_energy"] - ke) / ke < 1e-12
    # analytic TG kinetic energy at t=0: 1/4 N umax^2 (rho ~ 1)
    assert abs(d["kinetic_energy"] - 0.25 * n * n * float(s["umax"]) ** 2) / ke < 1e-3
    _, uxa, uya = o.taylor_green_eval(n, n, s["kx"], s["ky"], s["umax"], s["td"], 100.0)
    want = float(o.l2_norm(ux, uy, uxa, uya))
    assert abs(g.l2_error(uxa, uya) - want) / want < 1e-12
    plbm.dealloc_grid(g)


def test_cases_match_oracle(plbm):
    for prec, dt in (("f64", np.float64), ("f32", np.float32)):
        o = Oracle(prec)
        s = taylor_green_setup(o, 48, dt=1.0)
        tp = plbm.taylor_green_params(48, dt=1.0, dtype=dt)
        a = o.taylor_green_eval(48, 48, s["kx"], s["ky"], s["umax"], s["td"], dt(17.0))
        b = tp["case"].eval(17.0)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        vp = plbm.vortex_params(40, dtype=dt)
        c = vp["case"]
        a = o.vortex_eval(40, 40, c.U0, c.xc, c.yc, c.Rc, c.eps)
        b = c.eval(40, 40)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_sim_plugin_seam(plbm):
    """c_plbm_{init,step,vars,free,norm} against the DDF-shifted oracle of sim/sim.F90."""
    nx, ny, steps, omega = 48, 40, 25, 1.7
    o = Oracle("f64")
    rng = np.random.default_rng(11)
    p = 1e-3 * rng.standard_normal((ny, nx))
    u = 0.05 * rng.standard_normal((2, ny, nx))
    f1 = np.zeros((9, ny + 2, nx + 2))
    f2 = np.zeros_like(f1)
    P = lambda a: a.ctypes.data  # noqa: E731
    o._sim_eqinit(nx, ny, P(f1), P(p), P(u[0]), P(u[1]))
    f2[...] = f1
    for _ in range(steps):
        o._sim_step(nx, ny, P(f1), P(f2), omega)
        o._sim_bc(nx, ny, P(f2))
        f1, f2 = f2, f1
    rho_w, u_w, v_w = np.zeros((ny, nx)), np.zeros((ny, nx)), np.zeros((ny, nx))
    o._sim_macros(nx, ny, P(f1), P(rho_w), P(u_w), P(v_w))

    sim = plbm.SimPlugin()
    sim.init((nx, ny), 1.0, p, u)
    for _ in range(5):
        sim.step(omega)
    sim.step(omega, n=steps - 5)
    rho, uu = sim.vars()
    assert np.array_equal(rho, rho_w) and np.array_equal(uu[0], u_w) and np.array_equal(uu[1], v_w)
    assert abs(sim.norm(uu[0], u_w + 1e-3) - np.sqrt(((uu[0] - u_w - 1e-3) ** 2).sum() / ((u_w + 1e-3) ** 2).sum())) < 1e-12
    sim.free()
    # the same plugin under the reference's own symbol names (libslbm.so drop-in)
    ref_named = plbm.SimPlugin(name="slbm")
    ref_named.init((nx, ny), 1.0, p, u)
    ref_named.step(omega, n=steps)
    rho2, uu2 = ref_named.vars()
    assert np.array_equal(rho2, rho_w) and np.array_equal(uu2[0], u_w)
    ref_named.free()
    with pytest.raises(plbm.PlbmError):
        plbm.SimPlugin().init((nx, ny), 0.5, p, u)  # "Standard LBM only supports dt = 1.0!"


@pytest.mark.parametrize("dt", [1.0, 0.4])
@pytest.mark.parametrize("name,order", [("lw", 2), ("lw4", 4), ("lw6", 6)])
def test_sim_lw_plugin_seam(plbm, dt, name, order):
    """c_lw_* / c_lw4_* / c_lw6_* {init,step,vars,free}: the reference's Lax-Wendroff plugins (sim/sim_lw.F90,
    sim_lw4.F90, sim_lw6.F90) vs their oracle (haloed arrays + periodic halo copies), bit for bit."""
    nx, ny, steps, omega = 40, 56, 12, 1.4
    H = order // 2
    o = Oracle("f64")
    rng = np.random.default_rng(12)
    p = 1e-3 * rng.standard_normal((ny, nx))
    u = 0.05 * rng.standard_normal((2, ny, nx))
    f1 = np.zeros((9, ny + 2 * H, nx + 2 * H))
    f2 = np.zeros_like(f1)
    P = lambda a: a.ctypes.data  # noqa: E731
    if order == 2:
        eqinit = lambda f: o._sim_eqinit(nx, ny, P(f), P(p), P(u[0]), P(u[1]))  # noqa: E731
        stream = lambda a, b: o._lw_stream(nx, ny, P(a), P(b), dt)  # noqa: E731
        collide = lambda f: o._lw_collision(nx, ny, P(f), omega)  # noqa: E731
        bc = lambda f: o._lw_bc(nx, ny, P(f))  # noqa: E731
        macros = lambda f, r, a, b: o._sim_macros(nx, ny, P(f), P(r), P(a), P(b))  # noqa: E731
    else:
        eqinit = lambda f: o._simh_eqinit(nx, ny, H, P(f), P(p), P(u[0]), P(u[1]))  # noqa: E731
        stream = lambda a, b: o._lwh_stream(order, nx, ny, P(a), P(b), dt)  # noqa: E731
        collide = lambda f: o._lwh_collision(nx, ny, H, P(f), omega)  # noqa: E731
        bc = lambda f: o._lwh_bc(nx, ny, H, P(f))  # noqa: E731
        macros = lambda f, r, a, b: o._simh_macros(nx, ny, H, P(f), P(r), P(a), P(b))  # noqa: E731
    eqinit(f1)
    bc(f1)
    for _ in range(steps):
        stream(f1, f2)
        collide(f2)
        bc(f2)
        f1, f2 = f2, f1
    rho_w, u_w, v_w = np.zeros((ny, nx)), np.zeros((ny, nx)), np.zeros((ny, nx))
    macros(f1, rho_w, u_w, v_w)

    sim = plbm.SimPlugin(name=name)
    sim.init((nx, ny), dt, p, u)
    sim.step(omega)
    sim.step(omega, n=steps - 1)
    rho, uu = sim.vars()
    assert np.array_equal(rho, rho_w) and np.array_equal(uu[0], u_w) and np.array_equal(uu[1], v_w)
    sim.free()


@pytest.mark.parametrize("dt", [1.0, 0.3])
def test_sim_fvm_plugin_seam(plbm, dt):
    """c_fvm_{init,step,vars,free}: the reference's Heun finite-volume plugin (sim/sim_fvm.F90) vs its oracle."""
    from test_oracle_properties import _simfvm_run
    nx, ny, steps, omega, H = 40, 56, 9, 1.3, 2
    o = Oracle("f64")
    rng = np.random.default_rng(21)
    p = 1e-3 * rng.standard_normal((ny, nx))
    u = 0.05 * rng.standard_normal((2, ny, nx))
    f1 = _simfvm_run(o, nx, ny, p, u, dt, omega, steps)
    rho_w, u_w, v_w = np.zeros((ny, nx)), np.zeros((ny, nx)), np.zeros((ny, nx))
    P = lambda a: a.ctypes.data  # noqa: E731
    o._simh_macros(nx, ny, H, P(f1), P(rho_w), P(u_w), P(v_w))

    sim = plbm.SimPlugin(name="fvm")
    sim.init((nx, ny), dt, p, u)
    sim.step(omega)
    sim.step(omega, n=steps - 1)
    rho, uu = sim.vars()
    assert np.array_equal(rho, rho_w) and np.array_equal(uu[0], u_w) and np.array_equal(uu[1], v_w)
    sim.free()


def test_output_npy_and_checkpoint_roundtrip(plbm, tmp_path):
    """output_npy writes mf(ny,nx,3) in Fortran order like the reference; a PDF checkpoint restores a run
    bit for bit (continuing from the checkpoint == never stopping)."""
    nx, ny = 48, 40
    og, g = make_pair(plbm, nx, ny, "f64")
    g.collision, g.streaming = plbm.collide_trt, plbm.lbm_stream
    plbm.perform_lbm_step(g, 5)
    plbm.update_macros(g)
    g.filename = "results"
    plbm.set_output_folder(g, str(tmp_path / "out"))
    name = plbm.output_npy(g, step=5)
    assert name.endswith("out/results000000005.npy")
    mf = np.load(name)
    assert mf.shape == (ny, nx, 3) and np.isfortran(mf)
    assert np.array_equal(mf[:, :, 0], g.rho.T) and np.array_equal(mf[:, :, 2], g.uy.T)
    plbm.save_checkpoint(g, str(tmp_path / "ckpt"))
    plbm.perform_lbm_step(g, 7)
    want = g.download_f(g.iold)
    h = plbm.load_checkpoint(str(tmp_path / "ckpt"))
    h.collision, h.streaming = plbm.collide_trt, plbm.lbm_stream
    plbm.perform_lbm_step(h, 7)
    assert (h.iold, h.inew) == (g.iold, g.inew)
    assert np.array_equal(h.download_f(h.iold)[:, :, :ny], want[:, :, :ny])
    plbm.dealloc_grid(g)
    plbm.dealloc_grid(h)


def test_checkpoint_roundtrip_nf3_after_triple_steps(plbm, tmp_path):
    """perform_triple_step rotates (iold, inew, imid) through all three lattices; a checkpoint taken
    after an odd number of such steps must come back with the same roles (plbm_set_indices), so that the resumed run is
    bit-identical to the uninterrupted one and lagged update_macros reads the right lattice."""
    nx, ny = 40, 48
    for nsteps_before in (1, 2, 5):
        og = OracleGrid(nx, ny, "f64")
        f0 = np.nan_to_num(random_state(og.o, nx, ny), nan=0.0)
        g = plbm.alloc_grid(nx, ny, nf=3, precision="f64")
        plbm.set_properties(g, 0.02, 1.0, 0.25)
        g.upload_f(g.iold, f0)
        g.collision, g.streaming = plbm.collide_bgk, plbm.lbm_stream
        plbm.perform_triple_step(g, nsteps_before)
        saved = (g.iold, g.inew, g.imid)
        assert sorted(saved) == [1, 2, 3]
        plbm.save_checkpoint(g, str(tmp_path / f"ckpt3_{nsteps_before}"))
        plbm.update_macros(g)
        rho_lag = g.rho.copy()
        plbm.perform_triple_step(g, 3)
        want = [g.download_f(k) for k in (g.iold, g.inew, g.imid)]
        h = plbm.load_checkpoint(str(tmp_path / f"ckpt3_{nsteps_before}"))
        assert (h.iold, h.inew, h.imid) == saved
        h.collision, h.streaming = plbm.collide_bgk, plbm.lbm_stream
        plbm.update_macros(h)
        assert np.array_equal(h.rho, rho_lag)
        plbm.perform_triple_step(h, 3)
        assert (h.iold, h.inew, h.imid) == (g.iold, g.inew, g.imid)
        for w, k in zip(want, (h.iold, h.inew, h.imid)):
            assert np.array_equal(h.download_f(k)[:, :, :ny], w[:, :, :ny])
        plbm.dealloc_grid(g)
        plbm.dealloc_grid(h)


def test_set_indices_rejects_non_permutations(plbm):
    from periodic_lbm_b200.capi import lib
    g2, g3 = plbm.alloc_grid(8, 8, nf=2), plbm.alloc_grid(8, 8, nf=3)
    assert lib.plbm_set_indices(g2._h, 1, 2, -1) == 0 and (g2.iold, g2.inew, g2.imid) == (1, 2, -1)
    assert lib.plbm_set_indices(g2._h, 1, 1, -1) != 0 and lib.plbm_set_indices(g2._h, 3, 1, 2) != 0
    assert lib.plbm_set_indices(g3._h, 3, 1, 2) == 0 and (g3.iold, g3.inew, g3.imid) == (3, 1, 2)
    assert lib.plbm_set_indices(g3._h, 3, 3, 2) != 0 and lib.plbm_set_indices(g3._h, 0, 1, 2) != 0
    assert (g3.iold, g3.inew, g3.imid) == (3, 1, 2)
    plbm.dealloc_grid(g2)
    plbm.dealloc_grid(g3)


def test_diagnostics_propagate_nan(plbm):
    """a diverged field must not report max|u| = 0 / min|u| = 1e300 (fmax / fmin drop NaN)."""
    g = plbm.alloc_grid(24, 20)
    plbm.set_properties(g, 0.02, 1.0, 0.25)
    g.rho[:], g.ux[:], g.uy[:] = 1.0, 0.01, 0.02
    g.ux[7, 3] = np.nan
    plbm.set_pdf_to_equilibrium(g)
    plbm.update_macros(g, lagged=False)
    d = g.diagnostics()
    assert np.isnan(d["max_speed"]) and np.isnan(d["min_speed"]) and np.isnan(d["sum_rho"]) and np.isnan(d["kinetic_energy"])
    g.rho[:], g.ux[:], g.uy[:] = 1.0, 0.01, 0.02  # update_macros overwrote the host views (rho, uy of that node are NaN too)
    plbm.set_pdf_to_equilibrium(g)
    plbm.update_macros(g, lagged=False)
    d = g.diagnostics()
    assert d["max_speed"] == pytest.approx(np.hypot(0.01, 0.02), rel=1e-12) and d["min_speed"] == pytest.approx(np.hypot(0.01, 0.02), rel=1e-12)
    plbm.dealloc_grid(g)


def test_unfused_dugks_stream_after_a_fused_step(plbm):
    """after a fused perform_dugks_step lattice inew must behave as the reference's fbar+ for an unfused
    dugks_stream too (it used to read ftilde^n and leave the pending half-step collision behind)."""
    nx, ny = 36, 44
    og, g = make_pair(plbm, nx, ny, "f64", dt=0.3)
    p = og.props
    plbm.perform_dugks_step(g, 2)
    og.run(Oracle.SCHEME_DUGKS, Oracle.BGK, 2)
    plbm.dugks_stream(g)          # unfused pass on the lattices as the fused steps left them
    og.o.dugks_stream(og.lattice(og.iold), og.lattice(og.inew), ny, p["tau"], p["dt"], True)
    for k in (1, 2):
        assert np.array_equal(g.download_f(k)[:, :, :ny], og.lattice(k)[:, :, :ny])
    plbm.dealloc_grid(g)


def numpy_lattice_hash(f, ny):
    """restatement of plbm_lattice_hash (csrc/plbm_diag.cu): sum of splitmix64(bits + golden * (i + 1)) mod 2^64 over the
    rows 0..ny-1 of every line, i = (q * nx + x) * ny + y"""
    v = np.ascontiguousarray(f[:, :, :ny]).reshape(-1)
    bits = v.view(np.uint64) if v.dtype == np.float64 else v.view(np.uint32).astype(np.uint64)
    with np.errstate(over="ignore"):
        z = bits + np.uint64(0x9E3779B97F4A7C15) * np.arange(1, v.size + 1, dtype=np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        return int(np.add.reduce(z, dtype=np.uint64))


@pytest.mark.parametrize("prec", PRECS)
def test_lattice_hash_matches_its_numpy_restatement(plbm, prec):
    """plbm_lattice_hash: device-side 64-bit checksum (bench.py's selfcheck compares rings with single-GPU runs through it).
    Equal to the numpy restatement, blind to the padding rows, sensitive to one flipped bit and to a transposition."""
    nx, ny = 37, 53
    og, g = make_pair(plbm, nx, ny, prec)
    f = g.download_f(g.iold)
    h = g.lattice_hash(g.iold)
    assert h == numpy_lattice_hash(f, ny)
    f2 = f.copy()
    f2[:, :, ny:] = 123.0                      # padding rows do not count
    g.upload_f(g.iold, f2)
    assert g.lattice_hash(g.iold) == h
    f3 = f.copy()
    f3[4, 7, 11] = np.nextafter(f3[4, 7, 11], 2)  # one ulp somewhere
    g.upload_f(g.iold, f3)
    assert g.lattice_hash(g.iold) not in (h,) and g.lattice_hash(g.iold) == numpy_lattice_hash(f3, ny)
    f4 = f.copy()
    f4[2, 3, 5], f4[2, 3, 6] = f[2, 3, 6], f[2, 3, 5]  # same values, swapped places
    g.upload_f(g.iold, f4)
    assert g.lattice_hash(g.iold) != h
    plbm.dealloc_grid(g)


def test_error_paths(plbm):
    with pytest.raises(plbm.PlbmError):
        plbm.alloc_grid(0, 8)
    g = plbm.alloc_grid(8, 8)
    g.collision, g.streaming = plbm.collide_bgk, plbm.lbm_stream
    with pytest.raises(plbm.PlbmError):  # set_properties not called
        plbm.perform_lbm_step(g, 1)
    plbm.dealloc_grid(g)
