"""Deferred stepping (plbm_set_step_deferral): the reference drivers call perform_lbm_step ONCE per time step
(app/main_taylor_green.f90:98-119); with deferral on, the library counts those calls and runs them batched -- two
steps per pass over HBM -- when the budget is reached or anything else touches the grid.  Results, lattice roles
and indices must be exactly those of eager stepping (which tests/test_gpu_parity.py pins to the oracle)."""
import os

import numpy as np
import pytest

from conftest import random_state
from oracle.oracle import Oracle, OracleGrid

pytestmark = pytest.mark.gpu


def make(plbm, nx, ny, prec, f0, defer):
    g = plbm.alloc_grid(nx, ny, precision=prec)
    plbm.set_properties(g, 0.02, 1.0, 0.25)
    g.upload_f(g.iold, f0)
    g.upload_f(g.inew, np.zeros_like(f0))
    g.set_variant(5)  # never the cluster kernel: single steps -> k_lbm, batches -> the two-step kernels
    g.set_step_deferral(defer)
    g.streaming = plbm.lbm_stream
    return g


def state(plbm, g):
    io, inw = g.iold, g.inew  # reported BEFORE anything flushes: must already be the final roles
    plbm.update_macros(g)
    return (io, inw, g.iold, g.inew, g.download_f(g.iold), g.download_f(g.inew), g.rho.copy(), g.ux.copy(), g.uy.copy())


def same(a, b):
    return a[:4] == b[:4] and all(np.array_equal(x, y) for x, y in zip(a[4:], b[4:]))


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("nx,ny,nsteps,defer", [(40, 516, 37, 16), (64, 64, 9, 64), (16, 132, 16, 16), (9, 4, 7, 4)])
def test_deferred_single_steps_equal_eager_single_steps(plbm, nx, ny, nsteps, defer, prec):
    o = Oracle(prec)
    f0 = np.nan_to_num(random_state(o, nx, ny), nan=0.0)
    for coll in (plbm.collide_bgk, plbm.collide_trt, plbm.collide_rr):
        res, launches = [], []
        for d in (0, defer):
            g = make(plbm, nx, ny, prec, f0, d)
            g.collision = coll
            l0 = plbm.launch_count()
            for _ in range(nsteps):
                plbm.perform_lbm_step(g, 1)
            res.append(state(plbm, g))
            launches.append(plbm.launch_count() - l0)
            plbm.dealloc_grid(g)
        assert same(res[0], res[1])
        assert res[0][0:2] == res[0][2:4] and res[1][0:2] == res[1][2:4]  # indices reported before == after the flush
        if ny >= 8:
            assert launches[1] < launches[0], launches  # batched: fewer launches than one per step


def test_deferred_steps_match_the_oracle(plbm):
    nx, ny, nsteps = 40, 516, 21
    og = OracleGrid(nx, ny, "f64")
    og.set_properties(0.02, 1.0, 0.25)
    f0 = random_state(og.o, nx, ny)
    og.lattice(og.iold)[...] = f0
    og.lattice(og.inew)[...] = 0
    og.run(Oracle.SCHEME_LBM, Oracle.TRT, nsteps)
    g = make(plbm, nx, ny, "f64", np.nan_to_num(f0, nan=0.0), 8)
    g.collision = plbm.collide_trt
    for _ in range(nsteps):
        plbm.perform_lbm_step(g, 1)
    assert (g.iold, g.inew) == (og.iold, og.inew)
    assert np.array_equal(g.download_f(g.iold)[:, :, :ny], og.lattice(og.iold)[:, :, :ny])
    assert np.array_equal(g.download_f(g.inew)[:, :, :ny], og.lattice(og.inew)[:, :, :ny])
    plbm.dealloc_grid(g)


def test_changes_between_deferred_steps_take_effect_in_order(plbm):
    """a new omega, a new collision operator, new properties and an explicit batch in the middle of a run"""
    nx, ny = 24, 260
    o = Oracle("f64")
    f0 = np.nan_to_num(random_state(o, nx, ny), nan=0.0)
    res = []
    for d in (0, 32):
        g = make(plbm, nx, ny, "f64", f0, d)
        g.collision = plbm.collide_bgk
        for _ in range(5):
            plbm.perform_lbm_step(g, 1)
        g.omega = 1.25                      # pending steps must run with the old rate
        for _ in range(4):
            plbm.perform_lbm_step(g, 1)
        g.omega = 1.25                      # unchanged value (what the Fortran shim sends before every step): no flush needed
        plbm.perform_lbm_step(g, 2)
        g.collision = plbm.collide_rr       # operator switch
        for _ in range(3):
            plbm.perform_lbm_step(g, 1)
        plbm.set_properties(g, 0.05, 1.0, 0.25)
        plbm.perform_lbm_step(g, 1)
        plbm.perform_lbm_step(g, 40)        # >= the budget: runs at once, after what was pending
        plbm.perform_lbm_step(g, 1)
        st = state(plbm, g)                 # update_macros is an observer: it sees all 56 steps
        d0 = g.diagnostics()
        res.append((st, d0["sum_rho"], d0["kinetic_energy"]))
        plbm.dealloc_grid(g)
    assert same(res[0][0], res[1][0]) and res[0][1:] == res[1][1:]


def test_pair_kernel_selection(plbm):
    """plbm_lbm_pair_kernel: which kernel a call of >= 3 steps uses (bench accounting).  Grids on which the
    launcher of k_lbm2_bulk can fill one round of blocks get the bulk-copy flavour, smaller ones k_lbm2; the variants force either."""
    if os.environ.get("PLBM_PAIR_BULK", "") not in ("", "1"):
        pytest.skip("PLBM_PAIR_BULK overrides the default selection")
    for shape, prec, variant, want in (((64, 64), "f64", 0, "k_lbm2"), ((64, 64), "f64", 7, "k_lbm2_bulk"), ((64, 64), "f64", 6, "k_lbm2"),
                                       ((64, 64), "f64", 1, "k_lbm"), ((64, 8), "f64", 7, "k_lbm2"), ((64, 16), "f32", 7, "k_lbm2"),
                                       ((64, 67), "f64", 0, "k_lbm"), ((256, 256), "f64", 0, "k_lbm2"), ((1024, 1024), "f64", 0, "k_lbm2_bulk"),
                                       ((4096, 4096), "f64", 0, "k_lbm2_bulk"), ((4096, 4096), "f32", 0, "k_lbm2_bulk"),
                                       ((8192, 8192), "f32", 0, "k_lbm2_bulk")):
        g = plbm.alloc_grid(*shape, precision=prec)
        g.set_variant(variant)
        assert g.pair_kernel() == want, (shape, prec, variant, g.pair_kernel())
        # plbm_lbm_steps_per_pass: three steps per pass (k_lbmn_bulk) by default from 512^2 nodes, for bgk / trt / rr
        spp = g.steps_per_pass(plbm.collide_trt)
        if os.environ.get("PLBM_TRIPLES", "") in ("", "1"):
            assert spp == (3 if variant == 0 and shape[0] * shape[1] >= 512 * 512 else (1 if want == "k_lbm" else 2)), (shape, prec, variant, spp)
        plbm.dealloc_grid(g)


@pytest.mark.skipif(os.environ.get("PLBM_TEST_EXPERIMENTAL", "0") == "0",
                    reason="FMA-contracted tile kernels (csrc/plbm_fvm_tma_fma.cu, variant 3): set PLBM_TEST_EXPERIMENTAL=1")
@pytest.mark.parametrize("prec,rtol", [("f64", 1e-12), ("f32", 1e-5)])
@pytest.mark.parametrize("scheme", ["dugks", "fvm", "fdm_bardow"])
def test_fma_contracted_tile_kernels_stay_within_the_stated_tolerance(plbm, scheme, prec, rtol):
    """variant 3 = the TMA-pipelined tile kernels compiled with -fmad=true: NOT bit-identical, but within
    BASELINE.json's tolerance (1e-12 relative fp64, 1e-5 fp32) of the oracle after N steps, PDFs and rho/u."""
    nx = ny = 130
    nsteps = 10
    og = OracleGrid(nx, ny, prec)
    og.set_properties(0.02, 0.3, 0.25)
    f0 = random_state(og.o, nx, ny)
    og.lattice(og.iold)[...] = f0
    og.lattice(og.inew)[...] = 0
    g = plbm.alloc_grid(nx, ny, precision=prec)
    plbm.set_properties(g, 0.02, 0.3, 0.25)
    g.upload_f(g.iold, np.nan_to_num(f0, nan=0.0))
    g.upload_f(g.inew, np.zeros_like(f0))
    g.set_variant(3)
    if scheme == "dugks":
        og.run(Oracle.SCHEME_DUGKS, Oracle.BGK, nsteps)
        plbm.perform_dugks_step(g, nsteps)
    else:
        g.collision = plbm.collide_bgk
        g.streaming = plbm.stream_fvm_bardow if scheme == "fvm" else plbm.stream_fdm_bardow
        og.run(Oracle.SCHEME_FVM_BARDOW if scheme == "fvm" else Oracle.SCHEME_FDM_BARDOW, Oracle.BGK, nsteps)
        plbm.perform_step(g, nsteps)
    got = g.download_f(g.iold)[:, :, :ny].astype(np.float64)
    want = og.lattice(og.iold)[:, :, :ny].astype(np.float64)
    assert np.abs(got - want).max() <= rtol * np.abs(want).max()
    plbm.update_macros(g, lagged=False)
    r, u, v = og.update_macros(lagged=False)
    assert np.abs(g.rho - r).max() <= rtol * np.abs(r).max()
    assert np.abs(g.ux - u).max() <= rtol * max(np.abs(u).max(), np.abs(v).max())
    assert np.abs(g.uy - v).max() <= rtol * max(np.abs(u).max(), np.abs(v).max())
    plbm.dealloc_grid(g)


@pytest.mark.skipif(os.environ.get("PLBM_TEST_EXPERIMENTAL", "0") == "0",
                    reason="FMA-contracted two-step LBM kernels (csrc/plbm_lbm2_fma.cu, variant 11): set PLBM_TEST_EXPERIMENTAL=1")
@pytest.mark.parametrize("prec,rtol", [("f64", 1e-12), ("f32", 1e-5)])
@pytest.mark.parametrize("nx,ny", [(40, 516), (16, 24)])
def test_fma_contracted_two_step_kernels_stay_within_the_stated_tolerance(plbm, nx, ny, prec, rtol):
    """variant 11 = k_lbm2_bulk / k_lbm2 compiled with -fmad=true: within BASELINE.json's tolerance of the oracle
    after N steps (PDFs and rho/u), not bit-identical."""
    nsteps = 21
    for coll, ocoll in ((plbm.collide_bgk, Oracle.BGK), (plbm.collide_trt, Oracle.TRT), (plbm.collide_rr, Oracle.RR)):
        og = OracleGrid(nx, ny, prec)
        og.set_properties(0.02, 1.0, 0.25)
        f0 = random_state(og.o, nx, ny)
        og.lattice(og.iold)[...] = f0
        og.lattice(og.inew)[...] = 0
        og.run(Oracle.SCHEME_LBM, ocoll, nsteps)
        g = make(plbm, nx, ny, prec, np.nan_to_num(f0, nan=0.0), 0)
        g.set_variant(11)
        g.collision = coll
        plbm.perform_lbm_step(g, nsteps)
        assert (g.iold, g.inew) == (og.iold, og.inew)
        got = g.download_f(g.iold)[:, :, :ny].astype(np.float64)
        want = og.lattice(og.iold)[:, :, :ny].astype(np.float64)
        assert np.abs(got - want).max() <= rtol * np.abs(want).max()
        plbm.update_macros(g, lagged=False)
        r, u, v = og.update_macros(lagged=False)
        assert np.abs(g.rho - r).max() <= rtol * np.abs(r).max()
        assert np.abs(g.ux - u).max() <= rtol * max(np.abs(u).max(), np.abs(v).max())
        plbm.dealloc_grid(g)
