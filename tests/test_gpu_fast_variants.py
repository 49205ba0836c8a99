"""OPT-IN kernels that trade bit-identity for speed, gated at the tolerance the north-star states: PDFs and rho/u within
1e-12 relative (fp64) / 1e-5 (fp32) of the reference restatement after N steps, and the reference's golden Taylor-Green
sweep reproduced to its printed digits.

  variant 4   k_fv_march (csrc/plbm_fvm_march.cu): DUGKS / Bardow-FVM with every cell face reconstructed and relaxed ONCE
              and shared by its two cells (the reference evaluates each face twice, in two summation orders), FMA contraction
  variant 11  the two-step LBM kernels compiled with FMA contraction (csrc/plbm_lbm2_fma.cu)

The DEFAULT kernels stay bit-identical to the oracle (tests/test_gpu_parity.py)."""
import numpy as np
import pytest

from conftest import random_state
from oracle.oracle import Oracle, OracleGrid
from test_gpu_golden import assert_matches_golden, gpu_tg_run, load_rows

pytestmark = pytest.mark.gpu
RTOL = {"f64": 1e-12, "f32": 1e-5}


def _pair(plbm, nx, ny, prec, dt, variant):
    og = OracleGrid(nx, ny, prec)
    og.set_properties(0.02, dt, 0.25)
    f0 = random_state(og.o, nx, ny)
    og.lattice(og.iold)[...] = f0
    og.lattice(og.inew)[...] = 0
    g = plbm.alloc_grid(nx, ny, precision=prec)
    plbm.set_properties(g, 0.02, dt, 0.25)
    g.upload_f(g.iold, np.nan_to_num(f0, nan=0.0))
    g.set_variant(variant)
    return og, g


def _assert_close(got, want, ny, prec, what):
    got, want = got[:, :, :ny].astype(np.float64), want[:, :, :ny].astype(np.float64)
    rel = np.abs(got - want).max() / np.abs(want).max()
    elem = (np.abs(got - want) / np.abs(want)).max()   # PDFs are positive and O(w_q): element-wise relative is meaningful
    assert rel <= RTOL[prec] and elem <= 20 * RTOL[prec], (what, rel, elem)
    return rel


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("dugks", [True, False])
@pytest.mark.parametrize("nx,ny", [(64, 64), (67, 53), (16, 130), (5, 3), (300, 260), (3, 515)])
def test_marching_dugks_within_tolerance_of_the_oracle(plbm, nx, ny, dugks, prec):
    steps = 10
    og, g = _pair(plbm, nx, ny, prec, 0.3, 4)
    g.dugks = dugks
    plbm.perform_dugks_step(g, steps)
    og.run(Oracle.SCHEME_DUGKS if dugks else Oracle.SCHEME_DUGKS_OFF, Oracle.BGK, steps)
    assert (g.iold, g.inew) == (og.iold, og.inew)
    _assert_close(g.download_f(g.iold), og.lattice(og.iold), ny, prec, "ftilde^n")
    # lattice inew: fbar+ of the last step, materialised lazily by the bit-exact collision from ftilde^{n-1} (within tolerance too)
    _assert_close(g.download_f(g.inew), og.lattice(og.inew), ny, prec, "fbar+")
    plbm.update_macros(g, lagged=False)
    og.update_macros(lagged=False)
    for a, b in ((g.rho, og.rho), (g.ux, og.ux), (g.uy, og.uy)):
        assert np.abs(a.astype(np.float64) - b).max() <= RTOL[prec] * max(np.abs(b).max(), 1e-300) * 10
    plbm.dealloc_grid(g)


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("coll", ["bgk", "trt", "rr"])
@pytest.mark.parametrize("nx,ny", [(64, 64), (67, 53), (16, 130)])
def test_marching_bardow_fvm_within_tolerance_of_the_oracle(plbm, nx, ny, coll, prec):
    steps = 10
    og, g = _pair(plbm, nx, ny, prec, 0.3, 4)
    g.collision, g.streaming = getattr(plbm, "collide_" + coll), plbm.stream_fvm_bardow
    plbm.perform_step(g, steps)
    og.run(Oracle.SCHEME_FVM_BARDOW, {"bgk": Oracle.BGK, "trt": Oracle.TRT, "rr": Oracle.RR}[coll], steps)
    _assert_close(g.download_f(g.iold), og.lattice(og.iold), ny, prec, "f^n")
    plbm.dealloc_grid(g)


def test_marching_kernel_is_deterministic_and_differs_from_the_default_only_in_last_bits(plbm):
    nx, ny, steps = 96, 200, 8
    res = []
    for variant in (4, 4, 0):
        og, g = _pair(plbm, nx, ny, "f64", 0.3, variant)
        plbm.perform_dugks_step(g, steps)
        res.append(g.download_f(g.iold)[:, :, :ny])
        plbm.dealloc_grid(g)
    assert np.array_equal(res[0], res[1])
    assert not np.array_equal(res[0], res[2])  # it IS a different summation order: if this ever fails, make it the default
    assert np.abs(res[0] - res[2]).max() / np.abs(res[2]).max() < 1e-14


@pytest.mark.parametrize("r", [50.0, 20.0, 10.0])
def test_marching_dugks_reproduces_the_golden_sweep(plbm, r, monkeypatch):
    """graphs/fvm_dugks_64.txt through the fast kernel: same printed digits (up to 87,789 steps)."""
    gold = load_rows("ref_fvm_dugks_64.txt")[r]
    orig = plbm.alloc_grid

    def alloc(*a, **k):
        g = orig(*a, **k)
        g.set_variant(4)
        return g
    monkeypatch.setattr(plbm, "alloc_grid", alloc)
    l2, steps, t, g, _, _ = gpu_tg_run(plbm, 64, "dugks", dt_over_tau=r)
    plbm.dealloc_grid(g)
    assert_matches_golden(l2, gold, steps)


@pytest.mark.parametrize("r", [50.0, 20.0])
def test_marching_bardow_reproduces_the_golden_sweep(plbm, r, monkeypatch):
    gold = load_rows("ref_fvm_bardow_64.txt")[r]
    orig = plbm.alloc_grid

    def alloc(*a, **k):
        g = orig(*a, **k)
        g.set_variant(4)
        return g
    monkeypatch.setattr(plbm, "alloc_grid", alloc)
    l2, steps, t, g, _, _ = gpu_tg_run(plbm, 64, "fvm", plbm.collide_bgk, dt_over_tau=r)
    plbm.dealloc_grid(g)
    assert_matches_golden(l2, gold, steps)


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("coll", ["bgk", "trt", "rr"])
def test_fma_two_step_lbm_within_tolerance_of_the_oracle(plbm, coll, prec):
    """variant 11: k_lbm2 / k_lbm2_bulk with their multiply-adds contracted (RR fp64 +13 %, RR fp32 +16 % at 8192^2)."""
    nx, ny, steps = 96, 640, 21
    og, g = _pair(plbm, nx, ny, prec, 1.0, 11)
    g.collision, g.streaming = getattr(plbm, "collide_" + coll), plbm.lbm_stream
    plbm.perform_lbm_step(g, steps)
    og.run(Oracle.SCHEME_LBM, {"bgk": Oracle.BGK, "trt": Oracle.TRT, "rr": Oracle.RR}[coll], steps)
    _assert_close(g.download_f(g.iold), og.lattice(og.iold), ny, prec, "f^n")
    plbm.dealloc_grid(g)
