"""Pin the CPU oracle against the only golden data the reference ships for this path:
graphs/fvm_bardow_64.txt and graphs/fvm_dugks_64.txt (copied verbatim to tests/golden/ref_*.txt).
Recipe recovered in SURVEY.md App. B: TG 64^2, Re=100, umax=0.01/sqrt(3), dt=(dt/tau)*tau, stop at
first t >= ln2*td, L2 from the one-step-lagged update_macros, -DDUGKS for the DUGKS file."""
import os

import numpy as np
import pytest

from oracle.oracle import Oracle, OracleGrid, taylor_green_l2_run, taylor_green_setup

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load_rows(name):
    rows = np.loadtxt(os.path.join(GOLD, name))
    return {float(r): float(v) for r, v in rows}


# the CPU suite checks the cheap rows (large dt/tau = few steps); tests/test_gpu_golden.py checks more
@pytest.mark.parametrize("r", [50.0, 40.0, 30.0])
def test_oracle_reproduces_fvm_bardow_golden(r):
    gold = load_rows("ref_fvm_bardow_64.txt")[r]
    l2, steps, t, _ = taylor_green_l2_run(64, Oracle.SCHEME_FVM_BARDOW, Oracle.BGK, dt_over_tau=r, omp=True)
    assert abs(l2 - gold) / gold < 1e-7, (l2, gold)  # 8 printed digits


@pytest.mark.parametrize("r", [70.0, 60.0, 50.0])
def test_oracle_reproduces_fvm_dugks_golden(r):
    gold = load_rows("ref_fvm_dugks_64.txt")[r]
    l2, steps, t, _ = taylor_green_l2_run(64, Oracle.SCHEME_DUGKS, Oracle.BGK, dt_over_tau=r, omp=True)
    assert abs(l2 - gold) / gold < 1e-7, (l2, gold)


def test_lag_semantics_matter():
    """Only the lagged update_macros reproduces the golden file."""
    gold = load_rows("ref_fvm_bardow_64.txt")[50.0]
    l2, *_ = taylor_green_l2_run(64, Oracle.SCHEME_FVM_BARDOW, Oracle.BGK, dt_over_tau=50.0, lagged=False, omp=True)
    assert abs(l2 - gold) / gold > 1e-4


def test_dugks_without_macro_degenerates_to_bardow():
    """periodic_dugks built without -DDUGKS is Bardow's scheme."""
    a, *_ = taylor_green_l2_run(64, Oracle.SCHEME_DUGKS_OFF, Oracle.BGK, dt_over_tau=50.0, omp=True)
    b, *_ = taylor_green_l2_run(64, Oracle.SCHEME_FVM_BARDOW, Oracle.BGK, dt_over_tau=50.0, omp=True)
    assert abs(a - b) / b < 1e-9


def test_serial_and_openmp_builds_are_bit_identical():
    o = Oracle("f64")
    oo = Oracle("f64", omp=True)
    rng = np.random.default_rng(1)
    nx, ny = 37, 53
    f = o.alloc_f(nx, ny, fill=0.0)
    f[:, :, :ny] = 0.1 + 0.01 * rng.random((9, nx, ny))
    for name, args in [("collide_bgk", (1.3,)), ("collide_rr", (1.3,)), ("collide_trt", (1.3, 0.25)), ("kernel_bgk", (0.7,))]:
        a, b = f.copy(), f.copy()
        getattr(o, name)(a, ny, *args)
        getattr(oo, name)(b, ny, *args)
        assert np.array_equal(a, b), name


@pytest.mark.parametrize("collision,expect", [
    (Oracle.BGK, 1.525202021036551e-3), (Oracle.TRT, 9.993568347430525e-4), (Oracle.RR, 1.5119419007500739e-3)])
def test_lbm_tg64_matches_survey_expected_values(collision, expect):
    """SURVEY App. B.2 (not pinned by the reference: independent numpy restatement at survey time)."""
    l2, steps, t, g = taylor_green_l2_run(64, Oracle.SCHEME_LBM, collision, dt=1.0, omp=True)
    assert steps == 9732
    assert abs(float(g.props["omega"]) - 1.9566212177872033) < 1e-15
    assert abs(l2 - expect) / expect < 1e-10


@pytest.mark.parametrize("scheme,collision,kw,expect", [
    (Oracle.SCHEME_LBM, Oracle.BGK, dict(dt=1.0), 0.11020184847034406),
    (Oracle.SCHEME_LBM, Oracle.TRT, dict(dt=1.0), 0.11020497419215258),
    (Oracle.SCHEME_LBM, Oracle.RR, dict(dt=1.0), 0.11020205424303382),
    (Oracle.SCHEME_FVM_BARDOW, Oracle.BGK, dict(dt_over_tau=5.0), 0.11017918190099109),
    (Oracle.SCHEME_DUGKS, Oracle.BGK, dict(dt_over_tau=5.0), 0.11017737500673701)])
def test_single_element_anchor_after_100_steps(scheme, collision, kw, expect):
    """f(y=6, x=4, q=1) (1-based) of the current lattice after exactly 100 steps."""
    g = OracleGrid(64, 64)
    o = g.o
    s = taylor_green_setup(o, 64, **kw)
    g.set_properties(s["nu"], s["dt"], magic=0.25)
    p, ux, uy = o.taylor_green_eval(64, 64, s["kx"], s["ky"], s["umax"], s["td"], 0.0)
    g.rho, g.ux, g.uy = p / g.props["csqr"] + 1.0, ux, uy
    g.set_pdf_to_equilibrium()
    g.run(scheme, collision, 100)
    val = g.lattice(g.iold)[1, 3, 5]
    assert abs(val - expect) < 2e-15, (val, expect)


def test_fp32_oracle_tracks_fp64():
    a, *_ = taylor_green_l2_run(32, Oracle.SCHEME_LBM, Oracle.BGK, dt=1.0, precision="f64")
    b, *_ = taylor_green_l2_run(32, Oracle.SCHEME_LBM, Oracle.BGK, dt=1.0, precision="f32")
    assert abs(a - b) / a < 5e-2


def test_mass_conservation_and_moments():
    g = OracleGrid(48, 48)
    o = g.o
    s = taylor_green_setup(o, 48, dt=1.0)
    g.set_properties(s["nu"], s["dt"], magic=0.25)
    p, ux, uy = o.taylor_green_eval(48, 48, s["kx"], s["ky"], s["umax"], s["td"], 0.0)
    g.rho, g.ux, g.uy = p / g.props["csqr"] + 1.0, ux, uy
    m0 = g.rho.sum()
    g.set_pdf_to_equilibrium()
    for coll in (Oracle.BGK, Oracle.TRT, Oracle.RR):
        g.run(Oracle.SCHEME_LBM, coll, 50)
        rho, _, _ = g.update_macros(lagged=False)
        assert abs(rho.sum() - m0) / m0 < 1e-13
