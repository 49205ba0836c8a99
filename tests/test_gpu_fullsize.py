"""Parity at BASELINE.json's FULL sizes through a size-independent property: periodic tiling.

If the initial state of an N x N periodic grid is a T x T state repeated (N/T)^2 times, every node of
the big grid sees exactly the neighbourhood its image in the small grid sees, so after any number of
steps the big solution is the small solution repeated -- bit for bit (same arithmetic per node).  The
small run is itself checked against the CPU oracle here.  This exercises 64-bit indexing, the periodic
wrap at the far edges, the vector/tile paths and (for 32768^2) a lattice that fills the GPU."""
import numpy as np
import pytest

from conftest import random_state
from oracle.oracle import Oracle, OracleGrid

pytestmark = pytest.mark.gpu
T = 64


def small_state(prec):
    o = Oracle(prec)
    rng = np.random.default_rng(314)
    rho = (0.95 + 0.1 * rng.random((T, T))).astype(o.dtype)
    ux = (0.05 * (rng.random((T, T)) - 0.5)).astype(o.dtype)
    uy = (0.05 * (rng.random((T, T)) - 0.5)).astype(o.dtype)
    return rho, ux, uy


def run(plbm, n, prec, scheme, coll, steps, rho, ux, uy, nu, dt):
    g = plbm.alloc_grid(n, n, precision=prec)
    plbm.set_properties(g, nu, dt, 0.25)
    reps = n // T
    for dst, src in ((g.rho, rho), (g.ux, ux), (g.uy, uy)):
        dst.reshape(reps, T, reps, T)[...] = src[None, :, None, :]
    plbm.set_pdf_to_equilibrium(g)
    if scheme == "lbm":
        g.collision, g.streaming = getattr(plbm, coll), plbm.lbm_stream
        plbm.perform_lbm_step(g, steps)
    elif scheme == "dugks":
        plbm.perform_dugks_step(g, steps)
    else:
        g.collision, g.streaming = getattr(plbm, coll), plbm.stream_fvm_bardow
        plbm.perform_step(g, steps)
    plbm.update_macros(g, lagged=False)
    return g


CASES = [
    # BASELINE config, n, precision, scheme, collision, steps
    ("C2", 1024, "f64", "lbm", "collide_trt", 9),
    ("C3", 8192, "f64", "lbm", "collide_rr", 5),
    ("C3", 8192, "f32", "lbm", "collide_rr", 5),
    ("C4", 2048, "f64", "dugks", None, 5),
    ("C4", 2048, "f32", "dugks", None, 5),
    ("f1", 2048, "f64", "fvm", "collide_bgk", 5),
    ("C5", 32768, "f64", "lbm", "collide_bgk", 3),
]


@pytest.mark.parametrize("cfg,n,prec,scheme,coll,steps", CASES)
def test_big_grid_equals_tiled_small_grid(plbm, cfg, n, prec, scheme, coll, steps):
    if n == 32768:
        import torch

        free, total = torch.cuda.mem_get_info()
        need = (2 * 9 + 3) * n * n * 8 + (1 << 30)
        if free < need:
            pytest.skip(f"needs {need / 1e9:.0f} GB of free GPU memory, have {free / 1e9:.0f}")
    nu, dt = 0.02, (1.0 if scheme == "lbm" else 0.3)
    rho, ux, uy = small_state(prec)

    # small grid on the device ...
    gs = run(plbm, T, prec, scheme, coll, steps, rho, ux, uy, nu, dt)
    small = (gs.rho.copy(), gs.ux.copy(), gs.uy.copy())
    plbm.dealloc_grid(gs)
    # ... which the oracle confirms
    og = OracleGrid(T, T, prec)
    og.set_properties(nu, dt, 0.25)
    og.rho, og.ux, og.uy = rho, ux, uy
    og.set_pdf_to_equilibrium()
    osch = {"lbm": Oracle.SCHEME_LBM, "dugks": Oracle.SCHEME_DUGKS, "fvm": Oracle.SCHEME_FVM_BARDOW}[scheme]
    ocoll = {"collide_trt": Oracle.TRT, "collide_rr": Oracle.RR, "collide_bgk": Oracle.BGK, None: Oracle.BGK}[coll]
    og.run(osch, ocoll, steps)
    want = og.update_macros(lagged=False)
    assert all(np.array_equal(a, b) for a, b in zip(small, want))

    # big grid == the small one repeated
    gb = run(plbm, n, prec, scheme, coll, steps, rho, ux, uy, nu, dt)
    reps = n // T
    for big, sm in zip((gb.rho, gb.ux, gb.uy), small):
        v = big.reshape(reps, T, reps, T)
        for i in range(0, reps, 64):  # blockwise to bound the temporaries
            assert np.array_equal(v[i:i + 64], np.broadcast_to(sm[None, :, None, :], v[i:i + 64].shape)), f"{cfg}: tile mismatch"
    plbm.dealloc_grid(gb)
