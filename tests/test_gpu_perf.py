"""Performance guard (GPU): the fused stream+collide kernels must stay at / above the HBM roofline.

A register-count regression once dropped the RR fp64 kernel from 3 to 2 resident blocks per SM
(-24 %) without any test noticing; this test would have.  Thresholds are far below the measured
values (1.03-1.05 of the measured copy bandwidth for fp64, 0.97-1.05 for fp32)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6540.8


# floor of the algorithmic GB/s (144 B fp64 / 72 B fp32 per update) over the measured HBM peak:
#   one step per launch (k_lbm) measured 1.03-1.05 fp64, 0.96-1.05 fp32;
#   two steps per pass, raw columns by bulk async copies (k_lbm2_bulk) measured 1.83 / 1.82 / 1.51 (BGK / TRT / RR fp64),
#   1.62 / 1.10 (BGK / RR fp32); by per-thread loads (k_lbm2) 1.41 / 1.41 / 1.29, 1.31 / 1.08
CASES = [("f64", "collide_bgk", 1.50), ("f64", "collide_trt", 1.50), ("f64", "collide_rr", 1.25),
         ("f32", "collide_bgk", 1.35), ("f32", "collide_rr", 0.95)]


@pytest.mark.parametrize("two_step", [False, True])
@pytest.mark.parametrize("prec,coll,floor2", CASES)
def test_fused_lbm_kernel_is_at_the_hbm_roofline(plbm, prec, coll, floor2, two_step):
    import torch

    n, steps = 8192, 60
    stream = torch.cuda.Stream()
    g = plbm.alloc_grid(n, n, precision=prec)
    g.set_stream(stream.cuda_stream)
    plbm.set_properties(g, 0.05, 1.0, 0.25)
    g.rho[:], g.ux[:], g.uy[:] = 1.0, 0.01, -0.02
    plbm.set_pdf_to_equilibrium(g)
    g.collision, g.streaming = getattr(plbm, coll), plbm.lbm_stream
    plbm.perform_lbm_step(g, 5)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        if two_step:
            plbm.perform_lbm_step(g, steps)  # (steps-1)//2 launches of k_lbm2 + the closing k_lbm
        else:
            for _ in range(steps):           # one k_lbm launch per call
                plbm.perform_lbm_step(g, 1)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        best = max(best, n * n * (144 if prec == "f64" else 72) / (ms * 1e-3) / 1e9)
    plbm.dealloc_grid(g)
    floor = floor2 if two_step else 0.85
    assert best / peak_gbs() > floor, f"{coll} {prec} two_step={two_step}: {best:.0f} GB/s = {best / peak_gbs():.2f} of the measured HBM peak"
