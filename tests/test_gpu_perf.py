"""Performance guard (GPU): the fused stream+collide kernel must stay at the HBM roofline.

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


@pytest.mark.parametrize("prec,coll", [("f64", "collide_bgk"), ("f64", "collide_trt"), ("f64", "collide_rr"),
                                       ("f32", "collide_bgk"), ("f32", "collide_rr")])
def test_fused_lbm_kernel_is_at_the_hbm_roofline(plbm, prec, coll):
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
        plbm.perform_lbm_step(g, steps)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        best = max(best, n * n * (144 if prec == "f64" else 72) / (ms * 1e-3) / 1e9)
    plbm.dealloc_grid(g)
    assert best / peak_gbs() > 0.85, f"{coll} {prec}: {best:.0f} GB/s = {best / peak_gbs():.2f} of the measured HBM peak"
