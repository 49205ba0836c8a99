"""Kernel micro-benchmark (development tool): MLUPS and achieved GB/s of the step kernels per
variant / collision / precision, timed with CUDA events on the launching stream."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import periodic_lbm_b200 as p  # noqa: E402

PEAK = 6540.8
STREAM = torch.cuda.Stream()  # a real (non-default) stream: handle 0 would mean 'library-owned stream'
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def run(nx, ny, prec, scheme, coll, variant, steps, warmup=3, dugks=True):
    g = p.alloc_grid(nx, ny, precision=prec)
    g.set_stream(STREAM.cuda_stream)
    tp = p.taylor_green_params(nx, dt=1.0, dtype=g.dtype)
    if scheme == "lbm":
        p.set_properties(g, tp["nu"], 1.0, 0.25)
    else:
        p.set_properties(g, tp["nu"], min(5.0 * float(tp["tau"]), 0.5), 0.25)  # dt = 5 tau capped at CFL 0.5 (stable)
    # cheap IC: uniform rho with a small shear (values do not matter for bandwidth)
    g.rho[:] = 1.0
    g.ux[:] = 0.01
    g.uy[:] = -0.02
    p.set_pdf_to_equilibrium(g)
    g.set_variant(variant)
    g.dugks = dugks
    knobs = {k: v for k, v in os.environ.items() if k.startswith("PLBM_")}
    g.collision = {"bgk": p.collide_bgk, "trt": p.collide_trt, "rr": p.collide_rr, "split": p.collide_bgk_split}[coll]
    if scheme == "lbm":
        g.streaming = p.lbm_stream
        step = lambda n: p.perform_lbm_step(g, n)  # noqa: E731
    elif scheme in ("fvm", "fdmb", "fdms"):
        g.streaming = {"fvm": p.stream_fvm_bardow, "fdmb": p.stream_fdm_bardow, "fdms": p.stream_fdm_sofonea}[scheme]
        step = lambda n: p.perform_step(g, n)  # noqa: E731
    else:
        g.collision = g.streaming = None
        step = lambda n: p.perform_dugks_step(g, n)  # noqa: E731
    step(warmup)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(STREAM)
    step(steps)
    e1.record(STREAM)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    mlups = nx * ny / ms * 1e-3
    bpl = 144 if prec == "f64" else 72
    gbs = mlups * 1e6 * bpl / 1e9
    p.dealloc_grid(g)
    return dict(nx=nx, ny=ny, prec=prec, scheme=scheme, coll=coll, variant=variant, env=knobs, ms=round(ms, 4), mlups=round(mlups, 1),
                gbs=round(gbs, 1), frac=round(gbs / PEAK, 4))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--what", default="lbm")
    ap.add_argument("--case", default="", help="scheme,prec,coll,variant  e.g. dugks,f64,bgk,0 (runs only this)")
    a = ap.parse_args()
    n = a.n
    rows = []
    if a.case:
        scheme, prec, coll, variant = a.case.split(",")
        print(json.dumps(run(n, n, prec, scheme, coll, int(variant), a.steps)), flush=True)
        sys.exit(0)
    if "lbm" in a.what:
        for prec in ("f64", "f32"):
            for coll in ("bgk", "trt", "rr"):
                for variant in (0, 1, 2, 3):
                    rows.append(run(n, n, prec, "lbm", coll, variant, a.steps))
                    print(json.dumps(rows[-1]), flush=True)
    if "fvm" in a.what:
        for prec in ("f64", "f32"):
            rows.append(run(n // 2, n // 2, prec, "fvm", "bgk", 0, a.steps))
            print(json.dumps(rows[-1]), flush=True)
            for variant in (0, 1):
                rows.append(run(n // 2, n // 2, prec, "dugks", "bgk", variant, a.steps))
                print(json.dumps(rows[-1]), flush=True)
