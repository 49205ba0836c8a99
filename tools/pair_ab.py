"""A/B of the two flavours of the two-step kernel (development tool, no torch: starts in a second).

variant 6 = raw columns by per-thread loads (k_lbm2), 7 = by bulk async copies (k_lbm2_bulk); 9 / 10 = the experimental
depth-generic kernel k_lbmn_bulk with two / three steps per pass (bgk, trt, rr).  Timed by the
host clock around one perform_lbm_step(K) call with a device synchronize on both sides (K - 1 steps in pair
launches + the closing single step), best of `reps`; one JSON line per case.  `--once` runs a single short
call per case instead (for `ncu`)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import periodic_lbm_b200 as p  # noqa: E402

try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6540.8


def run(nx, ny, prec, coll, variant, steps, reps, once=False):
    g = p.alloc_grid(nx, ny, precision=prec)
    p.set_properties(g, 0.05, 1.0, 0.25)
    g.rho[:], g.ux[:], g.uy[:] = 1.0, 0.01, -0.02
    p.set_pdf_to_equilibrium(g)
    g.set_variant(variant)
    g.collision, g.streaming = getattr(p, "collide_" + coll), p.lbm_stream
    kernel = g.pair_kernel()  # what variants 5..8 use (9 / 10: k_lbmn_bulk, 11: the FMA build of this one)
    if variant in (9, 10) or (variant == 0 and g.steps_per_pass() == 3):
        kernel = g.triple_kernel() if variant == 0 else "k_lbmn_bulk"
    if once:
        p.perform_lbm_step(g, 4)  # one pair + two single steps; variant 10: one triple + one single step
        g.synchronize()
        p.dealloc_grid(g)
        return dict(nx=nx, ny=ny, prec=prec, coll=coll, variant=variant, once=True)
    p.perform_lbm_step(g, 5)
    g.synchronize()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        p.perform_lbm_step(g, steps)
        g.synchronize()
        best = min(best, (time.perf_counter() - t0) / steps)
    p.dealloc_grid(g)
    mlups = nx * ny / best * 1e-6
    gbs = mlups * 1e6 * (144 if prec == "f64" else 72) / 1e9
    knobs = {k: v for k, v in os.environ.items() if k.startswith("PLBM_")}
    return dict(nx=nx, ny=ny, prec=prec, coll=coll, variant=variant, kernel=kernel, env=knobs, steps=steps, ms_per_step=round(best * 1e3, 4),
                mlups=round(mlups, 1), algorithmic_gbs=round(gbs, 1), frac_of_hbm_peak=round(gbs / PEAK, 4))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="8192x8192:f64:bgk,8192x8192:f64:trt,8192x8192:f64:rr,8192x8192:f32:bgk,8192x8192:f32:rr",
                    help="comma list of NXxNY:prec:collision")
    ap.add_argument("--variants", default="6,7")
    ap.add_argument("--steps", type=int, default=41)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--once", action="store_true")
    a = ap.parse_args()
    for case in a.cases.split(","):
        shape, prec, coll = case.split(":")
        nx, ny = (int(v) for v in shape.split("x"))
        for variant in (int(v) for v in a.variants.split(",")):
            print(json.dumps(run(nx, ny, prec, coll, variant, a.steps, a.reps, a.once)), flush=True)
