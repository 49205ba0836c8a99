"""Run every kernel family once on awkward grid sizes -- meant to be executed under compute-sanitizer:
   compute-sanitizer --tool memcheck  python tools/sanitize.py
   compute-sanitizer --tool racecheck python tools/sanitize.py
   compute-sanitizer --tool initcheck python tools/sanitize.py   (padding rows are never read)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import periodic_lbm_b200 as p  # noqa: E402

# 9 / 10: depth-generic two / three steps per pass; 11: FMA twin of the two-step kernels; 12: scalar fp32 collisions (packed is the
# default); 3 / 4 on the FVM / DUGKS paths: FMA twin of the tile kernel / marching kernel (PLBM_MARCH_FORM=1: its sum form)
VARIANTS = tuple(int(v) for v in os.environ.get("PLBM_SANITIZE_VARIANTS", "0,1,2,3,4,5,6,7,8,9,10,11,12").split(","))
rng = np.random.default_rng(1)
for prec in ("f64", "f32"):
    for nx, ny in ((67, 53), (5, 3), (64, 64), (40, 130), (36, 520)):
        # 5..8: the two-step kernels also on grids the cluster kernel would take (6 per-thread loads, 7 bulk async
        # copies, 8 = 7 over the slab schedule's line ranges)
        for variant in VARIANTS:
            g = p.alloc_grid(nx, ny, nf=3, precision=prec)
            p.set_properties(g, 0.02, 0.3, 0.25)
            g.rho[:] = 1.0 + 0.01 * rng.random((nx, ny))
            g.ux[:] = 0.02 * rng.random((nx, ny))
            g.uy[:] = -0.01 * rng.random((nx, ny))
            p.set_pdf_to_equilibrium(g)
            g.set_variant(variant)
            for coll in (p.collide_bgk, p.collide_trt, p.collide_rr, p.collide_bgk_split, p.collide_trt_split, p.collide_bgk_improved):
                g.collision, g.streaming = coll, p.lbm_stream
                p.perform_lbm_step(g, 8 if variant == 10 else 5)
                p.perform_lbm_step(g, 1)
                p.lbm_stream(g)
                coll(g)
            if variant in (0, 2, 3, 4):
                for stream in (p.stream_fvm_bardow, p.stream_fdm_bardow, p.stream_fdm_sofonea):
                    g.collision, g.streaming = p.collide_rr, stream
                    p.perform_step(g, 2)
                    stream(g)
                for stencil in ("wls", "wls_gauss_v1", "wls_gauss_v2", "iso", "default"):
                    g.set_fdm_stencil(stencil)
                    g.collision, g.streaming = p.collide_trt, p.stream_fdm_bardow
                    p.perform_step(g, 1)
                g.collision, g.streaming = p.collide_bgk, p.lbm_stream
                p.perform_triple_step(g, 2)
            if variant in (0, 1, 2, 3, 4):
                g.collision = g.streaming = None
                for dugks in (True, False):
                    g.dugks = dugks
                    p.perform_dugks_step(g, 2)
                    p.update_macros(g)
            p.update_macros(g, lagged=False)
            g.diagnostics()
            p.vorticity_2nd(None, None, grid=g)
            p.vorticity_4th(None, None, grid=g)
            g.l2_error(g.ux, g.uy)
            g.synchronize()
            p.dealloc_grid(g)
sim = p.SimPlugin()
sim.init((33, 21), 1.0, np.zeros((21, 33)), np.zeros((2, 21, 33)))
sim.step(1.3, n=3)
sim.vars()
sim.free()
for name in ("lw", "lw4", "lw6", "fvm"):
    lw = p.SimPlugin(name=name)
    lw.init((33, 21), 0.5, np.zeros((21, 33)), np.zeros((2, 21, 33)))
    lw.step(1.3, n=3)
    lw.vars()
    lw.free()
print("sanitize run complete, launches:", p.launch_count())
