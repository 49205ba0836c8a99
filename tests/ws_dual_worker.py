"""Worker of tests/test_gpu_ws_dual.py: runs in its own process because the library reads its PLBM_* knobs once.

Checks, against the CPU oracle and bit for bit, perform_lbm_step calls whose three-step launches go to the kernel the environment
selects (PLBM_TRIPLE_WS) and which close with a dual triple when a third lattice buffer is available (PLBM_SPARE_LATTICE): both
lattices, the lattice indices and the lagged macroscopic fields after every call of a sequence of calls on the same grid."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import periodic_lbm_b200 as plbm  # noqa: E402
from conftest import random_state  # noqa: E402
from oracle.oracle import Oracle, OracleGrid  # noqa: E402
from periodic_lbm_b200.slab import launch_schedule  # noqa: E402


def main():
    want_kernel, want_dual = sys.argv[1], sys.argv[2] == "1"
    sizes = [(40, 516), (300, 260), (37, 2048), (8, 32), (16, 132), (70, 32), (6, 64), (9, 1000)]
    colls = ((plbm.collide_bgk, Oracle.BGK), (plbm.collide_trt, Oracle.TRT), (plbm.collide_rr, Oracle.RR),
             (plbm.collide_bgk_split, Oracle.BGK_SPLIT), (plbm.collide_trt_split, Oracle.TRT_SPLIT),
             (plbm.collide_bgk_improved, Oracle.BGK_IMPROVED))
    calls_of = [(3, 6), (5, 3), (8, 4), (9, 5), (11, 7), (6, 3, 3), (13, 1, 3), (20, 2, 6)]
    ncases = 0
    for prec in ("f64", "f32"):
        for i, (nx, ny) in enumerate(sizes):
            for c, (coll, ocoll) in enumerate(colls):
                calls = calls_of[(i + c) % len(calls_of)]
                og = OracleGrid(nx, ny, prec)
                og.set_properties(0.02, 1.0, 0.25)
                f0 = random_state(og.o, nx, ny, seed=100 + 7 * i + c)
                og.lattice(og.iold)[...] = f0
                og.lattice(og.inew)[...] = 0
                g = plbm.alloc_grid(nx, ny, precision=prec)
                plbm.set_properties(g, 0.02, 1.0, 0.25)
                g.upload_f(g.iold, np.nan_to_num(f0, nan=0.0))
                g.upload_f(g.inew, np.zeros_like(f0))
                g.collision, g.streaming = coll, plbm.lbm_stream
                assert g.steps_per_pass() == 3, (nx, ny, prec, g.steps_per_pass())
                auto = "k_lbmn_bulk" if coll in (plbm.collide_trt, plbm.collide_trt_split) else "k_lbm3_ws"  # the library's own rule
                assert g.triple_kernel() == (auto if want_kernel == "auto" else want_kernel), (g.triple_kernel(), coll.__name__)
                assert g.closing_triple() == want_dual, (g.closing_triple(), want_dual)
                for nsteps in calls:
                    l0 = plbm.launch_count()
                    plbm.perform_lbm_step(g, nsteps)
                    launches = plbm.launch_count() - l0
                    sched = launch_schedule(nsteps, pairs=g.pair_kernel() != "k_lbm", triples=True, dual=want_dual)
                    assert launches == len(sched), (nx, ny, prec, nsteps, launches, sched)
                    og.run(Oracle.SCHEME_LBM, ocoll, nsteps)
                    tag = (nx, ny, prec, coll.__name__, calls, nsteps)
                    assert (g.iold, g.inew) == (og.iold, og.inew), tag
                    for which_g, which_o in ((g.iold, og.iold), (g.inew, og.inew)):
                        got = g.download_f(which_g)[:, :, :ny]
                        want = og.lattice(which_o)[:, :, :ny]
                        assert np.array_equal(got, want), (tag, which_g, float(np.abs(got - want).max()), int((got != want).sum()))
                    plbm.update_macros(g)
                    r, u, v = og.update_macros(lagged=True)
                    assert np.array_equal(g.rho, r) and np.array_equal(g.ux, u) and np.array_equal(g.uy, v), tag
                plbm.dealloc_grid(g)
                ncases += 1
    print(json.dumps({"ok": True, "cases": ncases, "kernel": want_kernel, "dual": want_dual}))


if __name__ == "__main__":
    main()
