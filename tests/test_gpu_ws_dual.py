"""The warp-specialised three-step kernel (csrc/plbm_lbm3w.cu) and the closing dual triple (Grid::spare, csrc/plbm_api.cu
step_lbm_t) against the oracle, bit for bit, on awkward grids: every collision, fp64 and fp32, several calls per grid.

The library reads its PLBM_* knobs once per process, so every combination runs tests/ws_dual_worker.py in a process of its own:
PLBM_TRIPLES=2 makes the default stepping take triples on small grids, PLBM_SPARE_LATTICE=2 gives them the third lattice buffer."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("ws,spare", [(1, 2), (0, 2), (1, 0), (None, 2)])
def test_three_step_kernels_and_closing_dual_triple(plbm, ws, spare):
    """ws: 1 / 0 = k_lbm3_ws / k_lbmn_bulk forced for every collision, None = the library's own choice per collision"""
    env = dict(os.environ, PLBM_TRIPLES="2", PLBM_SPARE_LATTICE=str(spare))
    env.pop("PLBM_TRIPLE_WS", None)
    if ws is not None:
        env["PLBM_TRIPLE_WS"] = str(ws)
    kernel = "auto" if ws is None else ("k_lbm3_ws" if ws else "k_lbmn_bulk")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ws_dual_worker.py"), kernel, "1" if spare else "0"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["ok"] and out["cases"] == 96 and out["kernel"] == kernel and out["dual"] == bool(spare)
