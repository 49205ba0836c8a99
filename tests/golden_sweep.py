"""Print our GPU value next to every row of the reference golden files (helper script, lives under tests/ because it uses the oracle-backed test helpers)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import periodic_lbm_b200 as p
from test_gpu_golden import gpu_tg_run, load_rows
for name, scheme in (("ref_fvm_bardow_64.txt", "fvm"), ("ref_fvm_dugks_64.txt", "dugks")):
    for r, gold in sorted(load_rows(name).items(), reverse=True):
        l2, steps, t, g, _, _ = gpu_tg_run(p, 64, scheme, p.collide_bgk, dt_over_tau=r)
        p.dealloc_grid(g)
        print(f"{scheme:6s} r={r:7.4f} steps={steps:8d} gold={gold:.7E} ours={l2:.10E} rel={abs(l2-gold)/gold:.2e} abs={abs(l2-gold):.2e}", flush=True)
