import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import periodic_lbm_b200 as p
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prec = sys.argv[2] if len(sys.argv) > 2 else "f64"
g = p.alloc_grid(n, n, precision=prec)
p.set_properties(g, 0.02, 0.3, 0.25)
g.rho[:] = 1.0; g.ux[:] = 0.01; g.uy[:] = -0.02
p.set_pdf_to_equilibrium(g)
p.perform_dugks_step(g, 2)
g.synchronize()
print("dugks ok", g.download_f(g.iold)[1, 3, 5])
g.collision, g.streaming = p.collide_bgk, p.stream_fvm_bardow
p.perform_step(g, 2)
g.synchronize()
print("fvm ok", g.download_f(g.iold)[1, 3, 5])
