import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_built():
    """Never test a stale libplbm_b200.so: rebuild (incremental `make`) whenever the sources differ from the ones the
    library was built from (content hash, __graft_entry__.source_stamp -- file times do not survive the copy to the GPU box)."""
    import __graft_entry__ as ge

    if not ge.is_current():
        ge.build()
    from oracle import oracle as orc

    orc.build()


_ensure_built()

SEED = 20261017


def random_state(o, nx, ny, seed=SEED, amp=1e-3):
    """Random near-equilibrium lattice: f = feq(rho,u) * (1 + amp*xi), rho in [0.9,1.1],
    |u| < 0.1, xi ~ U(-1,1); padding rows are NaN-poisoned."""
    rng = np.random.default_rng(seed)
    T = o.dtype
    rho = (0.9 + 0.2 * rng.random((nx, ny))).astype(T)
    ang = 2 * np.pi * rng.random((nx, ny))
    mag = 0.1 * rng.random((nx, ny))
    ux = (mag * np.cos(ang)).astype(T)
    uy = (mag * np.sin(ang)).astype(T)
    f = o.alloc_f(nx, ny)
    o.set_pdf_to_equilibrium(rho, ux, uy, f)
    xi = (2 * rng.random(f.shape) - 1).astype(T)
    f[:, :, :ny] = f[:, :, :ny] * (T(1) + T(amp) * xi[:, :, :ny])
    return f


@pytest.fixture(scope="session")
def plbm():
    import periodic_lbm_b200 as p

    return p
