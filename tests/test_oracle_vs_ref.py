"""The C oracle against the UNMODIFIED reference kernels (oracle/_ref/libplbm_ref.so, built by
`make -C oracle ref` from the sources under /root/reference).  No Fortran compiler exists in the image this
repository is developed in, so this test normally SKIPS; it is the bit-for-bit check to run
on any machine that has gfortran."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle.oracle import Oracle, OracleGrid, padded_ld

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libplbm_ref.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built: no Fortran compiler in this image")


@pytest.mark.parametrize("scheme,collision", [(0, 0), (0, 1), (0, 2), (1, 0), (2, 0), (4, 0), (5, 0)])
def test_oracle_bitwise_equals_reference(scheme, collision):
    lib = C.CDLL(REF)
    nx = ny = 48
    ld = padded_ld(ny)
    rng = np.random.default_rng(2)
    rho = 0.95 + 0.1 * rng.random((nx, ny))
    ux = 0.05 * (rng.random((nx, ny)) - 0.5)
    uy = 0.05 * (rng.random((nx, ny)) - 0.5)
    nu, dt, magic, nsteps = 0.02, (1.0 if scheme == 0 else 0.3), 0.25, 25
    og = OracleGrid(nx, ny)
    og.set_properties(nu, dt, magic)
    og.rho, og.ux, og.uy = rho.copy(), ux.copy(), uy.copy()
    og.set_pdf_to_equilibrium()
    og.run(scheme, collision, nsteps)
    r, u, v = og.update_macros(lagged=True)

    rr, uu, vv = rho.copy(), ux.copy(), uy.copy()
    f = np.zeros((9, nx, ld))
    iold = C.c_int()
    P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    lib.ref_run(C.c_int(nx), C.c_int(ny), C.c_int(ld), C.c_int(scheme), C.c_int(collision), C.c_double(nu), C.c_double(dt),
                C.c_double(magic), C.c_int(nsteps), P(rr), P(uu), P(vv), P(f), C.byref(iold))
    assert iold.value == og.iold
    assert np.array_equal(f[:, :, :ny], og.lattice(og.iold)[:, :, :ny])
    assert np.array_equal(rr, r) and np.array_equal(uu, u) and np.array_equal(vv, v)
