"""The PRODUCT's per-node arithmetic source against the reference's source, on the CPU.

csrc/plbm_math.cuh and csrc/plbm_fv.cuh -- the `__device__` templates every kernel of libplbm_b200.so evaluates its nodes with -- are
compiled as host C++ here (tests/host_math/harness.cpp, a stub cuda_runtime.h, g++ -ffp-contract=off: the individually rounded
operations -fmad=false gives the device build) and fed, node by node, the inputs of tests/golden/refsrc_*.npz; the outputs must be the
ones the reference's executed Fortran source produced (oracle/f90_exec.py), bit for bit, in fp64 and fp32.  This is test
infrastructure, not a CPU path of the product: nothing in periodic_lbm_b200/ can reach it, and the kernels around these functions
(indexing, tiles, schedules, the packed fp32 type) are what the -m gpu parity tests check.  Neither the oracle nor a GPU is involved."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CX = [0, 1, 0, -1, 0, 1, -1, -1, 1]
CY = [0, 0, 1, 0, -1, 1, 1, -1, -1]
MODELS = {"bgk": 0, "trt": 1, "rr": 2, "bgk_cache": 3, "trt_split": 4}


@pytest.fixture(scope="module")
def hm(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    so = tmp_path_factory.mktemp("host_math") / "harness.so"
    r = subprocess.run([gxx, "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared",
                        "-I", os.path.join(ROOT, "tests", "host_math", "stub"), "-I", os.path.join(ROOT, "periodic_lbm_b200", "csrc"),
                        "-o", str(so), os.path.join(ROOT, "tests", "host_math", "harness.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(str(so))
    for sfx, R in (("f64", C.c_double), ("f32", C.c_float)):
        getattr(lib, f"hm_collide_{sfx}").argtypes = [C.c_int, C.c_void_p, R, R, C.c_int]
        getattr(lib, f"hm_equilibrium_{sfx}").argtypes = [R, R, R, C.c_void_p]
        getattr(lib, f"hm_macros_{sfx}").argtypes = [C.c_void_p, C.c_void_p]
        getattr(lib, f"hm_fv_{sfx}").argtypes = [C.c_int, C.c_int, C.c_void_p, R, R, C.c_void_p]
    return lib


def load(prec):
    with np.load(os.path.join(ROOT, "tests", "golden", f"refsrc_{prec}.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize("pre", [0, 1])
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_collisions(hm, prec, pre):
    """collide<T, MODEL> on every node of the streamed lattice == bgk_kernel, bgk_kernel_cache, trt_naive, trt_split, rr_kernel_naive;
    pre = 1: the two-part form the three-level kernel uses (node_reciprocals, then the collision with the reciprocal handed in)"""
    d = load(prec)
    nx, ny = 7, 5
    omega, _, _ = d["params"]
    lam = d["lambda_d"][0]
    fn = getattr(hm, f"hm_collide_{prec}")
    for name, model in MODELS.items():
        got = d["stream"].copy()
        for x in range(nx):
            for y in range(ny):
                node = np.ascontiguousarray(got[:, x, y])
                fn(model, node.ctypes.data, omega, lam, pre)
                got[:, x, y] = node
        assert np.array_equal(got[:, :, :ny], d[name][:, :, :ny]), name
    got = d["f16"].copy()
    for x in range(5):
        for y in range(16):
            node = np.ascontiguousarray(got[:, x, y])
            fn(5, node.ctypes.data, omega, lam, pre)
            got[:, x, y] = node
    assert np.array_equal(got, d["bgk_improved"]), "bgk_improved_kernel"


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_equilibrium_and_macros(hm, prec):
    d = load(prec)
    T = np.float64 if prec == "f64" else np.float32
    nx, ny = 7, 5
    mac = getattr(hm, f"hm_macros_{prec}")
    got = np.zeros((3, nx, ny), dtype=T)
    for x in range(nx):
        for y in range(ny):
            node, out = np.ascontiguousarray(d["stream"][:, x, y]), np.zeros(3, dtype=T)
            mac(node.ctypes.data, out.ctypes.data)
            got[:, x, y] = out
    assert np.array_equal(got, d["macros"]), "update_macros_kernel"
    # equilibrium, as set_pdf_to_equilibrium applied it in the whole-procedure runs: lattice `iold` before the first step is not
    # stored, but run_*.init + zero steps is what random_lattice() of the generator did with the reference's function
    eq = getattr(hm, f"hm_equilibrium_{prec}")
    init = d["run_lbm_bgk.init"]
    out = np.zeros(9, dtype=T)
    eq(init[0, 2, 3], init[1, 2, 3], init[2, 2, 3], out.ctypes.data)
    rho, ux, uy = (T(v) for v in init[:, 2, 3])
    indp = T(1) - T(1.5) * (ux * ux + uy * uy)
    assert out[0] == (T(4) / T(9)) * rho * indp  # feq(0) = w0*rho*indp, src/fvm_bardow.F90:113


def tile_of(f, x, y, n):
    """3 x 3 neighbourhood of node (x, y) of a periodic n x n lattice f[q, x, ld], laid out t[q * 9 + (dx + 1) * 3 + (dy + 1)]"""
    t = np.empty((9, 3, 3), dtype=f.dtype)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            t[:, dx + 1, dy + 1] = f[:, (x + dx) % n, (y + dy) % n]
    return np.ascontiguousarray(t)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_finite_volume_and_finite_difference_updates(hm, prec):
    """flux_update / fdm_update / fdm_stencil_update on the 3 x 3 neighbourhood of every node == fvm_bardow_kernel, kernel_stream with
    and without -DDUGKS (face reconstruction, update_ew / update_ns, flux update), fdm_bardow_kernel in its five builds, fdm_sofonea_kernel"""
    d = load(prec)
    T = np.float64 if prec == "f64" else np.float32
    n = 6
    dt = d["params"][2]
    tau = d["dugks_tau"][0]
    om_face = T(1) / (T(4) * (tau / dt) + T(1))  # dugks_stream, src/periodic_dugks.F90:180-181
    fn = getattr(hm, f"hm_fv_{prec}")
    fq = d["fq"]

    def sweep(mode, stencil, start):
        out = np.zeros_like(fq)
        for x in range(n):
            for y in range(n):
                t = tile_of(fq, x, y, n)
                fp = np.ascontiguousarray(start[:, x, y])
                fn(mode, stencil, t.ctypes.data, dt, om_face, fp.ctypes.data)
                out[:, x, y] = fp
        return out

    assert np.array_equal(sweep(1, 0, fq)[:, :, :n], d["fvm_bardow"][:, :, :n]), "fvm_bardow_kernel: fnew = fold - fluxes"
    assert np.array_equal(sweep(0, 0, d["dugks_fp_in"])[:, :, :n], d["dugks_stream_on"][:, :, :n]), "kernel_stream, -DDUGKS"
    assert np.array_equal(sweep(1, 0, d["dugks_fp_in"])[:, :, :n], d["dugks_stream_off"][:, :, :n]), "kernel_stream without -DDUGKS"
    assert np.array_equal(sweep(2, 0, fq)[:, :, :n], d["fdm_bardow_default"][:, :, :n])
    assert np.array_equal(sweep(3, 0, fq)[:, :, :n], d["fdm_sofonea"][:, :, :n])
    for k, name in enumerate(("wls", "wls_gauss_v1", "wls_gauss_v2", "iso"), start=1):
        assert np.array_equal(sweep(4, k, fq)[:, :, :n], d[f"fdm_bardow_{name}"][:, :, :n]), name
