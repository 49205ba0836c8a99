"""The C oracle against the reference's OWN SOURCE, bit for bit.

`tests/golden/refsrc_{f64,f32}.npz` hold inputs and outputs of the reference's Fortran kernels and host procedures, produced by
EXECUTING the source text of `/root/reference/src/*.f90|F90` statement by statement (oracle/f90_exec.py, a small interpreter with
Fortran's typing rules on IEEE scalars; oracle/make_refsrc_golden.py is the generating script).  No Fortran compiler exists in
this image, so this is as close as the oracle can get to "outputs of the reference itself run here": no human transcription sits
between the Fortran statements and the numbers.  It pins the rows that the reference's two golden files cannot reach:
lbm_stream, collide_trt, collide_rr, the -DSPLIT collisions, collide_bgk_improved, the finite-difference streaming schemes,
vorticity_2nd / 4th, update_macros, set_properties, and every fp32 result -- and re-pins Bardow FVM and DUGKS at kernel level.

Every comparison is np.array_equal.  Where /root/reference exists the fixtures are regenerated and must equal the committed ones."""
import os
import sys

import numpy as np
import pytest

from oracle.oracle import Oracle, OracleGrid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRECS = ["f64", "f32"]


@pytest.fixture(scope="module", params=PRECS)
def ref(request):
    prec = request.param
    with np.load(os.path.join(ROOT, "tests", "golden", f"refsrc_{prec}.npz")) as z:
        data = {k: z[k] for k in z.files}
    return prec, data, Oracle(prec)


def poisoned(o, f, ny):
    """the oracle's own allocation with the stored rows: the padding rows stay NaN, so a kernel that touched them would show"""
    g = o.alloc_f(f.shape[1], ny)
    g[:, :, :ny] = f[:, :, :ny]
    return g


def same(a, b, ny):
    return np.array_equal(a[..., :ny], b[..., :ny])


def test_lbm_kernels_on_a_non_square_grid(ref):
    prec, d, o = ref
    nx, ny = 7, 5
    omega, magic, _ = d["params"]
    f0 = poisoned(o, d["f0"], ny)
    fs = o.alloc_f(nx, ny)
    o.lbm_stream(f0, fs, ny)
    assert same(fs, d["stream"], ny), "lbm_stream_kernel (src/periodic_lbm.f90:45-127)"
    for name, call in (("bgk", lambda f: o.collide_bgk(f, ny, omega)), ("bgk_cache", lambda f: o.kernel_bgk(f, ny, omega)),
                       ("trt", lambda f: o.collide_trt(f, ny, omega, magic)), ("trt_split", lambda f: o.collide_trt_split(f, ny, omega, magic)),
                       ("rr", lambda f: o.collide_rr(f, ny, omega))):
        f = fs.copy()
        call(f)
        assert same(f, d[name], ny), name
    f = f0.copy()
    o.kernel_bgk(f, ny, omega)
    assert same(f, d["dugks_kernel_bgk"], ny), "kernel_bgk (src/periodic_dugks.F90:80-169)"
    lam = o.lambda_d(omega, magic)
    assert lam == d["lambda_d"][0] and o.magic_number(omega, lam) == d["lambda_d"][1]
    rho, ux, uy = o.update_macros(fs, ny)
    assert np.array_equal(np.stack([rho, ux, uy]), d["macros"]), "update_macros_kernel (src/fvm_bardow.F90:356-388)"
    assert np.array_equal(o.vorticity(ux, uy, 2), d["vorticity2"]) and np.array_equal(o.vorticity(ux, uy, 4), d["vorticity4"])


def test_improved_bgk(ref):
    prec, d, o = ref
    f = poisoned(o, d["f16"], 16)
    o.collide_bgk_improved(f, 16, d["params"][0])
    assert same(f, d["bgk_improved"], 16), "bgk_improved_kernel (src/collision_bgk_improved.f90:24-107)"


def test_finite_volume_and_finite_difference_kernels(ref):
    prec, d, o = ref
    n = 6
    dt = d["params"][2]
    fq = poisoned(o, d["fq"], n)
    fnew = o.alloc_f(n, n)
    o.stream_fvm_bardow(fq, fnew, n, dt)
    assert same(fnew, d["fvm_bardow"], n), "fvm_bardow_kernel (src/fvm_bardow.F90:410-509)"
    for stencil in ("default", "wls", "wls_gauss_v1", "wls_gauss_v2", "iso"):
        fnew = o.alloc_f(n, n)
        o.stream_fdm_bardow(fq, fnew, n, dt, stencil)
        assert same(fnew, d[f"fdm_bardow_{stencil}"], n), f"fdm_bardow_kernel, {stencil} build (src/fvm_bardow.F90:526-685)"
    fnew = o.alloc_f(n, n)
    o.stream_fdm_sofonea(fq, fnew, n, dt)
    assert same(fnew, d["fdm_sofonea"], n), "fdm_sofonea_kernel (src/fvm_bardow.F90:703-893)"


@pytest.mark.parametrize("dugks", [True, False])
def test_dugks_stream_kernel(ref, dugks):
    prec, d, o = ref
    n = 6
    fq, fp = poisoned(o, d["fq"], n), poisoned(o, d["dugks_fp_in"], n)
    o.dugks_stream(fq, fp, n, d["dugks_tau"][0], d["params"][2], dugks)
    assert same(fp, d["dugks_stream_on" if dugks else "dugks_stream_off"], n), "kernel_stream + update_ew / update_ns (src/periodic_dugks.F90:190-434)"


def test_flow_cases(ref):
    """taylor_green_t / eval_vortex_case (src/benchmarks): the fixtures carry this image's libm in their last bit (sin, cos, exp), the
    oracle calls the same libm -- equal bits here; two units in the last place are allowed should a machine's libm differ."""
    prec, d, o = ref

    def close(a, b):
        return np.array_equal(a, b) or np.all(np.abs(a - b) <= 2 * np.spacing(np.maximum(np.abs(a), np.abs(b))))

    n, kx, umax, nu, td = d["tg.params"]
    n = int(n)
    assert o.tg_decay_time(kx, kx, nu) == td
    for k in (0, 1):
        assert close(np.stack(o.taylor_green_eval(n, n, kx, kx, umax, td, d[f"tg.t{k}"][0])), d[f"tg.fields{k}"])
    u0, xc, yc, rc, eps = d["vortex.params"]
    assert close(np.stack(o.vortex_eval(9, 7, u0, xc, yc, rc, eps)), d["vortex.fields"])


RUNS = {"run_lbm_bgk": (Oracle.SCHEME_LBM, Oracle.BGK), "run_lbm_trt": (Oracle.SCHEME_LBM, Oracle.TRT), "run_lbm_rr": (Oracle.SCHEME_LBM, Oracle.RR),
        "run_fvm_bgk": (Oracle.SCHEME_FVM_BARDOW, Oracle.BGK), "run_dugks": (Oracle.SCHEME_DUGKS, Oracle.BGK)}


@pytest.mark.parametrize("name", sorted(RUNS))
def test_whole_procedures_on_a_lattice_grid(ref, name):
    """set_properties -> set_pdf_to_equilibrium -> N x perform_*step -> update_macros, as the reference's host procedures ran them in
    the interpreter (procedure pointers bound like app/main_taylor_green.f90:39-40 binds them) against OracleGrid."""
    prec, d, o = ref
    scheme, coll = RUNS[name]
    nsteps, nu, dt, magic = d[f"{name}.args"]
    og = OracleGrid(6, 6, prec)
    og.rho, og.ux, og.uy = (np.ascontiguousarray(a) for a in d[f"{name}.init"])
    p = og.set_properties(nu, dt, magic)
    assert np.array_equal(np.array([p["tau"], p["omega"], p["trt_magic"], p["csqr"]], dtype=o.dtype), d[f"{name}.props"]), "set_properties"
    og.set_pdf_to_equilibrium()
    og.run(scheme, coll, int(nsteps))
    assert [og.iold, og.inew] == list(d[f"{name}.idx"][:2])
    lat = d[f"{name}.lattices"]
    assert same(og.lattice(1), lat[0], 6) and same(og.lattice(2), lat[1], 6), "lattices after the run"
    rho, ux, uy = og.update_macros(lagged=True)
    assert np.array_equal(np.stack([rho, ux, uy]), d[f"{name}.macros"]), "update_macros reads lattice inew (SURVEY F3)"


def test_perform_triple_step(ref):
    """perform_triple_step (src/fvm_bardow.F90:322-340) with lbm_stream + collide_bgk on three lattices: the host logic composed from
    the oracle's kernels like tests/test_gpu_parity.py::test_perform_triple_step does."""
    prec, d, o = ref
    nsteps, nu, dt, magic = d["run_triple_lbm_bgk.args"]
    n = 6
    p = o.set_properties(nu, dt, magic)
    f = [o.alloc_f(n, n) for _ in range(3)]
    inew, iold, imid = 1, 2, 3
    init = d["run_triple_lbm_bgk.init"]
    o.set_pdf_to_equilibrium(*(np.ascontiguousarray(a) for a in init), f[iold - 1])
    for _ in range(int(nsteps)):
        o.lbm_stream(f[iold - 1], f[inew - 1], n)
        f[iold - 1][...] = f[inew - 1]
        o.collide_bgk(f[inew - 1], n, p["omega"])
        iold, inew, imid = inew, imid, iold
    assert [iold, inew, imid] == list(d["run_triple_lbm_bgk.idx"])
    lat = d["run_triple_lbm_bgk.lattices"]
    for k in (iold, imid):  # lattice inew is scratch at this point (never written in the first steps: zeros there, NaN here)
        assert same(f[k - 1], lat[k - 1], n), k


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="the reference's sources are not on this machine")
@pytest.mark.parametrize("prec", PRECS)
def test_committed_fixtures_are_what_the_reference_source_produces_today(prec):
    sys.path.insert(0, os.path.join(ROOT))
    from oracle.make_refsrc_golden import generate

    fresh = generate(prec)
    with np.load(os.path.join(ROOT, "tests", "golden", f"refsrc_{prec}.npz")) as z:
        assert sorted(z.files) == sorted(fresh)
        for k in z.files:
            assert np.array_equal(z[k], fresh[k], equal_nan=True), k


# ---- the plugins of sim/ (the second sandbox of the reference: c_slbm_*, c_lw_*, c_lw4_*, c_lw6_*, c_fvm_*) --------------------------
@pytest.fixture(scope="module")
def sim():
    with np.load(os.path.join(ROOT, "tests", "golden", "refsrc_sim_f64.npz")) as z:
        return {k: z[k] for k in z.files}, Oracle("f64")


P = lambda a: a.ctypes.data  # noqa: E731


def sim_macros(o, f, nx, ny, h):
    r, a, b = (np.zeros((ny, nx)) for _ in range(3))
    (o._sim_macros(nx, ny, P(f), P(r), P(a), P(b)) if h == 1 else o._simh_macros(nx, ny, h, P(f), P(r), P(a), P(b)))
    return np.stack([r, a, b])


def test_standard_lbm_plugin(sim):
    """slbm_step (sim/sim_slbm.F90:81-110) over lbm_collide_and_stream_fused + lbm_periodic_bc_push (sim/sim.F90), DDF-shifted"""
    d, o = sim
    nx, ny, steps, _, omega = d["args"]
    nx, ny = int(nx), int(ny)
    p, u = np.ascontiguousarray(d["p"]), np.ascontiguousarray(d["u"])
    f1 = np.zeros((9, ny + 2, nx + 2))
    o._sim_eqinit(nx, ny, P(f1), P(p), P(u[0]), P(u[1]))
    assert np.array_equal(f1, d["slbm.init"]), "lbm_eqinit_fields"
    f2 = f1.copy()
    for _ in range(int(steps)):
        o._sim_step(nx, ny, P(f1), P(f2), omega)
        o._sim_bc(nx, ny, P(f2))
        f1, f2 = f2, f1
    assert np.array_equal(f1, d["slbm.final"]), "halo included"
    assert np.array_equal(sim_macros(o, f1, nx, ny, 1), d["slbm.macros"]), "lbm_macros"


@pytest.mark.parametrize("name,h,order", [("lw", 1, 2), ("lw4", 2, 4), ("lw6", 3, 6)])
def test_lax_wendroff_plugins(sim, name, h, order):
    """<name>_step (sim/sim_lw.F90, sim_lw4.F90, sim_lw6.F90): <name>_stream, <name>_collision, <name>_bc and the move_alloc swap"""
    d, o = sim
    nx, ny, steps, dt, omega = d["args"]
    nx, ny = int(nx), int(ny)
    p, u = np.ascontiguousarray(d["p"]), np.ascontiguousarray(d["u"])
    f1 = np.zeros((9, ny + 2 * h, nx + 2 * h))
    if h == 1:
        o._sim_eqinit(nx, ny, P(f1), P(p), P(u[0]), P(u[1]))
        o._lw_bc(nx, ny, P(f1))
    else:
        o._simh_eqinit(nx, ny, h, P(f1), P(p), P(u[0]), P(u[1]))
        o._lwh_bc(nx, ny, h, P(f1))
    assert np.array_equal(f1, d[f"{name}.init"])
    f2 = np.zeros_like(f1)
    for _ in range(int(steps)):
        if h == 1:
            o._lw_stream(nx, ny, P(f1), P(f2), dt)
            o._lw_collision(nx, ny, P(f2), omega)
            o._lw_bc(nx, ny, P(f2))
        else:
            o._lwh_stream(order, nx, ny, P(f1), P(f2), dt)
            o._lwh_collision(nx, ny, h, P(f2), omega)
            o._lwh_bc(nx, ny, h, P(f2))
        f1, f2 = f2, f1
    assert np.array_equal(f1, d[f"{name}.final"])
    assert np.array_equal(sim_macros(o, f1, nx, ny, h), d[f"{name}.macros"])


def test_heun_finite_volume_plugin(sim):
    """fvm_step (sim/sim_fvm.F90:285-322): collisions, fvm_predict_hc, fvm_correct_hc, fvm_bc, the three-buffer rotation"""
    d, o = sim
    nx, ny = int(d["args"][0]), int(d["args"][1])
    dt, omega, steps = d["fvm.args"]
    p, u = np.ascontiguousarray(d["p"]), np.ascontiguousarray(d["u"])
    f1 = np.zeros((9, ny + 4, nx + 4))
    f2, fc = np.zeros_like(f1), np.zeros_like(f1)
    o._simh_eqinit(nx, ny, 2, P(f1), P(p), P(u[0]), P(u[1]))
    o._lwh_bc(nx, ny, 2, P(f1))
    assert np.array_equal(f1, d["fvm.init"])
    for _ in range(int(steps)):
        o._simfvm_step(nx, ny, P(f1), P(f2), P(fc), dt, omega)
        f1, fc = fc, f1
    assert np.array_equal(f1, d["fvm.final"])


@pytest.mark.skipif(not os.path.isdir("/root/reference/sim"), reason="the reference's sources are not on this machine")
def test_committed_sim_fixtures_are_what_the_reference_source_produces_today():
    from oracle.make_refsrc_golden import generate_sim

    fresh = generate_sim()
    with np.load(os.path.join(ROOT, "tests", "golden", "refsrc_sim_f64.npz")) as z:
        assert sorted(z.files) == sorted(fresh)
        for k in z.files:
            assert np.array_equal(z[k], fresh[k]), k


# ---- the configuration of the reference's golden files, a few steps of it ------------------------------------------------------------
TG64 = os.path.join(ROOT, "tests", "golden", "refsrc_tg64_f64.npz")


@pytest.mark.skipif(not os.path.exists(TG64), reason="generated on request: python oracle/make_refsrc_golden.py --tg64 (ten minutes)")
@pytest.mark.parametrize("name,scheme", [("dugks", Oracle.SCHEME_DUGKS), ("fvm_bgk", Oracle.SCHEME_FVM_BARDOW), ("lbm_bgk", Oracle.SCHEME_LBM)])
def test_golden_configuration_64x64(name, scheme):
    """Taylor-Green 64 x 64, Re = 100 -- what graphs/fvm_*_64.txt were produced with -- executed from the reference's source for a few
    steps (taylor_green_eval, set_properties without a magic number, set_pdf_to_equilibrium, perform_*step, update_macros): the oracle
    holds the same lattices (sha256 of the physical rows), indices and macroscopic fields."""
    import hashlib

    with np.load(TG64) as z:
        d = {k: z[k] for k in z.files}
    n = 64
    nsteps, dt = d[f"{name}.args"]
    og = OracleGrid(n, n, "f64")
    og.rho, og.ux, og.uy = (np.ascontiguousarray(a) for a in d[f"{name}.init"])
    p = og.set_properties(d["params"][2], dt, None)
    assert np.array_equal(np.array([p["tau"], p["omega"], p["trt_magic"], p["csqr"]]), d[f"{name}.props"])
    og.set_pdf_to_equilibrium()
    og.run(scheme, Oracle.BGK, int(nsteps))
    assert [og.iold, og.inew] == list(d[f"{name}.idx"])
    for k in (1, 2):
        digest = hashlib.sha256(np.ascontiguousarray(og.lattice(k)[:, :, :n]).tobytes()).digest()
        assert digest == d[f"{name}.digests"][k - 1].tobytes(), f"lattice {k}"
    assert np.array_equal(np.stack(og.update_macros(lagged=True)), d[f"{name}.macros"])
    # and the initial fields are the oracle's own Taylor-Green evaluation + density conversion
    pr, ux, uy = og.o.taylor_green_eval(n, n, d["params"][4], d["params"][4], d["params"][1], d["params"][5], 0.0)
    assert np.array_equal(np.stack([pr / p["csqr"] + 1.0, ux, uy]), d[f"{name}.init"])


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="the reference's sources are not on this machine")
@pytest.mark.parametrize("prec", PRECS)
def test_the_comparison_is_sensitive_to_one_reassociated_statement(prec):
    """Mutation check of the whole arrangement: re-associate ONE sum in the reference's TRT kernel text (mathematically the same
    value) and the interpreter's output must no longer equal the fixture the oracle reproduces -- bit-for-bit equality with the
    unmodified source is therefore a statement about the order of every operation, not a coincidence of loose arithmetic."""
    from oracle.f90_exec import Interp
    from oracle.make_refsrc_golden import F, ld_of

    with np.load(os.path.join(ROOT, "tests", "golden", f"refsrc_{prec}.npz")) as z:
        d = {k: z[k] for k in z.files}
    text = open("/root/reference/src/collision_trt.F90").read()
    old = "( vN + vS - fac1 * velY2 - t1x2 * feq_common )"
    assert text.count(old) == 1
    outs = []
    for src in (text, text.replace(old, "( vN + vS - t1x2 * feq_common - fac1 * velY2 )")):
        it = Interp(prec)
        for f in ("precision.F90", "fvm_bardow.F90"):
            it.load(f)
        it.load_text(src)
        f = d["stream"].copy()
        omega = d["params"][0]
        it.run("collision_trt", "collide_trt/trt_naive", 7, 5, F(f), ld_of(5), omega, d["lambda_d"][0])
        outs.append(f)
    assert np.array_equal(outs[0], d["trt"])          # the text as it is: the fixture
    assert not np.array_equal(outs[1], d["trt"])      # one re-associated sum: different last bits somewhere
    assert np.allclose(outs[1], d["trt"], rtol=1e-5 if prec == "f32" else 1e-13, atol=0)
