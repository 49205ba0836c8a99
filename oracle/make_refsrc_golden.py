"""TEST INFRASTRUCTURE: golden vectors from the reference's own Fortran source, executed by oracle/f90_exec.py.

    python oracle/make_refsrc_golden.py          # writes tests/golden/refsrc_f64.npz and refsrc_f32.npz

Needs /root/reference (this container); the fixtures travel, the reference does not.  Every array is stored in the ORACLE's
layout (lattices f[q, x, ld], fields [x, y]: the transposes of the Fortran arrays), inputs next to outputs, so that
tests/test_oracle_refsrc.py can feed the C oracle the same bits and demand the same bits back.

What runs, per precision (wp = real64 / real32, i.e. the reference built without / with -DPRECISION_SINGLE):
  kernels, called exactly as their host procedures call them
    periodic_lbm.f90          lbm_stream_kernel
    collision_bgk.F90         bgk_kernel, bgk_kernel_cache (-DSPLIT)
    collision_trt.F90         trt_naive, trt_split (-DSPLIT), lambda_d, magic_number
    collision_regularized.F90 rr_kernel_naive
    collision_bgk_improved.f90 bgk_improved_kernel
    fvm_bardow.F90            equilibrium (through set_pdf_to_equilibrium), update_macros_kernel, fvm_bardow_kernel,
                              fdm_bardow_kernel (default, -DFDM_WLS, -DFDM_WLS_GAUSS_V1, -DFDM_WLS_GAUSS_V2, -DFDM_ISO), fdm_sofonea_kernel
    periodic_dugks.F90        kernel_bgk, kernel_stream (+ update_ew / update_ns) with and without -DDUGKS
    vorticity.f90             vorticity_2nd, vorticity_4th
    benchmarks/*.f90          taylor_green_t (constructor, eval), eval_vortex_case -- transcendental functions from this machine's libm
  whole procedures on a lattice_grid object (set_properties, set_pdf_to_equilibrium, N x perform_*step, update_macros), with the
  procedure pointers grid%streaming / grid%collision bound like the drivers bind them
    perform_lbm_step  x  collide_bgk / collide_trt / collide_rr;  perform_step (stream_fvm_bardow + collide_bgk);
    perform_dugks_step (-DDUGKS);  perform_triple_step (lbm_stream + collide_bgk, three lattices).
  and the plugins of sim/ (generate_sim, refsrc_sim_f64.npz): slbm, lw, lw4, lw6, fvm -- init, N x <plugin>_step, lbm_macros."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != os.path.join(ROOT, "oracle")]  # `oracle` = the directory, not oracle.py
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.f90_exec import FArray, Interp  # noqa: E402

FILES = ["precision.F90", "fvm_bardow.F90", "collision_bgk.F90", "collision_trt.F90", "collision_regularized.F90",
         "collision_bgk_improved.f90", "periodic_lbm.f90", "periodic_dugks.F90", "vorticity.f90",
         "benchmarks/taylor_green.f90", "benchmarks/barotropic_vortex_case.F90"]
SEED = 20261018


def interp(prec, **macros):
    it = Interp(prec, macros)
    for f in FILES:
        it.load(f)
    return it


def ld_of(ny):
    return (ny + 15) // 16 * 16


def random_lattice(dtype, nx, ny, seed, it):
    """near-equilibrium lattice in the oracle's layout f[q, x, ld] (padding rows zero), built with the reference's equilibrium"""
    rng = np.random.default_rng(seed)
    f = np.zeros((9, nx, ld_of(ny)), dtype=dtype)
    for x in range(nx):
        for y in range(ny):
            rho = dtype(0.9 + 0.2 * rng.random())
            ang, mag = 2 * np.pi * rng.random(), 0.1 * rng.random()
            feq = it.run("fvm_bardow", "equilibrium", rho, dtype(mag * np.cos(ang)), dtype(mag * np.sin(ang)))
            f[:, x, y] = feq * (dtype(1) + dtype(1e-3) * (2 * rng.random(9) - 1).astype(dtype))
    return f


def F(a):
    """oracle layout -> the Fortran array it is the transpose of (a view: the kernels write through it)"""
    return a.transpose(*range(a.ndim - 1, -1, -1))


def new_grid(it, nx, ny, nf=2):
    """what alloc_grid leaves (src/fvm_bardow.F90:129-173), without the log file"""
    dt = it.wp
    f = np.zeros((nf, 9, nx, ld_of(ny)), dtype=dt)  # transposes of grid%f(ld, nx, 0:8, nf)
    mf = np.zeros((3, nx, ny), dtype=dt)
    g = {"nx": nx, "ny": ny, "f": FArray(F(f), [1, 1, 0, 1]), "mf": FArray(F(mf)),
         "rho": FArray(F(mf[0])), "ux": FArray(F(mf[1])), "uy": FArray(F(mf[2])),
         "inew": 1, "iold": 2, "imid": 3 if nf > 2 else -1, "_f": f, "_mf": mf}
    return g


def generate(prec):
    dtype = np.float64 if prec == "f64" else np.float32
    out = {}
    it = interp(prec)
    it_split = interp(prec, SPLIT=1)
    it_dugks = interp(prec, DUGKS=1)
    nx, ny = 7, 5
    ld = ld_of(ny)
    omega, magic, dt_fv = dtype(1.7), dtype(0.25), dtype(0.3)
    f0 = random_lattice(dtype, nx, ny, SEED, it)
    out["f0"] = f0
    out["params"] = np.array([omega, magic, dt_fv], dtype=dtype)

    # ---- kernels ------------------------------------------------------------------------------------------------------
    fs = np.zeros_like(f0)
    it.run("periodic_lbm", "lbm_stream/lbm_stream_kernel", nx, ny, ld, F(f0.copy()), F(fs))
    out["stream"] = fs.copy()

    def inplace(interp_, module, path, *args_after, pre=(), src=fs):
        f = src.copy()
        interp_.run(module, path, *pre, F(f), *args_after)
        return f

    out["bgk"] = inplace(it, "collision_bgk", "collide_bgk/bgk_kernel", omega, pre=(nx, ny, ld))
    out["bgk_cache"] = inplace(it_split, "collision_bgk", "collide_bgk/bgk_kernel_cache", omega, pre=(nx, ny, ld))
    lam = it.run("collision_trt", "lambda_d", omega, magic)
    out["lambda_d"] = np.array([lam, it.run("collision_trt", "magic_number", omega, lam)], dtype=dtype)
    out["trt"] = inplace(it, "collision_trt", "collide_trt/trt_naive", ld, omega, lam, pre=(nx, ny))
    out["trt_split"] = inplace(it_split, "collision_trt", "collide_trt/trt_split", ld, omega, lam, pre=(nx, ny))
    out["rr"] = inplace(it, "collision_regularized", "collide_rr/rr_kernel_naive", omega, pre=(nx, ny, ld))
    # collide_bgk_improved declares f1(ny, nx, 0:8): run it where ld == ny
    f16 = random_lattice(dtype, 5, 16, SEED + 1, it)
    out["f16"] = f16
    out["bgk_improved"] = inplace(it, "collision_bgk_improved", "collide_bgk_improved/bgk_improved_kernel", omega, pre=(5, 16), src=f16)
    rho, ux, uy = (np.zeros((nx, ny), dtype=dtype) for _ in range(3))
    it.run("fvm_bardow", "update_macros/update_macros_kernel", nx, ny, ld, F(fs.copy()), F(rho), F(ux), F(uy))
    out["macros"] = np.stack([rho, ux, uy])
    for order in (2, 4):
        om = np.zeros((nx, ny), dtype=dtype)
        it.run("vorticity", f"vorticity_{'2nd' if order == 2 else '4th'}", F(ux.copy()), F(uy.copy()), F(om))
        out[f"vorticity{order}"] = om

    # the finite-volume / finite-difference / DUGKS kernels on a SQUARE grid: stream_fvm_bardow and stream_fdm_sofonea loop
    # `do x = 1, ny / do y = 1, nx` (harmless iff nx == ny; the oracle follows the intended bounds)
    nq = 6
    fq = random_lattice(dtype, nq, nq, SEED + 2, it)
    out["fq"] = fq

    def two_lattice(interp_, module, path, *tail, src=fq):
        fnew = np.zeros_like(src)
        interp_.run(module, path, nq, nq, ld_of(nq), F(src.copy()), F(fnew), *tail)
        return fnew

    out["fvm_bardow"] = two_lattice(it, "fvm_bardow", "stream_fvm_bardow/fvm_bardow_kernel", dt_fv)
    out["fdm_bardow_default"] = two_lattice(it, "fvm_bardow", "stream_fdm_bardow/fdm_bardow_kernel", dt_fv)
    for name, macro in (("wls", "FDM_WLS"), ("wls_gauss_v1", "FDM_WLS_GAUSS_V1"), ("wls_gauss_v2", "FDM_WLS_GAUSS_V2"), ("iso", "FDM_ISO")):
        out[f"fdm_bardow_{name}"] = two_lattice(interp(prec, **{macro: 1}), "fvm_bardow", "stream_fdm_bardow/fdm_bardow_kernel", dt_fv)
    out["fdm_sofonea"] = two_lattice(it, "fvm_bardow", "stream_fdm_sofonea/fdm_sofonea_kernel", dt_fv)
    out["dugks_kernel_bgk"] = inplace(it, "periodic_dugks", "kernel_bgk", omega, pre=(nx, ny, ld), src=f0)
    # kernel_stream as dugks_stream calls it (src/periodic_dugks.F90:172-188): omega = 1 / (4 tau / dt + 1)
    tau = dtype(0.06)
    om_face = dtype(1) / (dtype(4) * (tau / dt_fv) + dtype(1))
    fp_in = random_lattice(dtype, nq, nq, SEED + 3, it)  # any second lattice: kernel_stream updates fp from the faces of ft
    out["dugks_fp_in"] = fp_in
    for name, interp_ in (("dugks_stream_on", it_dugks), ("dugks_stream_off", it)):
        fp = fp_in.copy()
        interp_.run("periodic_dugks", "kernel_stream", nq, nq, ld_of(nq), F(fq.copy()), F(fp), dt_fv, om_face)
        out[name] = fp
    out["dugks_tau"] = np.array([tau], dtype=dtype)

    # ---- whole procedures on a lattice_grid -------------------------------------------------------------------------------------
    def drive(interp_, step_proc, streaming, collision, nsteps, nu, dt, magic_, nf=2, step_module="fvm_bardow"):
        g = new_grid(interp_, 6, 6, nf)
        rng = np.random.default_rng(SEED + 7)
        g["_mf"][0] = (0.95 + 0.1 * rng.random((6, 6))).astype(dtype)
        g["_mf"][1] = (0.06 * (rng.random((6, 6)) - 0.5)).astype(dtype)
        g["_mf"][2] = (0.06 * (rng.random((6, 6)) - 0.5)).astype(dtype)
        init = g["_mf"].copy()
        interp_.run("fvm_bardow", "set_properties", g, dtype(nu), dtype(dt), dtype(magic_))
        interp_.run("fvm_bardow", "set_pdf_to_equilibrium", g)
        if streaming:
            g["streaming"] = lambda: interp_.run(streaming[0], streaming[1], g)
        if collision:
            g["collision"] = lambda: interp_.run(collision[0], collision[1], g)
        for _ in range(nsteps):
            interp_.run(step_module, step_proc, g)
        interp_.run("fvm_bardow", "update_macros", g)
        props = np.array([g["tau"], g["omega"], g["trt_magic"], g["csqr"]], dtype=dtype)
        return dict(init=init, props=props, lattices=g["_f"].copy(), idx=np.array([g["iold"], g["inew"], g["imid"]]), macros=g["_mf"].copy())

    runs = {
        "run_lbm_bgk": (it, "perform_lbm_step", ("periodic_lbm", "lbm_stream"), ("collision_bgk", "collide_bgk"), 4, 0.02, 1.0, 0.25, 2, "periodic_lbm"),
        "run_lbm_trt": (it, "perform_lbm_step", ("periodic_lbm", "lbm_stream"), ("collision_trt", "collide_trt"), 4, 0.02, 1.0, 0.1875, 2, "periodic_lbm"),
        "run_lbm_rr": (it, "perform_lbm_step", ("periodic_lbm", "lbm_stream"), ("collision_regularized", "collide_rr"), 4, 0.02, 1.0, 0.25, 2, "periodic_lbm"),
        "run_fvm_bgk": (it, "perform_step", ("fvm_bardow", "stream_fvm_bardow"), ("collision_bgk", "collide_bgk"), 3, 0.02, 0.3, 0.25, 2, "fvm_bardow"),
        "run_dugks": (it_dugks, "perform_dugks_step", ("periodic_dugks", "dugks_stream"), ("periodic_dugks", "dugks_collide"), 3, 0.02, 0.3, 0.25, 2,
                      "periodic_dugks"),
        "run_triple_lbm_bgk": (it, "perform_triple_step", ("periodic_lbm", "lbm_stream"), ("collision_bgk", "collide_bgk"), 4, 0.02, 1.0, 0.25, 3,
                               "fvm_bardow"),
    }
    for name, (interp_, step_proc, streaming, collision, nsteps, nu, dt, mg, nf, mod) in runs.items():
        r = drive(interp_, step_proc, streaming, collision, nsteps, nu, dt, mg, nf, mod)
        for k, v in r.items():
            out[f"{name}.{k}"] = v
        out[f"{name}.args"] = np.array([nsteps, nu, dt, mg], dtype=np.float64)
    # ---- the flow cases (src/benchmarks): sin / cos / exp come from this machine's libm, like in a gfortran binary built here ----------
    n = 12
    umax = dtype(0.01) / np.sqrt(dtype(3.0))
    nu_tg = (umax * dtype(n)) / dtype(100.0)
    kx = dtype(2) * it.constant("taylor_green", "pi") / dtype(n)
    case = it.run("taylor_green", "taylor_green_t_constructor", n, n, kx, kx, umax, nu_tg)
    out["tg.params"] = np.array([n, kx, umax, nu_tg, case["td"]], dtype=dtype)
    for k, t in enumerate((dtype(0.0), dtype(123.5))):
        p, ux, uy = (np.zeros((n, n), dtype=dtype) for _ in range(3))
        it.run("taylor_green", "taylor_green_eval", case, t, F(p), F(ux), F(uy))
        out[f"tg.fields{k}"] = np.stack([p, ux, uy])
        out[f"tg.t{k}"] = np.array([t], dtype=dtype)
    vc = {"u0": dtype(0.05), "xc": dtype(6.2), "yc": dtype(5.1), "rc": dtype(2.5), "eps": dtype(0.02), "rho0": dtype(1.0),
          "csqr": dtype(1) / dtype(3)}
    r, a, b = (np.zeros((9, 7), dtype=dtype) for _ in range(3))
    it.run("barotropic_vortex_case", "eval_vortex_case", vc, 9, 7, F(r), F(a), F(b))
    out["vortex.params"] = np.array([vc["u0"], vc["xc"], vc["yc"], vc["rc"], vc["eps"]], dtype=dtype)
    out["vortex.fields"] = np.stack([r, a, b])
    out["statements_executed"] = np.array([it.nstmt + it_split.nstmt + it_dugks.nstmt])
    return out


SIM_FILES = ["sim.F90", "sim_lw.F90", "sim_lw4.F90", "sim_lw6.F90", "sim_fvm.F90", "sim_slbm.F90"]


def generate_sim():
    """The plugins of sim/ (fp64: `wp => dp`), built as gfortran builds them (-D__GFORTRAN__ branches): the init sequence of
    <plugin>_init without its allocate statements (lbm_eqinit_fields + the halo fill), then <plugin>_step / slbm_step / fvm_step run
    as whole procedures on the plugin object (associate, keyword arguments, move_alloc swaps), then lbm_macros.  Arrays in the oracle's
    layout f[q, y, x] with the halo, fields [y, x]."""
    it = Interp("f64", {"__GFORTRAN__": 1}, src="/root/reference/sim")
    for f in SIM_FILES:
        it.load(f)
    out = {}
    nx, ny, steps = 7, 6, 3
    rng = np.random.default_rng(SEED + 11)
    p = 1e-3 * rng.standard_normal((ny, nx))
    u = 0.05 * rng.standard_normal((2, ny, nx))
    out["p"], out["u"] = p, u
    out["args"] = np.array([nx, ny, steps, 0.4, 1.4])  # nx, ny, steps, dt of the Lax-Wendroff / Heun plugins, omega
    dt, omega = 0.4, 1.4

    def arr(h):
        return np.zeros((9, ny + 2 * h, nx + 2 * h))

    def macros(f, h):
        r, a, b = (np.zeros((ny, nx)) for _ in range(3))
        it.run("lbm_primitives", "lbm_macros", nx, ny, h, F(f), F(r), F(a), F(b))
        return np.stack([r, a, b])

    # standard LBM behind c_slbm_* (sim/sim_slbm.F90, sim/sim.F90 lbm_primitives)
    g1, g2 = arr(1), arr(1)
    this = {"dt": np.float64(1.0), "flip": 0, "grid": {"nx": nx, "ny": ny, "f1": FArray(F(g1), [0, 0, 0]), "f2": FArray(F(g2), [0, 0, 0])}}
    it.run("lbm_primitives", "lbm_eqinit_fields", nx, ny, 1, this["grid"]["f1"].a, F(p), F(u[0]), F(u[1]))
    this["grid"]["f2"].a[...] = this["grid"]["f1"].a
    out["slbm.init"] = g1.copy()
    for _ in range(steps):
        it.run("sim_slbm_class", "slbm_step", this, np.float64(omega))
    out["slbm.final"] = F(this["grid"]["f1"].a).copy()
    out["slbm.macros"] = macros(out["slbm.final"], 1)
    # Lax-Wendroff plugins: lw (2nd order, halo 1), lw4 (halo 2), lw6 (halo 3)
    for name, h, cls, mod in (("lw", 1, "sim_lw_class", "lbm_lw"), ("lw4", 2, "sim_lw4_class", "lbm_lw4"), ("lw6", 3, "sim_lw6_class", "lbm_lw6")):
        g1, g2 = arr(h), arr(h)
        lo = [1 - h, 1 - h, 0]
        this = {"dt": np.float64(dt), "flip": 0, "grid": {"nx": nx, "ny": ny, "f1": FArray(F(g1), lo), "f2": FArray(F(g2), lo)}}
        it.run("lbm_primitives", "lbm_eqinit_fields", nx, ny, h, this["grid"]["f1"].a, F(p), F(u[0]), F(u[1]))
        it.run(mod, f"{name}_bc", nx, ny, this["grid"]["f1"].a)
        out[f"{name}.init"] = g1.copy()
        for _ in range(steps):
            it.run(cls, f"{name}_step", this, np.float64(omega))
        out[f"{name}.final"] = F(this["grid"]["f1"].a).copy()
        out[f"{name}.macros"] = macros(out[f"{name}.final"], h)
    # Heun finite-volume plugin (sim/sim_fvm.F90)
    g = [arr(2) for _ in range(3)]
    lo = [-1, -1, 0]
    this = {"dt": np.float64(0.3), "nx": nx, "ny": ny, "f1": FArray(F(g[0]), lo), "f2": FArray(F(g[1]), lo), "fc": FArray(F(g[2]), lo)}
    it.run("lbm_primitives", "lbm_eqinit_fields", nx, ny, 2, this["f1"].a, F(p), F(u[0]), F(u[1]))
    it.run("lbm_fvm", "fvm_bc", nx, ny, this["f1"].a)
    out["fvm.init"] = g[0].copy()
    for _ in range(2):
        it.run("sim_fvm_class", "fvm_step", this, np.float64(1.2))
    out["fvm.final"] = F(this["f1"].a).copy()
    out["fvm.args"] = np.array([0.3, 1.2, 2])
    out["statements_executed"] = np.array([it.nstmt])
    return out


def lattice_digest(f, ny):
    """sha256 of the physical rows of a lattice f[q, x, ld] (what a bitwise comparison would compare), as 32 bytes"""
    import hashlib

    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(f[:, :, :ny]).tobytes()).digest(), dtype=np.uint8)


def generate_tg64():
    """The configuration of the reference's own golden files (graphs/fvm_*_64.txt): Taylor-Green on 64 x 64, Re = 100,
    umax = 0.01 / sqrt(3), nu = umax n / Re, fp64 -- initial condition by taylor_green_eval (t = 0) and the density conversion of
    app/main_taylor_green.f90:133-149, set_properties WITHOUT a magic number (the default branch), set_pdf_to_equilibrium, then a few
    steps of each scheme, update_macros.  A 64 x 64 DUGKS step is 1.3 million Fortran statements (two and a half minutes in the
    interpreter), so this fixture is generated on request (`--tg64`) and not regenerated by the tests; lattices are stored as
    sha256 digests, the macroscopic fields in full."""
    n = 64
    out = {}
    its = {"dugks": interp("f64", DUGKS=1), "plain": interp("f64")}
    it = its["plain"]
    umax = np.float64(0.01) / np.sqrt(np.float64(3.0))
    nu = (umax * np.float64(n)) / np.float64(100.0)
    tau = np.float64(3.0) * nu
    kx = np.float64(2) * it.constant("taylor_green", "pi") / np.float64(n)
    case = it.run("taylor_green", "taylor_green_t_constructor", n, n, kx, kx, umax, nu)
    out["params"] = np.array([n, umax, nu, tau, kx, case["td"]])
    cases = {"dugks": ("dugks", "perform_dugks_step", ("periodic_dugks", "dugks_stream"), ("periodic_dugks", "dugks_collide"), 5.0 * tau, 2,
                       "periodic_dugks"),
             "fvm_bgk": ("plain", "perform_step", ("fvm_bardow", "stream_fvm_bardow"), ("collision_bgk", "collide_bgk"), 2.0 * tau, 2, "fvm_bardow"),
             "lbm_bgk": ("plain", "perform_lbm_step", ("periodic_lbm", "lbm_stream"), ("collision_bgk", "collide_bgk"), np.float64(1.0), 3,
                         "periodic_lbm")}
    for name, (which, step_proc, streaming, collision, dt, nsteps, mod) in cases.items():
        interp_ = its[which]
        g = new_grid(interp_, n, n)
        interp_.run("taylor_green", "taylor_green_eval", case, np.float64(0.0), g["rho"].a, g["ux"].a, g["uy"].a)
        interp_.run("fvm_bardow", "set_properties", g, nu, np.float64(dt))
        g["rho"].a[...] = g["rho"].a / g["csqr"] + np.float64(1.0)  # apply_initial_condition: grid%rho = grid%rho/grid%csqr + rho0
        out[f"{name}.init"] = g["_mf"].copy()
        interp_.run("fvm_bardow", "set_pdf_to_equilibrium", g)
        g["streaming"] = lambda i_=interp_, s_=streaming, g_=g: i_.run(s_[0], s_[1], g_)
        g["collision"] = lambda i_=interp_, c_=collision, g_=g: i_.run(c_[0], c_[1], g_)
        for _ in range(nsteps):
            interp_.run(mod, step_proc, g)
        interp_.run("fvm_bardow", "update_macros", g)
        out[f"{name}.args"] = np.array([nsteps, dt])
        out[f"{name}.props"] = np.array([g["tau"], g["omega"], g["trt_magic"], g["csqr"]])
        out[f"{name}.idx"] = np.array([g["iold"], g["inew"]])
        out[f"{name}.macros"] = g["_mf"].copy()
        out[f"{name}.digests"] = np.stack([lattice_digest(g["_f"][0], n), lattice_digest(g["_f"][1], n)])
    out["statements_executed"] = np.array([sum(i.nstmt for i in its.values())])
    return out


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    if "--tg64" in sys.argv:
        data = generate_tg64()
        path = os.path.join(ROOT, "tests", "golden", "refsrc_tg64_f64.npz")
        np.savez_compressed(path, **data)
        print(path, len(data), "arrays,", int(data["statements_executed"][0]), "Fortran statements executed")
        sys.exit(0)
    for prec in ("f64", "f32"):
        data = generate(prec)
        path = os.path.join(ROOT, "tests", "golden", f"refsrc_{prec}.npz")
        np.savez_compressed(path, **data)
        print(path, len(data), "arrays,", int(data["statements_executed"][0]), "Fortran statements executed")
    data = generate_sim()
    path = os.path.join(ROOT, "tests", "golden", "refsrc_sim_f64.npz")
    np.savez_compressed(path, **data)
    print(path, len(data), "arrays,", int(data["statements_executed"][0]), "Fortran statements executed")
