"""A SECOND, independent restatement of the LBM path of the reference, in whole-array numpy (test infrastructure).

The C oracle (oracle/plbm_oracle.c) is one reading of the Fortran; the reference ships no golden data for
lbm_stream, collide_trt and collide_rr ("parity unpinned by the reference").  This module is a second
reading, written separately from the Fortran text, statement by statement and in the Fortran's evaluation order
(left to right within a precedence level, parentheses kept, every literal rounded to the working precision like
`_wp`).  numpy's elementwise add / subtract / multiply / divide are correctly rounded IEEE operations and never
fused, so two faithful readings must agree BIT FOR BIT -- tests/test_oracle_second_restatement.py checks exactly
that, in fp64 and fp32.  Arrays are f[q, x, y] with y < ny (the padding rows of the oracle's layout are ignored).

  set_properties         src/fvm_bardow.F90:242-269
  equilibrium            src/fvm_bardow.F90:99-126   (set_pdf_to_equilibrium :272-305 applies it node by node)
  update_macros_kernel   src/fvm_bardow.F90:358-386
  lbm_stream_kernel      src/periodic_lbm.f90:45-127
  bgk_kernel             src/collision_bgk.F90:35-82
  trt_naive, lambda_d    src/collision_trt.F90:13-34, 64-160
  rr_kernel_naive        src/collision_regularized.F90:11-14, 40-202
  bgk_kernel_cache       src/collision_bgk.F90:84-176 (-DSPLIT; same text as periodic_dugks' kernel_bgk)
  trt_split              src/collision_trt.F90:162-290 (-DSPLIT)
  bgk_improved_kernel    src/collision_bgk_improved.f90:21-107
  vorticity_2nd / _4th   src/vorticity.f90:13-43, 46-87   (fields are u[x, y])
  fvm_bardow_kernel      src/fvm_bardow.F90:410-507   (square grids: the shipped loop bounds are swapped)
  fdm_bardow_kernel      src/fvm_bardow.F90:526-682   (default build and -DFDM_WLS / _GAUSS_V1 / _GAUSS_V2 / -DFDM_ISO)
  fdm_sofonea_kernel     src/fvm_bardow.F90:702-880
  periodic_dugks         src/periodic_dugks.F90:25-77 (step, dugks_collide), 80-169 (kernel_bgk), 172-304 (kernel_stream),
                         310-434 (update_ew / update_ns); built with -DDUGKS
  sim/ lw plugin         sim/sim_lw.F90:24-85 (lw_stream), 87-164 (lw_collision), 166-182 (lw_bc = periodic images)
  sim/ seam (slbm)       sim/sim.F90:119-131, 148-199, 352-383, 404-505, 568-624   (DDF-shifted populations f[k, j, i])
"""
import numpy as np

CX = (0, 1, 0, -1, 0, 1, -1, -1, 1)  # src/fvm_bardow.F90:87-88
CY = (0, 0, 1, 0, -1, 1, 1, -1, -1)


def set_properties(T, nu, dt, magic=None):
    nu, dt = T(nu), T(dt)
    csqr = T(1) / T(3)
    invcsqr = T(1) / csqr
    tau = invcsqr * nu
    omega = dt / (tau + T(0.5) * dt)
    trt_magic = T(magic) if magic is not None else (tau / dt) * (tau / dt)  # (tau/dt)**2: gfortran emits x*x
    return dict(tau=tau, omega=omega, trt_magic=trt_magic, csqr=csqr)


def lbm_stream(f):
    """fdst(y,x,q) = fsrc(y - cy_q, x - cx_q, q) with periodic neighbours (xm1, xp1, ym1, yp1 of the Fortran)."""
    out = np.empty_like(f)
    for q in range(9):
        out[q] = np.roll(f[q], shift=(CX[q], CY[q]), axis=(0, 1))
    return out


def equilibrium(T, rho, ux, uy):
    w0, ws, wd = T(4) / T(9), T(1) / T(9), T(1) / T(36)
    uxx = ux * ux
    uyy = uy * uy
    indp = T(1) - T(1.5) * (uxx + uyy)
    feq = [None] * 9
    feq[0] = w0 * rho * indp
    feq[1] = ws * rho * (indp + T(3) * ux + T(4.5) * uxx)
    feq[2] = ws * rho * (indp + T(3) * uy + T(4.5) * uyy)
    feq[3] = ws * rho * (indp - T(3) * ux + T(4.5) * uxx)
    feq[4] = ws * rho * (indp - T(3) * uy + T(4.5) * uyy)
    uxpy = ux + uy
    feq[5] = wd * rho * (indp + T(3) * uxpy + T(4.5) * uxpy * uxpy)
    feq[7] = wd * rho * (indp - T(3) * uxpy + T(4.5) * uxpy * uxpy)
    uxmy = ux - uy
    feq[6] = wd * rho * (indp - T(3) * uxmy + T(4.5) * uxmy * uxmy)
    feq[8] = wd * rho * (indp + T(3) * uxmy + T(4.5) * uxmy * uxmy)
    return feq


def update_macros(f):
    T = f.dtype.type
    fs = f
    rho = fs[0] + (((fs[5] + fs[7]) + (fs[6] + fs[8])) + ((fs[1] + fs[3]) + (fs[2] + fs[4])))
    invrho = T(1) / rho
    ux = invrho * (((fs[5] - fs[7]) + (fs[8] - fs[6])) + (fs[1] - fs[3]))
    uy = invrho * (((fs[5] - fs[7]) + (fs[6] - fs[8])) + (fs[2] - fs[4]))
    return rho, ux, uy


def collide_bgk(f, omega):
    T = f.dtype.type
    omega = T(omega)
    omegabar = T(1) - omega
    fs = f
    rho = (((fs[5] + fs[7]) + (fs[6] + fs[8])) + ((fs[1] + fs[3]) + (fs[2] + fs[4]))) + fs[0]
    invrho = T(1) / rho
    ux = invrho * (((fs[5] - fs[7]) + (fs[8] - fs[6])) + (fs[1] - fs[3]))
    uy = invrho * (((fs[5] - fs[7]) + (fs[6] - fs[8])) + (fs[2] - fs[4]))
    feq = equilibrium(T, rho, ux, uy)
    return np.stack([omegabar * fs[q] + omega * feq[q] for q in range(9)])


def lambda_d(T, omega, x):
    omega, x = T(omega), T(x)
    return (T(4) - T(2) * omega) / (T(4) * x * omega + T(2) - omega)


def collide_trt(f, omega, magic):
    T = f.dtype.type
    t0 = T(4) / T(9)
    t1x2 = (T(1) / T(9)) * T(2)
    t2x2 = (T(1) / T(36)) * T(2)
    inv2csq2 = T(1) / (T(2) * (T(1) / T(3)) * (T(1) / T(3)))
    fac1 = t1x2 * inv2csq2
    fac2 = t2x2 * inv2csq2
    lam_e = T(omega)
    lam_d = lambda_d(T, omega, magic)
    les = T(0.5) * lam_e
    lds = T(0.5) * lam_d
    vC, vE, vN, vW, vS, vNE, vNW, vSW, vSE = (f[q] for q in range(9))
    rho = (((vNE + vSW) + (vNW + vSE)) + ((vE + vW) + (vN + vS))) + vC
    velX = ((vNE - vSW) + (vSE - vNW)) + (vE - vW)
    velY = ((vNE - vSW) + (vNW - vSE)) + (vN - vS)
    velX2 = velX * velX
    velY2 = velY * velY
    feq_common = rho - T(1.5) * (velX2 + velY2)
    out = [None] * 9
    out[0] = vC * (T(1) - lam_e) + lam_e * t0 * feq_common
    velXPY = velX + velY
    sym = les * (vNE + vSW - fac2 * velXPY * velXPY - t2x2 * feq_common)
    asym = lds * (vNE - vSW - T(3) * t2x2 * velXPY)
    out[5] = vNE - sym - asym
    out[7] = vSW - sym + asym
    velXMY = velX - velY
    sym = les * (vSE + vNW - fac2 * velXMY * velXMY - t2x2 * feq_common)
    asym = lds * (vSE - vNW - T(3) * t2x2 * velXMY)
    out[8] = vSE - sym - asym
    out[6] = vNW - sym + asym
    sym = les * (vN + vS - fac1 * velY2 - t1x2 * feq_common)
    asym = lds * (vN - vS - T(3) * t1x2 * velY)
    out[2] = vN - sym - asym
    out[4] = vS - sym + asym
    sym = les * (vE + vW - fac1 * velX2 - t1x2 * feq_common)
    asym = lds * (vE - vW - T(3) * t1x2 * velX)
    out[1] = vE - sym - asym
    out[3] = vW - sym + asym
    return np.stack(out)


def collide_rr(f, omega):
    T = f.dtype.type
    w0, ws, wd, csqr = T(4) / T(9), T(1) / T(9), T(1) / T(36), T(1) / T(3)
    omega = T(omega)
    omega_w0 = w0 * (T(1) - omega)
    omega_ws = ws * (T(1) - omega)
    omega_wd = wd * (T(1) - omega)
    vC, vE, vN, vW, vS, vNE, vNW, vSW, vSE = (f[q] for q in range(9))
    rho = (((vNE + vSW) + (vNW + vSE)) + ((vE + vW) + (vN + vS))) + vC
    invrho = T(1) / rho
    ux = invrho * (((vNE - vSW) + (vSE - vNW)) + (vE - vW))
    uy = invrho * (((vNE - vSW) + (vNW - vSE)) + (vN - vS))
    uxx = ux * ux
    uyy = uy * uy
    uxxy = uxx * uy
    uyyx = uyy * ux
    uxxyy = uxx * uyy
    indp0 = T(1) - T(1.5) * (uxx + uyy)
    indps = indp0 - T(4.5) * uxxyy
    indpd = indp0 + T(9) * uxxyy
    indp0 = indp0 + T(2.25) * uxxyy
    feq = [None] * 9
    feq[0] = w0 * rho * indp0
    feq[1] = ws * rho * (indps + T(3) * ux + T(4.5) * (uxx - uyyx))
    feq[3] = ws * rho * (indps - T(3) * ux + T(4.5) * (uxx + uyyx))
    feq[2] = ws * rho * (indps + T(3) * uy + T(4.5) * (uyy - uxxy))
    feq[4] = ws * rho * (indps - T(3) * uy + T(4.5) * (uyy + uxxy))
    vC = vC - feq[0]
    vE = vE - feq[1]
    vN = vN - feq[2]
    vW = vW - feq[3]
    vS = vS - feq[4]
    axx = csqr * (T(2) * (vE + vW) - (vN + vS) - vC)
    ayy = csqr * (T(2) * (vN + vS) - (vE + vW) - vC)
    u3p = uxxy + uyyx
    uxpy = ux + uy
    indp57 = indpd + T(4.5) * uxpy * uxpy
    feq[5] = wd * rho * (indp57 + T(3) * uxpy + T(9) * u3p)
    feq[7] = wd * rho * (indp57 - T(3) * uxpy - T(9) * u3p)
    u3m = uxxy - uyyx
    uxmy = ux - uy
    indp68 = indpd + T(4.5) * uxmy * uxmy
    feq[6] = wd * rho * (indp68 - T(3) * uxmy + T(9) * u3m)
    feq[8] = wd * rho * (indp68 + T(3) * uxmy - T(9) * u3m)
    vNE = vNE - feq[5]
    vNW = vNW - feq[6]
    vSW = vSW - feq[7]
    vSE = vSE - feq[8]
    tmp = T(2) * csqr * (vNE + vNW + vSW + vSE)
    axx = axx + tmp
    ayy = ayy + tmp
    axy = (vNE + vSW) - (vNW + vSE)
    axxy = T(2) * ux * axy + uy * axx
    ayyx = T(2) * uy * axy + ux * ayy
    axxyy = T(2) * (ux * ayyx + uy * axxy) - uxx * ayy - uyy * axx - T(4) * ux * uy * axy
    indp0 = T(-1.5) * (axx + ayy)
    indps = indp0 - T(4.5) * axxyy
    indpd = T(9) * axxyy - T(2) * indp0
    indp0 = indp0 + T(2.25) * axxyy
    vC = indp0
    vE = indps + T(4.5) * (axx - ayyx)
    vW = indps + T(4.5) * (axx + ayyx)
    vN = indps + T(4.5) * (ayy - axxy)
    vS = indps + T(4.5) * (ayy + axxy)
    vNE = indpd + T(9) * (axxy + ayyx + axy)
    vSW = indpd - T(9) * (axxy + ayyx - axy)
    vNW = indpd + T(9) * (axxy - ayyx - axy)
    vSE = indpd - T(9) * (axxy - ayyx + axy)
    out = [feq[0] + omega_w0 * vC, feq[1] + omega_ws * vE, feq[2] + omega_ws * vN, feq[3] + omega_ws * vW, feq[4] + omega_ws * vS,
           feq[5] + omega_wd * vNE, feq[6] + omega_wd * vNW, feq[7] + omega_wd * vSW, feq[8] + omega_wd * vSE]
    return np.stack(out)


def vorticity_2nd(ux, uy):
    T = ux.dtype.type
    duydx = T(0.5) * (np.roll(uy, -1, axis=0) - np.roll(uy, 1, axis=0))  # uy(y,xp1) - uy(y,xm1)
    duxdy = T(0.5) * (np.roll(ux, -1, axis=1) - np.roll(ux, 1, axis=1))  # ux(yp1,x) - ux(ym1,x)
    return duydx - duxdy


def vorticity_4th(ux, uy):
    """as shipped: t1 = 1/12 weighs the +-1 neighbours and t2 = 2/3 the (m2 - p2) pair"""
    T = ux.dtype.type
    t1, t2 = T(1) / T(12), T(2) / T(3)
    duydx = t1 * (np.roll(uy, -1, axis=0) - np.roll(uy, 1, axis=0)) + t2 * (np.roll(uy, 2, axis=0) - np.roll(uy, -2, axis=0))
    duxdy = t1 * (np.roll(ux, -1, axis=1) - np.roll(ux, 1, axis=1)) + t2 * (np.roll(ux, 2, axis=1) - np.roll(ux, -2, axis=1))
    return duydx - duxdy


# ---- the sim/ plugin seam (sim/sim_slbm.F90 over sim/sim.F90): populations are stored shifted by their weight -----
def sim_equilibrium(rho, ux, uy):
    """equilibrium() with ddf_shift = .true. (sim/sim.F90:352-383)"""
    T = rho.dtype.type
    w0, ws, wd = T(4) / T(9), T(1) / T(9), T(1) / T(36)
    w = (w0, ws, ws, ws, ws, wd, wd, wd, wd)
    rho0 = T(1)
    uxx = ux * ux
    uyy = uy * uy
    uxpy = ux + uy
    uxmy = ux - uy
    indp = T(-1.5) * (uxx + uyy)
    feq = [None] * 9
    feq[0] = w0 * rho * indp
    feq[1] = ws * rho * (indp + T(3) * ux + T(4.5) * uxx)
    feq[2] = ws * rho * (indp + T(3) * uy + T(4.5) * uyy)
    feq[3] = ws * rho * (indp - T(3) * ux + T(4.5) * uxx)
    feq[4] = ws * rho * (indp - T(3) * uy + T(4.5) * uyy)
    feq[5] = wd * rho * (indp + T(3) * uxpy + T(4.5) * uxpy * uxpy)
    feq[7] = wd * rho * (indp - T(3) * uxpy + T(4.5) * uxpy * uxpy)
    feq[6] = wd * rho * (indp - T(3) * uxmy + T(4.5) * uxmy * uxmy)
    feq[8] = wd * rho * (indp + T(3) * uxmy + T(4.5) * uxmy * uxmy)
    return [feq[k] + w[k] * (rho - rho0) for k in range(9)]


def sim_eqinit(p, u, v):
    """lbm_eqinit_fields (sim/sim.F90:181-199): rho = rho0 + p / csqr, f = equilibrium(rho, u, v); fields [j, i]"""
    T = p.dtype.type
    csqr = T(1) / T(3)
    rho = T(1) + p / csqr
    return np.stack(sim_equilibrium(rho, u, v))


def sim_macros(f):
    """lbm_macros (sim/sim.F90:148-179)"""
    T = f.dtype.type
    rho = (((f[5] + f[7]) + (f[6] + f[8])) + ((f[1] + f[3]) + (f[2] + f[4]))) + f[0]
    rho = rho + T(1)
    u = (((f[5] - f[7]) + (f[8] - f[6])) + (f[1] - f[3])) / rho
    v = (((f[5] - f[7]) + (f[6] - f[8])) + (f[2] - f[4])) / rho
    return rho, u, v


def sim_step(f, omega):
    """lbm_collide_and_stream_fused with push = .true. followed by lbm_periodic_bc_push (sim/sim.F90:404-505,
    568-624): collide, scatter to (i + cx, j + cy), fold the halo ring back = a periodic shift of the interior."""
    T = f.dtype.type
    omega = T(omega)
    rho, ux, uy = sim_macros(f)
    feq = sim_equilibrium(rho, ux, uy)
    out = np.empty_like(f)
    for k in range(9):
        post = omega * (feq[k] - f[k]) + f[k]
        out[k] = np.roll(post, shift=(CY[k], CX[k]), axis=(0, 1))  # axes (j, i)
    return out


def stream_fvm_bardow(f, dt):
    """fvm_bardow_kernel: bilinear face values from the 3 x 3 neighbourhood, conservative update; f[q, x, y]"""
    T = f.dtype.type
    dt = T(dt)
    p2, p8 = T(0.5), T(0.125)
    out = np.empty_like(f)
    out[0] = f[0]
    sh = lambda a, dx, dy: np.roll(a, shift=(-dx, -dy), axis=(0, 1))  # noqa: E731  value at (x + dx, y + dy)
    for q in range(1, 9):
        cxq = dt * T(CX[q])
        cyq = dt * T(CY[q])
        fc = f[q]
        fe, fn, fw, fs = sh(fc, 1, 0), sh(fc, 0, 1), sh(fc, -1, 0), sh(fc, 0, -1)
        fne, fnw, fsw, fse = sh(fc, 1, 1), sh(fc, -1, 1), sh(fc, -1, -1), sh(fc, 1, -1)
        cfw = p2 * (fc + fw) - p2 * cxq * (fc - fw) - p8 * cyq * (fnw + fn - fsw - fs)
        cfn = p2 * (fc + fn) - p2 * cyq * (fn - fc) - p8 * cxq * (fne + fe - fnw - fw)
        cfe = p2 * (fc + fe) - p2 * cxq * (fe - fc) - p8 * cyq * (fne + fn - fse - fs)
        cfs = p2 * (fc + fs) - p2 * cyq * (fc - fs) - p8 * cxq * (fse + fe - fsw - fw)
        out[q] = fc - cxq * (cfe - cfw) - cyq * (cfn - cfs)
    return out


# ---- DUGKS (src/periodic_dugks.F90, -DDUGKS) ----------------------------------------------------------------------
def _dugks_moments(f):
    T = f[0].dtype.type
    rho = (((f[5] + f[7]) + (f[6] + f[8])) + ((f[1] + f[3]) + (f[2] + f[4]))) + f[0]
    invrho = T(1) / rho
    ux = invrho * (((f[5] - f[7]) + (f[8] - f[6])) + (f[1] - f[3]))
    uy = invrho * (((f[5] - f[7]) + (f[6] - f[8])) + (f[2] - f[4]))
    indp = T(1) / T(3) - T(0.5) * (ux * ux + uy * uy)
    return rho, ux, uy, indp


def _dugks_relax(f, omega, which):
    """kernel_bgk (which = 'all'), update_ew ('ew': populations with cx != 0), update_ns ('ns': cy != 0); f is a list of 9"""
    T = f[0].dtype.type
    omega = T(omega)
    w0, ws, wd = T(4) / T(9), T(1) / T(9), T(1) / T(36)
    omegabar = T(1) - omega
    omega_w0 = T(3) * omega * w0
    omega_ws = T(3) * omega * ws
    omega_wd = T(3) * omega * wd
    rho, ux, uy, indp = _dugks_moments(f)
    g = list(f)
    if which == "all":
        g[0] = omegabar * f[0] + omega_w0 * rho * indp
    if which in ("all", "ew"):
        t13 = indp + T(1.5) * ux * ux
        g[1] = omegabar * f[1] + omega_ws * rho * (t13 + ux)
        g[3] = omegabar * f[3] + omega_ws * rho * (t13 - ux)
    if which in ("all", "ns"):
        t24 = indp + T(1.5) * uy * uy
        g[2] = omegabar * f[2] + omega_ws * rho * (t24 + uy)
        g[4] = omegabar * f[4] + omega_ws * rho * (t24 - uy)
    velxpy = ux + uy
    t57 = indp + T(1.5) * velxpy * velxpy
    g[5] = omegabar * f[5] + omega_wd * rho * (t57 + velxpy)
    g[7] = omegabar * f[7] + omega_wd * rho * (t57 - velxpy)
    velxmy = ux - uy
    t68 = indp + T(1.5) * velxmy * velxmy
    g[6] = omegabar * f[6] + omega_wd * rho * (t68 - velxmy)
    g[8] = omegabar * f[8] + omega_wd * rho * (t68 + velxmy)
    return g


def dugks_step(ftilde, grid_omega, tau, dt):
    """perform_dugks_step: dugks_collide, dugks_stream (the swap is the caller's).  Returns (ftilde_new, fbar_plus):
    what the reference leaves in lattice `inew` / `iold` BEFORE the swap, i.e. `iold` / `inew` after it."""
    T = ftilde.dtype.type
    tau, dt = T(tau), T(dt)
    tau_d = tau / dt
    omega = T(1) / (tau_d + T(0.5))
    fnew = _dugks_relax([ftilde[q] for q in range(9)], grid_omega, "all")     # ftilde^+ (full step, grid%omega)
    omega = T(0.75) * omega
    fold = _dugks_relax([ftilde[q] for q in range(9)], omega, "all")          # fbar^+ (half step)
    omega_f = T(1) / (T(4) * tau_d + T(1))
    p2, p8 = T(0.5), T(0.125)
    sh = lambda a, dx, dy: np.roll(a, shift=(-dx, -dy), axis=(0, 1))  # noqa: E731  value at (x + dx, y + dy)
    cfw, cfn, cfe, cfs = [None] * 9, [None] * 9, [None] * 9, [None] * 9
    for q in range(9):
        cxq = dt * T(CX[q])
        cyq = dt * T(CY[q])
        fc = fold[q]
        fe, fn, fw, fs = sh(fc, 1, 0), sh(fc, 0, 1), sh(fc, -1, 0), sh(fc, 0, -1)
        fne, fnw, fsw, fse = sh(fc, 1, 1), sh(fc, -1, 1), sh(fc, -1, -1), sh(fc, 1, -1)
        cfw[q] = p2 * (fc + fw) - p2 * cxq * (fc - fw) - p8 * cyq * (fnw + fn - fsw - fs)
        cfn[q] = p2 * (fc + fn) - p2 * cyq * (fn - fc) - p8 * cxq * (fne + fe - fnw - fw)
        cfe[q] = p2 * (fc + fe) - p2 * cxq * (fe - fc) - p8 * cyq * (fne + fn - fse - fs)
        cfs[q] = p2 * (fc + fs) - p2 * cyq * (fc - fs) - p8 * cxq * (fse + fe - fsw - fw)
    cfw = _dugks_relax(cfw, omega_f, "ew")
    cfe = _dugks_relax(cfe, omega_f, "ew")
    cfn = _dugks_relax(cfn, omega_f, "ns")
    cfs = _dugks_relax(cfs, omega_f, "ns")
    out = [fnew[0]]
    for q in range(1, 9):
        cxq = dt * T(CX[q])
        cyq = dt * T(CY[q])
        out.append(fnew[q] - cxq * (cfe[q] - cfw[q]) - cyq * (cfn[q] - cfs[q]))
    return np.stack(out), np.stack(fold)


def _neighbours(fc):
    sh = lambda a, dx, dy: np.roll(a, shift=(-dx, -dy), axis=(0, 1))  # noqa: E731  value at (x + dx, y + dy)
    return (sh(fc, 1, 0), sh(fc, 0, 1), sh(fc, -1, 0), sh(fc, 0, -1), sh(fc, 1, 1), sh(fc, -1, 1), sh(fc, -1, -1), sh(fc, 1, -1))


def stream_fdm_bardow(f, dt, stencil="default"):
    """fdm_bardow_kernel: second-order Taylor (Lax-Wendroff) streaming; `stencil` = the cpp macro of the build"""
    T = f.dtype.type
    dt = T(dt)
    p2 = T(0.5)
    two_thirds, one_sixth = T(2) / T(3), T(1) / T(6)
    five_sixths, one_twelth = T(10) / T(12), T(1) / T(12)
    one_third = T(1) / T(3)
    out = np.empty_like(f)
    out[0] = f[0]
    for q in range(1, 9):
        cxq = dt * T(CX[q])
        cyq = dt * T(CY[q])
        cxxq = T(0.5) * cxq * cxq
        cyyq = T(0.5) * cyq * cyq
        cxyq = cxq * cyq
        fc = f[q]
        fe, fn, fw, fs, fne, fnw, fsw, fse = _neighbours(fc)
        if stencil == "wls":
            dfx = one_sixth * ((fne - fnw) + (fe - fw) + (fse - fsw))
            dfy = one_sixth * ((fne - fse) + (fn - fs) + (fnw - fsw))
            dfxx = one_third * (fne - T(2) * fn + fnw) + one_third * (fe - T(2) * fc + fw) + one_third * (fse - T(2) * fs + fsw)
            dfyy = one_third * (fne - T(2) * fe + fse) + one_third * (fn - T(2) * fc + fs) + one_third * (fnw - T(2) * fw + fsw)
            dfxy = T(0.25) * (fne - fnw + fsw - fse)
        elif stencil in ("wls_gauss_v1", "wls_gauss_v2"):
            if stencil == "wls_gauss_v1":
                p1s, p1d = T(0.2880584423829145035434), T(0.1059707788085427065949)
                p2c, p2d1 = T(-1.152233769531658458263), T(0.5761168847658292291314)
                p2d2, p2d = T(-0.4238831152341712149578), T(0.2119415576170855242122)
            else:
                p1s, p1d = T(0.3934930210807994210853), T(0.05325348945960039354075)
                p2c, p2d1 = T(-1.573972084323197018207), T(0.7869860421615988421706)
                p2d2, p2d = T(-0.2130139578384016019186), T(0.1065069789192007732037)
            dfx = p1s * (fe - fw) + p1d * (fne - fnw) + p1d * (fse - fsw)
            dfy = p1s * (fn - fs) + p1d * (fne - fse) + p1d * (fnw - fsw)
            dfxx = p2c * fc + p2d1 * (fe + fw) + p2d2 * (fn + fs) + p2d * (fne + fnw + fsw + fse)
            dfyy = p2c * fc + p2d2 * (fe + fw) + p2d1 * (fn + fs) + p2d * (fne + fnw + fsw + fse)
            dfxy = T(0.25) * (fne - fnw + fsw - fse)
        elif stencil == "iso":
            dfx = p2 * (one_sixth * (fne - fnw) + two_thirds * (fe - fw) + one_sixth * (fse - fsw))
            dfy = p2 * (one_sixth * (fne - fse) + two_thirds * (fn - fs) + one_sixth * (fnw - fsw))
            dfxx = one_twelth * (fne - T(2) * fn + fnw) + five_sixths * (fe - T(2) * fc + fw) + one_twelth * (fse - T(2) * fs + fsw)
            dfyy = one_twelth * (fne - T(2) * fe + fse) + five_sixths * (fn - T(2) * fc + fs) + one_twelth * (fnw - T(2) * fw + fsw)
            dfxy = T(0.25) * (fne - fse - fnw + fsw)
        else:
            dfx = p2 * (fe - fw)
            dfy = p2 * (fn - fs)
            dfxx = fe - T(2) * fc + fw
            dfyy = fn - T(2) * fc + fs
            dfxy = T(0.25) * (fne - fse - fnw + fsw)
        out[q] = fc - cxq * dfx - cyq * dfy + (cxxq * dfxx + cxyq * dfxy + cyyq * dfyy)
    return out


def stream_fdm_sofonea(f, dt):
    """fdm_sofonea_kernel: one-dimensional Lax-Wendroff along each lattice direction (square grids)"""
    T = f.dtype.type
    dt = T(dt)
    p2 = T(0.5) / np.sqrt(T(2))
    sh = lambda a, dx, dy: np.roll(a, shift=(-dx, -dy), axis=(0, 1))  # noqa: E731
    out = np.empty_like(f)
    out[0] = f[0]
    for q in range(1, 9):
        fc = f[q]
        fu = sh(fc, CX[q], CY[q])
        fd = sh(fc, -CX[q], -CY[q])
        if q <= 4:
            du1 = T(0.5) * (fu - fd)
            du2 = fu - T(2) * fc + fd
        else:
            du1 = p2 * (fu - fd)
            du2 = T(0.5) * (fu - T(2) * fc + fd)
        out[q] = fc + dt * (T(0.5) * dt * du2 - du1)
    return out


def collide_bgk_split(f, omega):
    """bgk_kernel_cache (-DSPLIT): the re-associated BGK form, identical in text to periodic_dugks' kernel_bgk"""
    return np.stack(_dugks_relax([f[q] for q in range(9)], omega, "all"))


def collide_trt_split(f, omega, magic):
    """trt_split (-DSPLIT): like trt_naive but fac1 * vel * vel is evaluated left to right (no vel2 temporaries)"""
    T = f.dtype.type
    t0 = T(4) / T(9)
    t1x2 = (T(1) / T(9)) * T(2)
    t2x2 = (T(1) / T(36)) * T(2)
    inv2csq2 = T(1) / (T(2) * (T(1) / T(3)) * (T(1) / T(3)))
    fac1 = t1x2 * inv2csq2
    fac2 = t2x2 * inv2csq2
    lam_e = T(omega)
    lam_d = lambda_d(T, omega, magic)
    les = T(0.5) * lam_e
    lds = T(0.5) * lam_d
    vC, vE, vN, vW, vS, vNE, vNW, vSW, vSE = (f[q] for q in range(9))
    rho = (((vNE + vSW) + (vNW + vSE)) + ((vE + vW) + (vN + vS))) + vC
    velX = ((vNE - vSW) + (vSE - vNW)) + (vE - vW)
    velY = ((vNE - vSW) + (vNW - vSE)) + (vN - vS)
    feq_common = rho - T(1.5) * (velX * velX + velY * velY)
    out = [None] * 9
    out[0] = vC * (T(1) - lam_e) + lam_e * t0 * feq_common
    velXPY = velX + velY
    sym = les * (vNE + vSW - fac2 * velXPY * velXPY - t2x2 * feq_common)
    asym = lds * (vNE - vSW - T(3) * t2x2 * velXPY)
    out[5] = vNE - sym - asym
    out[7] = vSW - sym + asym
    velXMY = velX - velY
    sym = les * (vSE + vNW - fac2 * velXMY * velXMY - t2x2 * feq_common)
    asym = lds * (vSE - vNW - T(3) * t2x2 * velXMY)
    out[8] = vSE - sym - asym
    out[6] = vNW - sym + asym
    sym = les * (vN + vS - fac1 * velY * velY - t1x2 * feq_common)
    asym = lds * (vN - vS - T(3) * t1x2 * velY)
    out[2] = vN - sym - asym
    out[4] = vS - sym + asym
    sym = les * (vE + vW - fac1 * velX * velX - t1x2 * feq_common)
    asym = lds * (vE - vW - T(3) * t1x2 * velX)
    out[1] = vE - sym - asym
    out[3] = vW - sym + asym
    return np.stack(out)


def collide_bgk_improved(f, omega):
    T = f.dtype.type
    omega = T(omega)
    one_third, two_thirds = T(1) / T(3), T(2) / T(3)
    fac = T(4.5) - T(2.25) * omega
    omegabar = T(1) - omega
    vC, vE, vN, vW, vS, vNE, vNW, vSW, vSE = (f[q] for q in range(9))
    rho = (((vNE + vSW) + (vNW + vSE)) + ((vE + vW) + (vN + vS))) + vC
    invrho = T(1) / rho
    sumX1 = vE + vNE + vSE
    sumXN = vW + vNW + vSW
    sumY1 = vN + vNE + vNW
    sumYN = vS + vSE + vSW
    m10 = invrho * (sumX1 - sumXN)
    m01 = invrho * (sumY1 - sumYN)
    u2 = m10 * m10
    v2 = m01 * m01
    m20 = invrho * (sumX1 + sumXN)
    m02 = invrho * (sumY1 + sumYN)
    Gx = fac * u2 * (m20 - one_third - u2)
    Gy = fac * v2 * (m02 - one_third - v2)
    X0 = -two_thirds + u2 + Gx
    X1 = -(X0 + T(1) + m10) * T(0.5)
    XN = X1 + m10
    Y0 = -two_thirds + v2 + Gy
    Y1 = -(Y0 + T(1) + m01) * T(0.5)
    YN = Y1 + m01
    rho_omega = rho * omega
    X0 = X0 * rho_omega
    X1 = X1 * rho_omega
    XN = XN * rho_omega
    out = [omegabar * vC + X0 * Y0, omegabar * vE + X1 * Y0, omegabar * vN + X0 * Y1, omegabar * vW + XN * Y0, omegabar * vS + X0 * YN,
           omegabar * vNE + X1 * Y1, omegabar * vNW + XN * Y1, omegabar * vSW + XN * YN, omegabar * vSE + X1 * YN]
    return np.stack(out)


# ---- the 2nd-order Lax-Wendroff plugin of sim/ (sim/sim_lw.F90), populations DDF-shifted, fields f[k, j, i] --------------
def lw_stream(f, dt):
    T = f.dtype.type
    dt = T(dt)
    at = lambda a, di, dj: np.roll(a, shift=(-dj, -di), axis=(0, 1))  # noqa: E731  value at (i + di, j + dj)
    out = np.empty_like(f)
    out[0] = f[0]
    for k in range(1, 9):
        vx = dt * T(CX[k])
        vy = dt * T(CY[k])
        vxx = T(0.5) * vx * vx
        vyy = T(0.5) * vy * vy
        vxy = vx * vy
        c = f[k]
        dfx = T(0.5) * (at(c, 1, 0) - at(c, -1, 0))
        dfy = T(0.5) * (at(c, 0, 1) - at(c, 0, -1))
        dfxx = at(c, 1, 0) - T(2) * c + at(c, -1, 0)
        dfyy = at(c, 0, 1) - T(2) * c + at(c, 0, -1)
        dfxy = T(0.25) * (at(c, 1, 1) - at(c, -1, 1) + at(c, -1, -1) - at(c, 1, -1))
        out[k] = c - vx * dfx - vy * dfy + (vxx * dfxx + vxy * dfxy + vyy * dfyy)
    return out


def lw_collision(f, omega):
    T = f.dtype.type
    omega = T(omega)
    rho = f[0] + (((f[5] + f[7]) + (f[6] + f[8])) + ((f[1] + f[3]) + (f[2] + f[4]))) + T(1)
    irho = T(1) / rho
    ux = (((f[5] - f[7]) + (f[8] - f[6])) + (f[1] - f[3])) * irho
    uy = (((f[5] - f[7]) + (f[6] - f[8])) + (f[2] - f[4])) * irho
    feq = sim_equilibrium(rho, ux, uy)
    return np.stack([f[k] + omega * (feq[k] - f[k]) for k in range(9)])
