/*
 * plbm_oracle_impl.h -- TEST INFRASTRUCTURE ONLY (CPU oracle), never shipped.
 *
 * Precision-generic body of the oracle.  Included twice by plbm_oracle.c with
 *   REAL = double, SFX(x) = x##_f64      (reference default build)
 *   REAL = float,  SFX(x) = x##_f32      (reference -DPRECISION_SINGLE build,
 *                                         src/precision.F90:11-15)
 *
 * Every function is a literal restatement (same loop nests, same
 * parenthesisation, same evaluation order) of one Fortran routine of
 * ivan-pi/periodic-lbm; the routine is cited above each function as
 * file:line relative to the reference root.  The translation unit must be
 * compiled with -ffp-contract=off so no FMA contraction changes the rounding
 * (gfortran -O3 on baseline x86-64 emits none either).
 *
 * Array layout is the reference's: f(ld,nx,0:8) column-major, i.e.
 *   f[y + ld*(x + nx*q)],  0-based y in [0,ny), x in [0,nx), q in [0,9)
 * (src/fvm_bardow.F90:144-153), macroscopic fields are (ny,nx) unpadded.
 */

#define R(x) ((REAL)(x))
/* periodic neighbours: the reference writes mod(x,nx)+1 and mod(nx+x-2,nx)+1 on 1-based
 * indices (e.g. src/periodic_lbm.f90:63-64); same integers, without the division. */
#define WRAP_P1(i, n) ((i) + 1 == (n) ? 0 : (i) + 1)
#define WRAP_M1(i, n) ((i) == 0 ? (n)-1 : (i)-1)

/* src/fvm_bardow.F90:90-95 (and the identical parameter blocks in
 * collision_bgk.F90:11-13, collision_regularized.F90:11-14,
 * periodic_dugks.F90:17-21).  Evaluated in working precision. */
#define W0 (R(4.0) / R(9.0))
#define WS (R(1.0) / R(9.0))
#define WD (R(1.0) / R(36.0))
#define CSQR (R(1.0) / R(3.0))
#define INVCSQR (R(1.0) / CSQR)
#define ONE_THIRD (R(1.0) / R(3.0))

#define FIDX(y, x, q) ((size_t)(y) + (size_t)ld * ((size_t)(x) + (size_t)nx * (size_t)(q)))
#define MIDX(y, x) ((size_t)(y) + (size_t)ny * (size_t)(x))

/* velocity set, src/fvm_bardow.F90:87-88 */
static const int SFX(ocx)[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
static const int SFX(ocy)[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};

/* ------------------------------------------------------------------ */
/* src/fvm_bardow.F90:99-126  equilibrium(rho,ux,uy)                   */
void SFX(orc_equilibrium)(REAL rho, REAL ux, REAL uy, REAL *feq)
{
    REAL uxx = ux * ux;
    REAL uyy = uy * uy;
    REAL indp = R(1.0) - R(1.5) * (uxx + uyy);

    feq[0] = W0 * rho * (indp);
    feq[1] = WS * rho * (indp + R(3.0) * ux + R(4.5) * uxx);
    feq[2] = WS * rho * (indp + R(3.0) * uy + R(4.5) * uyy);
    feq[3] = WS * rho * (indp - R(3.0) * ux + R(4.5) * uxx);
    feq[4] = WS * rho * (indp - R(3.0) * uy + R(4.5) * uyy);

    REAL uxpy = ux + uy;
    feq[5] = WD * rho * (indp + R(3.0) * uxpy + R(4.5) * uxpy * uxpy);
    feq[7] = WD * rho * (indp - R(3.0) * uxpy + R(4.5) * uxpy * uxpy);

    REAL uxmy = ux - uy;
    feq[6] = WD * rho * (indp - R(3.0) * uxmy + R(4.5) * uxmy * uxmy);
    feq[8] = WD * rho * (indp + R(3.0) * uxmy + R(4.5) * uxmy * uxmy);
}

/* src/fvm_bardow.F90:242-269  set_properties(grid,nu,dt[,magic])
 * out[0]=tau, out[1]=omega, out[2]=trt_magic, out[3]=csqr */
void SFX(orc_set_properties)(REAL nu, REAL dt, REAL magic, int has_magic, REAL *out)
{
    REAL tau = INVCSQR * nu;
    out[0] = tau;
    out[1] = dt / (tau + R(0.5) * dt);
    if (has_magic)
        out[2] = magic;
    else
        out[2] = (tau / dt) * (tau / dt);
    out[3] = CSQR;
}

/* src/fvm_bardow.F90:272-305  set_pdf_to_equilibrium (writes lattice iold) */
void SFX(orc_set_pdf_to_equilibrium)(int nx, int ny, int ld, const REAL *rho, const REAL *ux,
                                     const REAL *uy, REAL *pdf)
{
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y) {
            REAL feq[9];
            SFX(orc_equilibrium)(rho[MIDX(y, x)], ux[MIDX(y, x)], uy[MIDX(y, x)], feq);
            for (int q = 0; q < 9; ++q) pdf[FIDX(y, x, q)] = feq[q];
        }
}

/* src/fvm_bardow.F90:343-390  update_macros_kernel */
void SFX(orc_update_macros)(int nx, int ny, int ld, const REAL *f, REAL *grho, REAL *gux, REAL *guy)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y) {
            REAL fs[9];
            for (int q = 0; q < 9; ++q) fs[q] = f[FIDX(y, x, q)];
            REAL rho = fs[0] + (((fs[5] + fs[7]) + (fs[6] + fs[8])) + ((fs[1] + fs[3]) + (fs[2] + fs[4])));
            grho[MIDX(y, x)] = rho;
            REAL invrho = R(1.0) / rho;
            REAL ux = invrho * (((fs[5] - fs[7]) + (fs[8] - fs[6])) + (fs[1] - fs[3]));
            REAL uy = invrho * (((fs[5] - fs[7]) + (fs[6] - fs[8])) + (fs[2] - fs[4]));
            gux[MIDX(y, x)] = ux;
            guy[MIDX(y, x)] = uy;
        }
}

/* ------------------------------------------------------------------ */
/* src/periodic_lbm.f90:45-127  lbm_stream_kernel (pull, periodic wrap) */
void SFX(orc_lbm_stream)(int nx, int ny, int ld, const REAL *fsrc, REAL *fdst)
{
#pragma omp parallel
    {
#pragma omp for schedule(static)
        for (int x = 0; x < nx; ++x)
            for (int y = 0; y < ny; ++y) fdst[FIDX(y, x, 0)] = fsrc[FIDX(y, x, 0)];

#pragma omp for schedule(static)
        for (int x = 0; x < nx; ++x) {
            int xp1 = WRAP_P1(x, nx);          /* mod(x,nx)+1         (1-based) */
            int xm1 = WRAP_M1(x, nx);     /* mod(nx+x-2,nx)+1    (1-based) */
            for (int y = 0; y < ny; ++y) {
                fdst[FIDX(y, x, 1)] = fsrc[FIDX(y, xm1, 1)];
                fdst[FIDX(y, x, 3)] = fsrc[FIDX(y, xp1, 3)];
            }
        }
#pragma omp for schedule(static)
        for (int x = 0; x < nx; ++x)
            for (int y = 0; y < ny; ++y) {
                int yp1 = WRAP_P1(y, ny);
                int ym1 = WRAP_M1(y, ny);
                fdst[FIDX(y, x, 2)] = fsrc[FIDX(ym1, x, 2)];
                fdst[FIDX(y, x, 4)] = fsrc[FIDX(yp1, x, 4)];
            }
#pragma omp for schedule(static)
        for (int x = 0; x < nx; ++x) {
            int xp1 = WRAP_P1(x, nx);
            int xm1 = WRAP_M1(x, nx);
            for (int y = 0; y < ny; ++y) {
                int yp1 = WRAP_P1(y, ny);
                int ym1 = WRAP_M1(y, ny);
                fdst[FIDX(y, x, 5)] = fsrc[FIDX(ym1, xm1, 5)];
                fdst[FIDX(y, x, 7)] = fsrc[FIDX(yp1, xp1, 7)];
            }
        }
#pragma omp for schedule(static)
        for (int x = 0; x < nx; ++x) {
            int xp1 = WRAP_P1(x, nx);
            int xm1 = WRAP_M1(x, nx);
            for (int y = 0; y < ny; ++y) {
                int yp1 = WRAP_P1(y, ny);
                int ym1 = WRAP_M1(y, ny);
                fdst[FIDX(y, x, 6)] = fsrc[FIDX(ym1, xp1, 6)];
                fdst[FIDX(y, x, 8)] = fsrc[FIDX(yp1, xm1, 8)];
            }
        }
    }
}

/* ------------------------------------------------------------------ */
/* src/collision_bgk.F90:35-82  bgk_kernel (default build) */
void SFX(orc_collide_bgk)(int nx, int ny, int ld, REAL *f1, REAL omega)
{
    REAL omegabar = R(1.0) - omega;
#pragma omp parallel for collapse(2) schedule(static)
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y) {
            REAL fs[9], feq[9];
            for (int q = 0; q < 9; ++q) fs[q] = f1[FIDX(y, x, q)];

            REAL rho = (((fs[5] + fs[7]) + (fs[6] + fs[8])) + ((fs[1] + fs[3]) + (fs[2] + fs[4]))) + fs[0];
            REAL invrho = R(1.0) / rho;
            REAL ux = invrho * (((fs[5] - fs[7]) + (fs[8] - fs[6])) + (fs[1] - fs[3]));
            REAL uy = invrho * (((fs[5] - fs[7]) + (fs[6] - fs[8])) + (fs[2] - fs[4]));

            SFX(orc_equilibrium)(rho, ux, uy, feq);

            for (int q = 0; q < 9; ++q) f1[FIDX(y, x, q)] = omegabar * fs[q] + omega * feq[q];
        }
}

/* src/collision_bgk.F90:84-176 bgk_kernel_cache (-DSPLIT)  ==
 * src/periodic_dugks.F90:80-169 kernel_bgk  (textually the same arithmetic:
 * re-associated BGK with omega_w = 3*omega*w and indp = 1/3 - 1/2 |u|^2).
 * Per-node form; the reference's column blocking does not change any
 * per-node operation order. */
static inline void SFX(orc_bgk_split_node)(REAL *fs, REAL omega)
{
    REAL omegabar = R(1.0) - omega;
    REAL omega_w0 = R(3.0) * omega * W0;
    REAL omega_ws = R(3.0) * omega * WS;
    REAL omega_wd = R(3.0) * omega * WD;

    REAL rho = (((fs[5] + fs[7]) + (fs[6] + fs[8])) + ((fs[1] + fs[3]) + (fs[2] + fs[4]))) + fs[0];
    REAL invrho = R(1.0) / rho;
    REAL ux = invrho * (((fs[5] - fs[7]) + (fs[8] - fs[6])) + (fs[1] - fs[3]));
    REAL uy = invrho * (((fs[5] - fs[7]) + (fs[6] - fs[8])) + (fs[2] - fs[4]));
    REAL indp = ONE_THIRD - R(0.5) * (ux * ux + uy * uy);

    fs[0] = omegabar * fs[0] + omega_w0 * rho * indp;

    REAL vel_trm_13 = indp + R(1.5) * ux * ux;
    fs[1] = omegabar * fs[1] + omega_ws * rho * (vel_trm_13 + ux);
    fs[3] = omegabar * fs[3] + omega_ws * rho * (vel_trm_13 - ux);

    REAL vel_trm_24 = indp + R(1.5) * uy * uy;
    fs[2] = omegabar * fs[2] + omega_ws * rho * (vel_trm_24 + uy);
    fs[4] = omegabar * fs[4] + omega_ws * rho * (vel_trm_24 - uy);

    REAL velxpy = ux + uy;
    REAL vel_trm_57 = indp + R(1.5) * velxpy * velxpy;
    fs[5] = omegabar * fs[5] + omega_wd * rho * (vel_trm_57 + velxpy);
    fs[7] = omegabar * fs[7] + omega_wd * rho * (vel_trm_57 - velxpy);

    REAL velxmy = ux - uy;
    REAL vel_trm_68 = indp + R(1.5) * velxmy * velxmy;
    fs[6] = omegabar * fs[6] + omega_wd * rho * (vel_trm_68 - velxmy);
    fs[8] = omegabar * fs[8] + omega_wd * rho * (vel_trm_68 + velxmy);
}

void SFX(orc_kernel_bgk)(int nx, int ny, int ld, REAL *f, REAL omega)
{
#pragma omp parallel for schedule(static)
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y) {
            REAL fs[9];
            for (int q = 0; q < 9; ++q) fs[q] = f[FIDX(y, x, q)];
            SFX(orc_bgk_split_node)(fs, omega);
            for (int q = 0; q < 9; ++q) f[FIDX(y, x, q)] = fs[q];
        }
}

/* ------------------------------------------------------------------ */
/* src/collision_trt.F90:24-34 */
REAL SFX(orc_magic_number)(REAL le, REAL ld_)
{
    return (R(2.0) - le) * (R(2.0) - ld_) / (R(4.0) * le * ld_);
}
REAL SFX(orc_lambda_d)(REAL omega, REAL x)
{
    return (R(4.0) - R(2.0) * omega) / (R(4.0) * x * omega + R(2.0) - omega);
}

/* src/collision_trt.F90:64-160 trt_naive; constants :13-20 */
void SFX(orc_collide_trt)(int nx, int ny, int ld, REAL *f1, REAL lambda_e, REAL lambda_d)
{
    const REAL t0 = R(4.0) / R(9.0);
    const REAL t1x2 = (R(1.0) / R(9.0)) * R(2.0);
    const REAL t2x2 = (R(1.0) / R(36.0)) * R(2.0);
    const REAL inv2csq2 = R(1.0) / (R(2.0) * (R(1.0) / R(3.0)) * (R(1.0) / R(3.0)));
    const REAL fac1 = t1x2 * inv2csq2;
    const REAL fac2 = t2x2 * inv2csq2;

    REAL lambda_e_scaled = R(0.5) * lambda_e;
    REAL lambda_d_scaled = R(0.5) * lambda_d;

#pragma omp parallel for schedule(static)
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y) {
            REAL vC = f1[FIDX(y, x, 0)];
            REAL vE = f1[FIDX(y, x, 1)];
            REAL vN = f1[FIDX(y, x, 2)];
            REAL vW = f1[FIDX(y, x, 3)];
            REAL vS = f1[FIDX(y, x, 4)];
            REAL vNE = f1[FIDX(y, x, 5)];
            REAL vNW = f1[FIDX(y, x, 6)];
            REAL vSW = f1[FIDX(y, x, 7)];
            REAL vSE = f1[FIDX(y, x, 8)];

            REAL rho = (((vNE + vSW) + (vNW + vSE)) + ((vE + vW) + (vN + vS))) + vC;
            REAL velX = (((vNE - vSW) + (vSE - vNW)) + (vE - vW));
            REAL velY = (((vNE - vSW) + (vNW - vSE)) + (vN - vS));
            REAL velX2 = velX * velX;
            REAL velY2 = velY * velY;

            REAL feq_common = rho - R(1.5) * (velX2 + velY2);

            f1[FIDX(y, x, 0)] = vC * (R(1.0) - lambda_e) + lambda_e * t0 * feq_common;

            REAL velXPY = velX + velY;
            REAL sym_NE_SW = lambda_e_scaled * (vNE + vSW - fac2 * velXPY * velXPY - t2x2 * feq_common);
            REAL asym_NE_SW = lambda_d_scaled * (vNE - vSW - R(3.0) * t2x2 * velXPY);
            f1[FIDX(y, x, 5)] = vNE - sym_NE_SW - asym_NE_SW;
            f1[FIDX(y, x, 7)] = vSW - sym_NE_SW + asym_NE_SW;

            REAL velXMY = velX - velY;
            REAL sym_SE_NW = lambda_e_scaled * (vSE + vNW - fac2 * velXMY * velXMY - t2x2 * feq_common);
            REAL asym_SE_NW = lambda_d_scaled * (vSE - vNW - R(3.0) * t2x2 * velXMY);
            f1[FIDX(y, x, 8)] = vSE - sym_SE_NW - asym_SE_NW;
            f1[FIDX(y, x, 6)] = vNW - sym_SE_NW + asym_SE_NW;

            REAL sym_N_S = lambda_e_scaled * (vN + vS - fac1 * velY2 - t1x2 * feq_common);
            REAL asym_N_S = lambda_d_scaled * (vN - vS - R(3.0) * t1x2 * velY);
            f1[FIDX(y, x, 2)] = vN - sym_N_S - asym_N_S;
            f1[FIDX(y, x, 4)] = vS - sym_N_S + asym_N_S;

            REAL sym_E_W = lambda_e_scaled * (vE + vW - fac1 * velX2 - t1x2 * feq_common);
            REAL asym_E_W = lambda_d_scaled * (vE - vW - R(3.0) * t1x2 * velX);
            f1[FIDX(y, x, 1)] = vE - sym_E_W - asym_E_W;
            f1[FIDX(y, x, 3)] = vW - sym_E_W + asym_E_W;
        }
}

/* src/collision_trt.F90:162-290 trt_split (-DSPLIT).  Per node: trt_naive except that the axis
 * pairs evaluate `fac1 * vel * vel` left to right instead of fac1 * (vel*vel). */
void SFX(orc_collide_trt_split)(int nx, int ny, int ld, REAL *f1, REAL lambda_e, REAL lambda_d)
{
    const REAL t0 = R(4.0) / R(9.0);
    const REAL t1x2 = (R(1.0) / R(9.0)) * R(2.0);
    const REAL t2x2 = (R(1.0) / R(36.0)) * R(2.0);
    const REAL inv2csq2 = R(1.0) / (R(2.0) * (R(1.0) / R(3.0)) * (R(1.0) / R(3.0)));
    const REAL fac1 = t1x2 * inv2csq2;
    const REAL fac2 = t2x2 * inv2csq2;
    REAL lambda_e_scaled = R(0.5) * lambda_e;
    REAL lambda_d_scaled = R(0.5) * lambda_d;
#pragma omp parallel for schedule(static)
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y) {
            REAL vC = f1[FIDX(y, x, 0)], vE = f1[FIDX(y, x, 1)], vN = f1[FIDX(y, x, 2)];
            REAL vW = f1[FIDX(y, x, 3)], vS = f1[FIDX(y, x, 4)], vNE = f1[FIDX(y, x, 5)];
            REAL vNW = f1[FIDX(y, x, 6)], vSW = f1[FIDX(y, x, 7)], vSE = f1[FIDX(y, x, 8)];
            REAL rho = (((vNE + vSW) + (vNW + vSE)) + ((vE + vW) + (vN + vS))) + vC;
            REAL velX = (((vNE - vSW) + (vSE - vNW)) + (vE - vW));
            REAL velY = (((vNE - vSW) + (vNW - vSE)) + (vN - vS));
            REAL feq_common = rho - R(1.5) * (velX * velX + velY * velY);
            f1[FIDX(y, x, 0)] = vC * (R(1.0) - lambda_e) + lambda_e * t0 * feq_common;

            REAL velXPY = velX + velY;
            REAL sym_NE_SW = lambda_e_scaled * (vNE + vSW - fac2 * velXPY * velXPY - t2x2 * feq_common);
            REAL asym_NE_SW = lambda_d_scaled * (vNE - vSW - R(3.0) * t2x2 * velXPY);
            f1[FIDX(y, x, 5)] = vNE - sym_NE_SW - asym_NE_SW;
            f1[FIDX(y, x, 7)] = vSW - sym_NE_SW + asym_NE_SW;

            REAL velXMY = velX - velY;
            REAL sym_SE_NW = lambda_e_scaled * (vSE + vNW - fac2 * velXMY * velXMY - t2x2 * feq_common);
            REAL asym_SE_NW = lambda_d_scaled * (vSE - vNW - R(3.0) * t2x2 * velXMY);
            f1[FIDX(y, x, 8)] = vSE - sym_SE_NW - asym_SE_NW;
            f1[FIDX(y, x, 6)] = vNW - sym_SE_NW + asym_SE_NW;

            REAL sym_N_S = lambda_e_scaled * (vN + vS - fac1 * velY * velY - t1x2 * feq_common);
            REAL asym_N_S = lambda_d_scaled * (vN - vS - R(3.0) * t1x2 * velY);
            f1[FIDX(y, x, 2)] = vN - sym_N_S - asym_N_S;
            f1[FIDX(y, x, 4)] = vS - sym_N_S + asym_N_S;

            REAL sym_E_W = lambda_e_scaled * (vE + vW - fac1 * velX * velX - t1x2 * feq_common);
            REAL asym_E_W = lambda_d_scaled * (vE - vW - R(3.0) * t1x2 * velX);
            f1[FIDX(y, x, 1)] = vE - sym_E_W - asym_E_W;
            f1[FIDX(y, x, 3)] = vW - sym_E_W + asym_E_W;
        }
}

/* src/collision_bgk_improved.f90:24-107 bgk_improved_kernel.  The reference declares f1(ny,nx,0:8)
 * (it ignores the padded leading dimension, SURVEY F9: only correct when ny % 16 == 0); the
 * restatement indexes with ld, identical in that case. */
void SFX(orc_collide_bgk_improved)(int nx, int ny, int ld, REAL *f1, REAL omega)
{
    const REAL one_third = R(1.0) / R(3.0), two_thirds = R(2.0) / R(3.0);
    REAL fac = R(4.5) - R(2.25) * omega;
    REAL omegabar = R(1.0) - omega;
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y) {
            REAL vC = f1[FIDX(y, x, 0)], vE = f1[FIDX(y, x, 1)], vN = f1[FIDX(y, x, 2)];
            REAL vW = f1[FIDX(y, x, 3)], vS = f1[FIDX(y, x, 4)], vNE = f1[FIDX(y, x, 5)];
            REAL vNW = f1[FIDX(y, x, 6)], vSW = f1[FIDX(y, x, 7)], vSE = f1[FIDX(y, x, 8)];
            REAL rho = (((vNE + vSW) + (vNW + vSE)) + ((vE + vW) + (vN + vS))) + vC;
            REAL invrho = R(1.0) / rho;
            REAL sumX1 = vE + vNE + vSE;
            REAL sumXN = vW + vNW + vSW;
            REAL sumY1 = vN + vNE + vNW;
            REAL sumYN = vS + vSE + vSW;
            REAL m10 = invrho * (sumX1 - sumXN);
            REAL m01 = invrho * (sumY1 - sumYN);
            REAL u2 = m10 * m10;
            REAL v2 = m01 * m01;
            REAL m20 = invrho * (sumX1 + sumXN);
            REAL m02 = invrho * (sumY1 + sumYN);
            REAL Gx = fac * u2 * (m20 - one_third - u2);
            REAL Gy = fac * v2 * (m02 - one_third - v2);
            REAL X0 = -two_thirds + u2 + Gx;
            REAL X1 = -(X0 + R(1.0) + m10) * R(0.5);
            REAL XN = X1 + m10;
            REAL Y0 = -two_thirds + v2 + Gy;
            REAL Y1 = -(Y0 + R(1.0) + m01) * R(0.5);
            REAL YN = Y1 + m01;
            REAL rho_omega = rho * omega;
            X0 = X0 * rho_omega;
            X1 = X1 * rho_omega;
            XN = XN * rho_omega;
            f1[FIDX(y, x, 0)] = omegabar * vC + X0 * Y0;
            f1[FIDX(y, x, 1)] = omegabar * vE + X1 * Y0;
            f1[FIDX(y, x, 2)] = omegabar * vN + X0 * Y1;
            f1[FIDX(y, x, 3)] = omegabar * vW + XN * Y0;
            f1[FIDX(y, x, 4)] = omegabar * vS + X0 * YN;
            f1[FIDX(y, x, 5)] = omegabar * vNE + X1 * Y1;
            f1[FIDX(y, x, 6)] = omegabar * vNW + XN * Y1;
            f1[FIDX(y, x, 7)] = omegabar * vSW + XN * YN;
            f1[FIDX(y, x, 8)] = omegabar * vSE + X1 * YN;
        }
}

/* ------------------------------------------------------------------ */
/* src/collision_regularized.F90:40-202 rr_kernel_naive */
void SFX(orc_collide_rr)(int nx, int ny, int ld, REAL *f1, REAL omega)
{
    const REAL csqr = R(1.0) / R(3.0);
    REAL omega_w0 = W0 * (R(1.0) - omega);
    REAL omega_ws = WS * (R(1.0) - omega);
    REAL omega_wd = WD * (R(1.0) - omega);

#pragma omp parallel for schedule(static)
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y) {
            REAL feq[9];
            REAL vC = f1[FIDX(y, x, 0)];
            REAL vE = f1[FIDX(y, x, 1)];
            REAL vN = f1[FIDX(y, x, 2)];
            REAL vW = f1[FIDX(y, x, 3)];
            REAL vS = f1[FIDX(y, x, 4)];
            REAL vNE = f1[FIDX(y, x, 5)];
            REAL vNW = f1[FIDX(y, x, 6)];
            REAL vSW = f1[FIDX(y, x, 7)];
            REAL vSE = f1[FIDX(y, x, 8)];

            REAL rho = (((vNE + vSW) + (vNW + vSE)) + ((vE + vW) + (vN + vS))) + vC;
            REAL invrho = R(1.0) / rho;
            REAL ux = invrho * (((vNE - vSW) + (vSE - vNW)) + (vE - vW));
            REAL uy = invrho * (((vNE - vSW) + (vNW - vSE)) + (vN - vS));

            REAL uxx = ux * ux;
            REAL uyy = uy * uy;
            REAL uxxy = uxx * uy;
            REAL uyyx = uyy * ux;
            REAL uxxyy = uxx * uyy;

            REAL indp0 = R(1.0) - R(1.5) * (uxx + uyy);
            REAL indps = indp0 - R(4.5) * uxxyy;
            REAL indpd = indp0 + R(9.0) * uxxyy;
            indp0 = indp0 + R(2.25) * uxxyy;

            feq[0] = W0 * rho * indp0;
            feq[1] = WS * rho * (indps + R(3.0) * ux + R(4.5) * (uxx - uyyx));
            feq[3] = WS * rho * (indps - R(3.0) * ux + R(4.5) * (uxx + uyyx));
            feq[2] = WS * rho * (indps + R(3.0) * uy + R(4.5) * (uyy - uxxy));
            feq[4] = WS * rho * (indps - R(3.0) * uy + R(4.5) * (uyy + uxxy));

            vC = vC - feq[0];
            vE = vE - feq[1];
            vN = vN - feq[2];
            vW = vW - feq[3];
            vS = vS - feq[4];

            REAL axx = csqr * (R(2.0) * (vE + vW) - (vN + vS) - vC);
            REAL ayy = csqr * (R(2.0) * (vN + vS) - (vE + vW) - vC);

            REAL u3p = uxxy + uyyx;
            REAL uxpy = ux + uy;
            REAL indp57 = indpd + R(4.5) * uxpy * uxpy;
            feq[5] = WD * rho * (indp57 + R(3.0) * uxpy + R(9.0) * u3p);
            feq[7] = WD * rho * (indp57 - R(3.0) * uxpy - R(9.0) * u3p);

            REAL u3m = uxxy - uyyx;
            REAL uxmy = ux - uy;
            REAL indp68 = indpd + R(4.5) * uxmy * uxmy;
            feq[6] = WD * rho * (indp68 - R(3.0) * uxmy + R(9.0) * u3m);
            feq[8] = WD * rho * (indp68 + R(3.0) * uxmy - R(9.0) * u3m);

            vNE = vNE - feq[5];
            vNW = vNW - feq[6];
            vSW = vSW - feq[7];
            vSE = vSE - feq[8];

            REAL tmp = R(2.0) * csqr * (vNE + vNW + vSW + vSE);
            axx = axx + tmp;
            ayy = ayy + tmp;

            REAL axy = ((vNE + vSW) - (vNW + vSE));

            REAL axxy = R(2.0) * ux * axy + uy * axx;
            REAL ayyx = R(2.0) * uy * axy + ux * ayy;
            REAL axxyy = R(2.0) * (ux * ayyx + uy * axxy) - uxx * ayy - uyy * axx - R(4.0) * ux * uy * axy;

            indp0 = -R(1.5) * (axx + ayy);
            indps = indp0 - R(4.5) * axxyy;
            indpd = R(9.0) * axxyy - R(2.0) * indp0;
            indp0 = indp0 + R(2.25) * axxyy;

            vC = indp0;
            vE = indps + R(4.5) * (axx - ayyx);
            vW = indps + R(4.5) * (axx + ayyx);
            vN = indps + R(4.5) * (ayy - axxy);
            vS = indps + R(4.5) * (ayy + axxy);
            vNE = indpd + R(9.0) * (axxy + ayyx + axy);
            vSW = indpd - R(9.0) * (axxy + ayyx - axy);
            vNW = indpd + R(9.0) * (axxy - ayyx - axy);
            vSE = indpd - R(9.0) * (axxy - ayyx + axy);

            f1[FIDX(y, x, 0)] = feq[0] + omega_w0 * vC;
            f1[FIDX(y, x, 1)] = feq[1] + omega_ws * vE;
            f1[FIDX(y, x, 2)] = feq[2] + omega_ws * vN;
            f1[FIDX(y, x, 3)] = feq[3] + omega_ws * vW;
            f1[FIDX(y, x, 4)] = feq[4] + omega_ws * vS;
            f1[FIDX(y, x, 5)] = feq[5] + omega_wd * vNE;
            f1[FIDX(y, x, 6)] = feq[6] + omega_wd * vNW;
            f1[FIDX(y, x, 7)] = feq[7] + omega_wd * vSW;
            f1[FIDX(y, x, 8)] = feq[8] + omega_wd * vSE;
        }
}

/* ------------------------------------------------------------------ */
/* 2nd-order half-step back-traced face reconstruction shared by
 * src/fvm_bardow.F90:449-473 and src/periodic_dugks.F90:238-263
 * (identical text). c[0..3] = cfw, cfn, cfe, cfs. */
static inline void SFX(orc_faces)(const REAL *fq, int ld, int x, int y, int xp1, int xm1, int yp1, int ym1,
                                  REAL cxq, REAL cyq, REAL *c)
{
    const REAL p2 = R(0.5), p8 = R(0.125);
#define FQ(yy, xx) fq[(size_t)(yy) + (size_t)ld * (size_t)(xx)]
    REAL fc = FQ(y, x);
    REAL fe = FQ(y, xp1);
    REAL fn = FQ(yp1, x);
    REAL fw = FQ(y, xm1);
    REAL fs = FQ(ym1, x);
    REAL fne = FQ(yp1, xp1);
    REAL fnw = FQ(yp1, xm1);
    REAL fsw = FQ(ym1, xm1);
    REAL fse = FQ(ym1, xp1);
#undef FQ
    c[0] = p2 * (fc + fw) - p2 * cxq * (fc - fw) - p8 * cyq * (fnw + fn - fsw - fs);
    c[1] = p2 * (fc + fn) - p2 * cyq * (fn - fc) - p8 * cxq * (fne + fe - fnw - fw);
    c[2] = p2 * (fc + fe) - p2 * cxq * (fe - fc) - p8 * cyq * (fne + fn - fse - fs);
    c[3] = p2 * (fc + fs) - p2 * cyq * (fc - fs) - p8 * cxq * (fse + fe - fsw - fw);
}

/* src/fvm_bardow.F90:410-507 fvm_bardow_kernel.
 * NOTE: the reference loops `do x = 1, ny` / `do y = 1, nx`
 * (swapped bounds, :438,:444) which is only well defined for nx == ny; the
 * restatement uses the intended bounds, identical for square grids. */
void SFX(orc_stream_fvm_bardow)(int nx, int ny, int ld, const REAL *fold, REAL *fnew, REAL dt)
{
#pragma omp parallel
    {
#pragma omp for schedule(static)
        for (int x = 0; x < nx; ++x)
            for (int y = 0; y < ny; ++y) fnew[FIDX(y, x, 0)] = fold[FIDX(y, x, 0)];

        for (int q = 1; q < 9; ++q) {
            REAL cxq = dt * R(SFX(ocx)[q]);
            REAL cyq = dt * R(SFX(ocy)[q]);
            const REAL *fq = fold + FIDX(0, 0, q);
#pragma omp for schedule(static)
            for (int x = 0; x < nx; ++x) {
                int xp1 = WRAP_P1(x, nx);
                int xm1 = WRAP_M1(x, nx);
                for (int y = 0; y < ny; ++y) {
                    int yp1 = WRAP_P1(y, ny);
                    int ym1 = WRAP_M1(y, ny);
                    REAL c[4];
                    SFX(orc_faces)(fq, ld, x, y, xp1, xm1, yp1, ym1, cxq, cyq, c);
                    REAL fc = fq[(size_t)y + (size_t)ld * x];
                    fnew[FIDX(y, x, q)] = fc - cxq * (c[2] - c[0]) - cyq * (c[1] - c[3]);
                }
            }
        }
    }
}

/* src/fvm_bardow.F90:525-683 fdm_bardow_kernel, default build (none of FDM_WLS, FDM_WLS_GAUSS_V1/V2,
 * FDM_ISO defined): second-order Lax-Wendroff with plain central differences. */
void SFX(orc_stream_fdm_bardow)(int nx, int ny, int ld, const REAL *fold, REAL *fnew, REAL dt)
{
    const REAL p2 = R(0.5);
#pragma omp parallel
    {
#pragma omp for schedule(static)
        for (int x = 0; x < nx; ++x)
            for (int y = 0; y < ny; ++y) fnew[FIDX(y, x, 0)] = fold[FIDX(y, x, 0)];
        for (int q = 1; q < 9; ++q) {
            REAL cxq = dt * R(SFX(ocx)[q]);
            REAL cyq = dt * R(SFX(ocy)[q]);
            REAL cxxq = R(0.5) * cxq * cxq;
            REAL cyyq = R(0.5) * cyq * cyq;
            REAL cxyq = cxq * cyq;
            const REAL *fq = fold + FIDX(0, 0, q);
#define FQ(yy, xx) fq[(size_t)(yy) + (size_t)ld * (size_t)(xx)]
#pragma omp for schedule(static)
            for (int x = 0; x < nx; ++x) {
                int xp1 = WRAP_P1(x, nx);
                int xm1 = WRAP_M1(x, nx);
                for (int y = 0; y < ny; ++y) {
                    int yp1 = WRAP_P1(y, ny);
                    int ym1 = WRAP_M1(y, ny);
                    REAL fc = FQ(y, x), fe = FQ(y, xp1), fn = FQ(yp1, x), fw = FQ(y, xm1), fs = FQ(ym1, x);
                    REAL fne = FQ(yp1, xp1), fnw = FQ(yp1, xm1), fsw = FQ(ym1, xm1), fse = FQ(ym1, xp1);
                    REAL dfx = p2 * (fe - fw);
                    REAL dfy = p2 * (fn - fs);
                    REAL dfxx = fe - R(2.0) * fc + fw;
                    REAL dfyy = fn - R(2.0) * fc + fs;
                    REAL dfxy = R(0.25) * (fne - fse - fnw + fsw);
                    fnew[FIDX(y, x, q)] = fc - cxq * dfx - cyq * dfy + (cxxq * dfxx + cxyq * dfxy + cyyq * dfyy);
                }
            }
#undef FQ
        }
    }
}

/* stream_fdm_bardow built with -DFDM_WLS (stencil 1), -DFDM_WLS_GAUSS_V1 (2), -DFDM_WLS_GAUSS_V2 (3) or
 * -DFDM_ISO (4): src/fvm_bardow.F90:591-660.  stencil 0 is the default build (orc_stream_fdm_bardow). */
void SFX(orc_stream_fdm_bardow_stencil)(int nx, int ny, int ld, const REAL *fold, REAL *fnew, REAL dt, int stencil)
{
    const REAL p2 = R(0.5);
    const REAL two_thirds = R(2.0) / R(3.0), one_sixth = R(1.0) / R(6.0);
    const REAL five_sixths = R(10.0) / R(12.0), one_twelth = R(1.0) / R(12.0);
    const REAL one_third = R(1.0) / R(3.0);
    REAL p1s = R(0.0), p1d = R(0.0), p2c = R(0.0), p2d1 = R(0.0), p2d2 = R(0.0), p2d = R(0.0);
    if (stencil == 2) {
        p1s = R(0.2880584423829145035434), p1d = R(0.1059707788085427065949);
        p2c = R(-1.152233769531658458263), p2d1 = R(0.5761168847658292291314);
        p2d2 = R(-0.4238831152341712149578), p2d = R(0.2119415576170855242122);
    } else if (stencil == 3) {
        p1s = R(0.3934930210807994210853), p1d = R(0.05325348945960039354075);
        p2c = R(-1.573972084323197018207), p2d1 = R(0.7869860421615988421706);
        p2d2 = R(-0.2130139578384016019186), p2d = R(0.1065069789192007732037);
    }
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y) fnew[FIDX(y, x, 0)] = fold[FIDX(y, x, 0)];
    for (int q = 1; q < 9; ++q) {
        REAL cxq = dt * R(SFX(ocx)[q]);
        REAL cyq = dt * R(SFX(ocy)[q]);
        REAL cxxq = R(0.5) * cxq * cxq;
        REAL cyyq = R(0.5) * cyq * cyq;
        REAL cxyq = cxq * cyq;
        const REAL *fq = fold + FIDX(0, 0, q);
#define FQ(yy, xx) fq[(size_t)(yy) + (size_t)ld * (size_t)(xx)]
        for (int x = 0; x < nx; ++x) {
            int xp1 = WRAP_P1(x, nx);
            int xm1 = WRAP_M1(x, nx);
            for (int y = 0; y < ny; ++y) {
                int yp1 = WRAP_P1(y, ny);
                int ym1 = WRAP_M1(y, ny);
                REAL fc = FQ(y, x), fe = FQ(y, xp1), fn = FQ(yp1, x), fw = FQ(y, xm1), fs = FQ(ym1, x);
                REAL fne = FQ(yp1, xp1), fnw = FQ(yp1, xm1), fsw = FQ(ym1, xm1), fse = FQ(ym1, xp1);
                REAL dfx, dfy, dfxx, dfyy, dfxy;
                if (stencil == 1) {
                    dfx = one_sixth * ((fne - fnw) + (fe - fw) + (fse - fsw));
                    dfy = one_sixth * ((fne - fse) + (fn - fs) + (fnw - fsw));
                    dfxx = one_third * (fne - R(2.0) * fn + fnw) + one_third * (fe - R(2.0) * fc + fw) + one_third * (fse - R(2.0) * fs + fsw);
                    dfyy = one_third * (fne - R(2.0) * fe + fse) + one_third * (fn - R(2.0) * fc + fs) + one_third * (fnw - R(2.0) * fw + fsw);
                    dfxy = R(0.25) * (fne - fnw + fsw - fse);
                } else if (stencil == 2 || stencil == 3) {
                    dfx = p1s * (fe - fw) + p1d * (fne - fnw) + p1d * (fse - fsw);
                    dfy = p1s * (fn - fs) + p1d * (fne - fse) + p1d * (fnw - fsw);
                    dfxx = p2c * fc + p2d1 * (fe + fw) + p2d2 * (fn + fs) + p2d * (fne + fnw + fsw + fse);
                    dfyy = p2c * fc + p2d2 * (fe + fw) + p2d1 * (fn + fs) + p2d * (fne + fnw + fsw + fse);
                    dfxy = R(0.25) * (fne - fnw + fsw - fse);
                } else if (stencil == 4) {
                    dfx = p2 * (one_sixth * (fne - fnw) + two_thirds * (fe - fw) + one_sixth * (fse - fsw));
                    dfy = p2 * (one_sixth * (fne - fse) + two_thirds * (fn - fs) + one_sixth * (fnw - fsw));
                    dfxx = one_twelth * (fne - R(2.0) * fn + fnw) + five_sixths * (fe - R(2.0) * fc + fw) + one_twelth * (fse - R(2.0) * fs + fsw);
                    dfyy = one_twelth * (fne - R(2.0) * fe + fse) + five_sixths * (fn - R(2.0) * fc + fs) + one_twelth * (fnw - R(2.0) * fw + fsw);
                    dfxy = R(0.25) * (fne - fse - fnw + fsw);
                } else {
                    dfx = p2 * (fe - fw);
                    dfy = p2 * (fn - fs);
                    dfxx = fe - R(2.0) * fc + fw;
                    dfyy = fn - R(2.0) * fc + fs;
                    dfxy = R(0.25) * (fne - fse - fnw + fsw);
                }
                fnew[FIDX(y, x, q)] = fc - cxq * dfx - cyq * dfy + (cxxq * dfxx + cxyq * dfxy + cyyq * dfyy);
            }
        }
#undef FQ
    }
}

/* src/fvm_bardow.F90:702-891 fdm_sofonea_kernel: per population, 1-D Lax-Wendroff along its own
 * characteristic; fu = f(x + c_q), fd = f(x - c_q).  (Loop bounds swapped in the reference, F9.) */
void SFX(orc_stream_fdm_sofonea)(int nx, int ny, int ld, const REAL *fold, REAL *fnew, REAL dt)
{
    const REAL p2 = R(0.5) / MSQRT(R(2.0));
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y) fnew[FIDX(y, x, 0)] = fold[FIDX(y, x, 0)];
    for (int q = 1; q < 9; ++q) {
        const int cx = SFX(ocx)[q], cy = SFX(ocy)[q];
        for (int x = 0; x < nx; ++x) {
            int xu = cx == 1 ? WRAP_P1(x, nx) : (cx == -1 ? WRAP_M1(x, nx) : x);
            int xd = cx == 1 ? WRAP_M1(x, nx) : (cx == -1 ? WRAP_P1(x, nx) : x);
            for (int y = 0; y < ny; ++y) {
                int yu = cy == 1 ? WRAP_P1(y, ny) : (cy == -1 ? WRAP_M1(y, ny) : y);
                int yd = cy == 1 ? WRAP_M1(y, ny) : (cy == -1 ? WRAP_P1(y, ny) : y);
                REAL fc = fold[FIDX(y, x, q)], fu = fold[FIDX(yu, xu, q)], fd = fold[FIDX(yd, xd, q)];
                REAL du1, du2;
                if (q <= 4) {
                    du1 = R(0.5) * (fu - fd);
                    du2 = fu - R(2.0) * fc + fd;
                } else {
                    du1 = p2 * (fu - fd);
                    du2 = R(0.5) * (fu - R(2.0) * fc + fd);
                }
                fnew[FIDX(y, x, q)] = fc + dt * (R(0.5) * dt * du2 - du1);
            }
        }
    }
}

/* ------------------------------------------------------------------ */
/* src/periodic_dugks.F90:310-434 update_ew / update_ns (-DDUGKS):
 * face moments from all nine face values, relaxation of the flux-carrying
 * populations only (ew: 1,3,5,7,6,8 ; ns: 2,4,5,7,6,8). */
static inline void SFX(orc_face_relax)(REAL *f, REAL omega, int ew)
{
    REAL omegabar = R(1.0) - omega;
    REAL omega_ws = R(3.0) * omega * WS;
    REAL omega_wd = R(3.0) * omega * WD;

    REAL rho = (((f[5] + f[7]) + (f[6] + f[8])) + ((f[1] + f[3]) + (f[2] + f[4]))) + f[0];
    REAL invrho = R(1.0) / rho;
    REAL ux = invrho * (((f[5] - f[7]) + (f[8] - f[6])) + (f[1] - f[3]));
    REAL uy = invrho * (((f[5] - f[7]) + (f[6] - f[8])) + (f[2] - f[4]));
    REAL indp = ONE_THIRD - R(0.5) * (ux * ux + uy * uy);

    if (ew) {
        REAL vel_trm_13 = indp + R(1.5) * ux * ux;
        f[1] = omegabar * f[1] + omega_ws * rho * (vel_trm_13 + ux);
        f[3] = omegabar * f[3] + omega_ws * rho * (vel_trm_13 - ux);
    } else {
        REAL vel_trm_24 = indp + R(1.5) * uy * uy;
        f[2] = omegabar * f[2] + omega_ws * rho * (vel_trm_24 + uy);
        f[4] = omegabar * f[4] + omega_ws * rho * (vel_trm_24 - uy);
    }
    REAL velxpy = ux + uy;
    REAL vel_trm_57 = indp + R(1.5) * velxpy * velxpy;
    f[5] = omegabar * f[5] + omega_wd * rho * (vel_trm_57 + velxpy);
    f[7] = omegabar * f[7] + omega_wd * rho * (vel_trm_57 - velxpy);

    REAL velxmy = ux - uy;
    REAL vel_trm_68 = indp + R(1.5) * velxmy * velxmy;
    f[6] = omegabar * f[6] + omega_wd * rho * (vel_trm_68 - velxmy);
    f[8] = omegabar * f[8] + omega_wd * rho * (vel_trm_68 + velxmy);
}

/* src/periodic_dugks.F90:190-304 kernel_stream.  ft = fbar^{+}, fp = ftilde^{+}
 * (updated in place for q = 1..8).  dugks != 0 selects the -DDUGKS branch. */
void SFX(orc_dugks_kernel_stream)(int nx, int ny, int ld, const REAL *ft, REAL *fp, REAL dt, REAL omega,
                                  int dugks)
{
#pragma omp parallel for schedule(static)
    for (int x = 0; x < nx; ++x) {
        int xp1 = WRAP_P1(x, nx);
        int xm1 = WRAP_M1(x, nx);
        for (int y = 0; y < ny; ++y) {
            int yp1 = WRAP_P1(y, ny);
            int ym1 = WRAP_M1(y, ny);
            REAL cfw[9], cfn[9], cfe[9], cfs[9];
            for (int q = 0; q < 9; ++q) {
                REAL cxq = dt * R(SFX(ocx)[q]);
                REAL cyq = dt * R(SFX(ocy)[q]);
                REAL c[4];
                SFX(orc_faces)(ft + FIDX(0, 0, q), ld, x, y, xp1, xm1, yp1, ym1, cxq, cyq, c);
                cfw[q] = c[0];
                cfn[q] = c[1];
                cfe[q] = c[2];
                cfs[q] = c[3];
            }
            if (dugks) {
                SFX(orc_face_relax)(cfw, omega, 1);
                SFX(orc_face_relax)(cfe, omega, 1);
                SFX(orc_face_relax)(cfn, omega, 0);
                SFX(orc_face_relax)(cfs, omega, 0);
            }
            for (int q = 1; q < 9; ++q) {
                REAL cxq = dt * R(SFX(ocx)[q]);
                REAL cyq = dt * R(SFX(ocy)[q]);
                fp[FIDX(y, x, q)] = fp[FIDX(y, x, q)] - cxq * (cfe[q] - cfw[q]) - cyq * (cfn[q] - cfs[q]);
            }
        }
    }
}

/* src/periodic_dugks.F90:46-77 dugks_collide (+ copy_field :441-486).
 * omega_grid = grid%omega, tau_d = grid%tau/grid%dt (:40-44). */
void SFX(orc_dugks_collide)(int nx, int ny, int ld, REAL *fold, REAL *fnew, REAL omega_grid, REAL tau,
                            REAL dt, int dugks)
{
    REAL tau_d = tau / dt;
    size_t n = (size_t)ld * nx * 9;
    /* copy_field copies whole padded columns */
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) fnew[i] = fold[i];

    REAL omega = R(1.0) / (tau_d + R(0.5));
    SFX(orc_kernel_bgk)(nx, ny, ld, fnew, omega_grid);
    if (dugks) omega = R(0.75) * omega;
    SFX(orc_kernel_bgk)(nx, ny, ld, fold, omega);
}

/* src/periodic_dugks.F90:172-188 dugks_stream */
void SFX(orc_dugks_stream)(int nx, int ny, int ld, const REAL *fold, REAL *fnew, REAL tau, REAL dt, int dugks)
{
    REAL tau_d = tau / dt;
    REAL omega = R(1.0) / (R(4.0) * tau_d + R(1.0));
    SFX(orc_dugks_kernel_stream)(nx, ny, ld, fold, fnew, dt, omega, dugks);
}

/* ------------------------------------------------------------------ */
/* src/vorticity.f90:13-43 */
void SFX(orc_vorticity_2nd)(int nx, int ny, const REAL *ux, const REAL *uy, REAL *omega)
{
#pragma omp parallel for schedule(static)
    for (int x = 0; x < nx; ++x) {
        int xp1 = WRAP_P1(x, nx);
        int xm1 = WRAP_M1(x, nx);
        for (int y = 0; y < ny; ++y) {
            int yp1 = WRAP_P1(y, ny);
            int ym1 = WRAP_M1(y, ny);
            REAL duydx = R(0.5) * (uy[MIDX(y, xp1)] - uy[MIDX(y, xm1)]);
            REAL duxdy = R(0.5) * (ux[MIDX(yp1, x)] - ux[MIDX(ym1, x)]);
            omega[MIDX(y, x)] = duydx - duxdy;
        }
    }
}

/* src/vorticity.f90:46-87.  The weights are as in the reference (1/12 on the
 * +-1 pair, 2/3 on the reversed +-2 pair: SURVEY F9) -- reproduced, not fixed. */
void SFX(orc_vorticity_4th)(int nx, int ny, const REAL *ux, const REAL *uy, REAL *omega)
{
    const REAL t1 = R(1.0) / R(12.0), t2 = R(2.0) / R(3.0);
#pragma omp parallel for schedule(static)
    for (int x = 0; x < nx; ++x) {
        int xp1 = WRAP_P1(x, nx);
        int xm1 = WRAP_M1(x, nx);
        int xp2 = ((x + 2) % nx);
        int xm2 = ((nx + x - 2) % nx);
        for (int y = 0; y < ny; ++y) {
            int yp1 = WRAP_P1(y, ny);
            int ym1 = WRAP_M1(y, ny);
            int yp2 = (y + 2) % ny;
            int ym2 = (ny + y - 2) % ny;
            REAL duydx = t1 * (uy[MIDX(y, xp1)] - uy[MIDX(y, xm1)]) + t2 * (uy[MIDX(y, xm2)] - uy[MIDX(y, xp2)]);
            REAL duxdy = t1 * (ux[MIDX(yp1, x)] - ux[MIDX(ym1, x)]) + t2 * (ux[MIDX(ym2, x)] - ux[MIDX(yp2, x)]);
            omega[MIDX(y, x)] = duydx - duxdy;
        }
    }
}

/* ------------------------------------------------------------------ */
/* src/benchmarks/taylor_green.f90:31-84.  td = 1/(nu (kx^2+ky^2)) (:42).
 * (The reference inner loop runs to self%nx, :68 -- square grids only.) */
REAL SFX(orc_tg_decay_time)(REAL kx, REAL ky, REAL nu)
{
    return R(1.0) / (nu * (kx * kx + ky * ky));
}

void SFX(orc_taylor_green_eval)(int nx, int ny, REAL kx, REAL ky, REAL umax, REAL td, REAL t, REAL *p,
                                REAL *ux, REAL *uy)
{
    for (int x = 0; x < nx; ++x) {
        REAL xx = R(x) + R(0.5);
        for (int y = 0; y < ny; ++y) {
            REAL yy = R(y) + R(0.5);
            ux[MIDX(y, x)] = -umax * MSQRT(ky / kx) * MCOS(kx * xx) * MSIN(ky * yy) * MEXP(-t / td);
            uy[MIDX(y, x)] = umax * MSQRT(kx / ky) * MSIN(kx * xx) * MCOS(ky * yy) * MEXP(-t / td);
            p[MIDX(y, x)] = -R(0.25) * (umax * umax) *
                            ((ky / kx) * MCOS(R(2.0) * kx * xx) + (kx / ky) * MCOS(R(2.0) * ky * yy)) *
                            MEXP(-R(2.0) * t / td);
        }
    }
}

/* src/benchmarks/barotropic_vortex_case.F90:34-83 eval_vortex_case */
void SFX(orc_vortex_eval)(int nx, int ny, REAL U0, REAL xc, REAL yc, REAL Rc, REAL eps, REAL rho0,
                          REAL csqr, REAL *rho, REAL *ux, REAL *uy)
{
    const REAL half = R(1.0) / R(2.0);
    REAL Rcsqr = Rc * Rc;
    REAL vMa_sq = eps * eps / csqr;
    for (int x = 0; x < nx; ++x) {
        REAL xl = R(x) + R(0.5);
        for (int y = 0; y < ny; ++y) {
            REAL yl = R(y) + R(0.5);
            REAL xr = xl - xc;
            REAL yr = yl - yc;
            REAL rsqr = xr * xr + yr * yr;
            ux[MIDX(y, x)] = U0 - eps * (yr / Rc) * MEXP(-half * rsqr / Rcsqr);
            uy[MIDX(y, x)] = eps * (xr / Rc) * MEXP(-half * rsqr / Rcsqr);
            rho[MIDX(y, x)] = rho0 * MEXP(-half * vMa_sq * MEXP(-rsqr / Rcsqr));
        }
    }
}

/* libgfortran NORM2 (generic/norm2 template, not vendored in the reference):
 * one-pass scaled sum of squares; restated from its published algorithm. */
static REAL SFX(orc_norm2)(const REAL *v, size_t n)
{
    REAL scale = R(1.0), result = R(0.0);
    for (size_t i = 0; i < n; ++i) {
        if (v[i] != R(0.0)) {
            REAL absx = MFABS(v[i]);
            if (scale < absx) {
                REAL val = scale / absx;
                result = R(1.0) + result * val * val;
                scale = absx;
            } else {
                REAL val = absx / scale;
                result += val * val;
            }
        }
    }
    return scale * MSQRT(result);
}

/* app/main_taylor_green.f90:174-212 calc_L2_norm:
 * norm2(hypot(ux-uxa, uy-uya)) / norm2(hypot(uxa,uya)); scratch: 2*nx*ny */
REAL SFX(orc_l2_norm)(int nx, int ny, const REAL *ux, const REAL *uy, const REAL *uxa, const REAL *uya,
                      REAL *scratch)
{
    size_t n = (size_t)nx * ny;
    REAL *a = scratch, *b = scratch + n;
    for (size_t i = 0; i < n; ++i) {
        a[i] = MHYPOT(ux[i] - uxa[i], uy[i] - uya[i]);
        b[i] = MHYPOT(uxa[i], uya[i]);
    }
    return SFX(orc_norm2)(a, n) / SFX(orc_norm2)(b, n);
}

/* ------------------------------------------------------------------ */
/* Step orchestrators on a two-lattice state f[2] (indices 1-based like the
 * reference: inew=1, iold=2 after alloc_grid, src/fvm_bardow.F90:171-172).
 *   scheme 0: perform_lbm_step   src/periodic_lbm.f90:15-29  (lbm_stream)
 *   scheme 1: perform_step       src/fvm_bardow.F90:307-320  (stream_fvm_bardow)
 *   scheme 2: perform_dugks_step src/periodic_dugks.F90:25-38 (-DDUGKS)
 *   scheme 3: perform_dugks_step without -DDUGKS
 *   scheme 4 / 5: perform_step with stream_fdm_bardow / stream_fdm_sofonea (src/fvm_bardow.F90:511, 688)
 * collision: 0 bgk, 1 trt, 2 rr, 3 bgk -DSPLIT (schemes 0/1 only).
 * idx[0]=iold, idx[1]=inew (1-based), updated in place. */
void SFX(orc_run)(int nx, int ny, int ld, REAL *f1, REAL *f2, int *idx, int scheme, int collision,
                  REAL omega, REAL tau, REAL dt, REAL trt_magic, long nsteps)
{
    REAL *f[3] = {0, f1, f2};
    for (long s = 0; s < nsteps; ++s) {
        REAL *fo = f[idx[0]], *fn = f[idx[1]];
        if (scheme == 0 || scheme == 1 || scheme == 4 || scheme == 5) {
            if (scheme == 0)
                SFX(orc_lbm_stream)(nx, ny, ld, fo, fn);
            else if (scheme == 1)
                SFX(orc_stream_fvm_bardow)(nx, ny, ld, fo, fn, dt);
            else if (scheme == 4)
                SFX(orc_stream_fdm_bardow)(nx, ny, ld, fo, fn, dt);
            else
                SFX(orc_stream_fdm_sofonea)(nx, ny, ld, fo, fn, dt);
            switch (collision) {
            case 0: SFX(orc_collide_bgk)(nx, ny, ld, fn, omega); break;
            case 1: SFX(orc_collide_trt)(nx, ny, ld, fn, omega, SFX(orc_lambda_d)(omega, trt_magic)); break;
            case 2: SFX(orc_collide_rr)(nx, ny, ld, fn, omega); break;
            case 4: SFX(orc_collide_trt_split)(nx, ny, ld, fn, omega, SFX(orc_lambda_d)(omega, trt_magic)); break;
            case 5: SFX(orc_collide_bgk_improved)(nx, ny, ld, fn, omega); break;
            default: SFX(orc_kernel_bgk)(nx, ny, ld, fn, omega); break;
            }
        } else {
            int dugks = (scheme == 2);
            SFX(orc_dugks_collide)(nx, ny, ld, fo, fn, omega, tau, dt, dugks);
            SFX(orc_dugks_stream)(nx, ny, ld, fo, fn, tau, dt, dugks);
        }
        int t = idx[0];
        idx[0] = idx[1];
        idx[1] = t;
    }
}

/* ------------------------------------------------------------------ */
/* sim/ plugin seam (secondary): DDF-shifted standard LBM, x fastest,
 * arrays (0:nx+1, 0:ny+1, 0:8).  sim/sim.F90:353-401 (equilibrium),
 * :181-199 (lbm_eqinit), :404-505 (collide_and_stream_fused, push),
 * :568-624 (periodic_bc_push), :148-179 (lbm_macros). */
#define SIDX(i, j, k) ((size_t)(i) + (size_t)(nx + 2) * ((size_t)(j) + (size_t)(ny + 2) * (size_t)(k)))

static inline void SFX(orc_sim_equilibrium)(REAL rho, REAL ux, REAL uy, REAL *feq)
{
    const REAL rho0 = R(1.0);
    const REAL w[9] = {W0, WS, WS, WS, WS, WD, WD, WD, WD};
    REAL uxx = ux * ux, uyy = uy * uy;
    REAL uxpy = ux + uy, uxmy = ux - uy;
    REAL indp = -R(1.5) * (uxx + uyy);
    feq[0] = W0 * rho * (indp);
    feq[1] = WS * rho * (indp + R(3.0) * ux + R(4.5) * uxx);
    feq[2] = WS * rho * (indp + R(3.0) * uy + R(4.5) * uyy);
    feq[3] = WS * rho * (indp - R(3.0) * ux + R(4.5) * uxx);
    feq[4] = WS * rho * (indp - R(3.0) * uy + R(4.5) * uyy);
    feq[5] = WD * rho * (indp + R(3.0) * uxpy + R(4.5) * uxpy * uxpy);
    feq[7] = WD * rho * (indp - R(3.0) * uxpy + R(4.5) * uxpy * uxpy);
    feq[6] = WD * rho * (indp - R(3.0) * uxmy + R(4.5) * uxmy * uxmy);
    feq[8] = WD * rho * (indp + R(3.0) * uxmy + R(4.5) * uxmy * uxmy);
    for (int k = 0; k < 9; ++k) feq[k] = feq[k] + w[k] * (rho - rho0);
}

/* p(nx,ny), u(nx,ny), v(nx,ny): x fastest. halo cells are zeroed (the
 * reference leaves them undefined; they are never read before written). */
void SFX(orc_sim_eqinit)(int nx, int ny, REAL *f, const REAL *p, const REAL *u, const REAL *v)
{
    const REAL rho0 = R(1.0);
    for (size_t i = 0; i < (size_t)(nx + 2) * (ny + 2) * 9; ++i) f[i] = R(0.0);
    for (int j = 1; j <= ny; ++j)
        for (int i = 1; i <= nx; ++i) {
            size_t m = (size_t)(i - 1) + (size_t)nx * (j - 1);
            REAL rho = rho0 + p[m] / CSQR;
            REAL feq[9];
            SFX(orc_sim_equilibrium)(rho, u[m], v[m], feq);
            for (int k = 0; k < 9; ++k) f[SIDX(i, j, k)] = feq[k];
        }
}

void SFX(orc_sim_collide_and_stream)(int nx, int ny, const REAL *fsrc, REAL *fdst, REAL omega)
{
    const REAL rho0 = R(1.0);
    const REAL w[9] = {W0, WS, WS, WS, WS, WD, WD, WD, WD};
    static const int cx[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
    static const int cy[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
    for (int j = 1; j <= ny; ++j)
        for (int i = 1; i <= nx; ++i) {
            REAL f[9], feq[9];
            for (int k = 0; k < 9; ++k) f[k] = fsrc[SIDX(i, j, k)];
            REAL rho = (((f[5] + f[7]) + (f[6] + f[8])) + ((f[1] + f[3]) + (f[2] + f[4]))) + f[0];
            rho = rho + rho0;
            REAL ux = (((f[5] - f[7]) + (f[8] - f[6])) + (f[1] - f[3])) / rho;
            REAL uy = (((f[5] - f[7]) + (f[6] - f[8])) + (f[2] - f[4])) / rho;
            REAL uxx = ux * ux, uyy = uy * uy;
            REAL indp = -R(1.5) * (uxx + uyy);
            feq[0] = W0 * rho * (indp);
            feq[1] = WS * rho * (indp + R(3.0) * ux + R(4.5) * uxx);
            feq[2] = WS * rho * (indp + R(3.0) * uy + R(4.5) * uyy);
            feq[3] = WS * rho * (indp - R(3.0) * ux + R(4.5) * uxx);
            feq[4] = WS * rho * (indp - R(3.0) * uy + R(4.5) * uyy);
            REAL uxpy = ux + uy;
            feq[5] = WD * rho * (indp + R(3.0) * uxpy + R(4.5) * uxpy * uxpy);
            feq[7] = WD * rho * (indp - R(3.0) * uxpy + R(4.5) * uxpy * uxpy);
            REAL uxmy = ux - uy;
            feq[6] = WD * rho * (indp - R(3.0) * uxmy + R(4.5) * uxmy * uxmy);
            feq[8] = WD * rho * (indp + R(3.0) * uxmy + R(4.5) * uxmy * uxmy);
            for (int k = 0; k < 9; ++k) feq[k] = feq[k] + w[k] * (rho - rho0);
            for (int k = 0; k < 9; ++k) fdst[SIDX(i + cx[k], j + cy[k], k)] = omega * (feq[k] - f[k]) + f[k];
        }
}

void SFX(orc_sim_periodic_bc_push)(int nx, int ny, REAL *f)
{
    for (int j = 1; j <= ny; ++j) { /* EAST, WEST */
        f[SIDX(nx, j, 6)] = f[SIDX(0, j, 6)];
        f[SIDX(nx, j, 3)] = f[SIDX(0, j, 3)];
        f[SIDX(nx, j, 7)] = f[SIDX(0, j, 7)];
        f[SIDX(1, j, 5)] = f[SIDX(nx + 1, j, 5)];
        f[SIDX(1, j, 1)] = f[SIDX(nx + 1, j, 1)];
        f[SIDX(1, j, 8)] = f[SIDX(nx + 1, j, 8)];
    }
    for (int i = 1; i <= nx; ++i) { /* NORTH, SOUTH */
        f[SIDX(i, ny, 7)] = f[SIDX(i, 0, 7)];
        f[SIDX(i, ny, 4)] = f[SIDX(i, 0, 4)];
        f[SIDX(i, ny, 8)] = f[SIDX(i, 0, 8)];
        f[SIDX(i, 1, 6)] = f[SIDX(i, ny + 1, 6)];
        f[SIDX(i, 1, 2)] = f[SIDX(i, ny + 1, 2)];
        f[SIDX(i, 1, 5)] = f[SIDX(i, ny + 1, 5)];
    }
    /* corners last */
    f[SIDX(nx, ny, 7)] = f[SIDX(0, 0, 7)];
    f[SIDX(nx, 1, 6)] = f[SIDX(0, ny + 1, 6)];
    f[SIDX(1, ny, 8)] = f[SIDX(nx + 1, 0, 8)];
    f[SIDX(1, 1, 5)] = f[SIDX(nx + 1, ny + 1, 5)];
}

void SFX(orc_sim_macros)(int nx, int ny, const REAL *fsrc, REAL *rho, REAL *u, REAL *v)
{
    const REAL rho0 = R(1.0);
    for (int j = 1; j <= ny; ++j)
        for (int i = 1; i <= nx; ++i) {
            REAL f[9];
            for (int k = 0; k < 9; ++k) f[k] = fsrc[SIDX(i, j, k)];
            size_t m = (size_t)(i - 1) + (size_t)nx * (j - 1);
            rho[m] = (((f[5] + f[7]) + (f[6] + f[8])) + ((f[1] + f[3]) + (f[2] + f[4]))) + f[0];
            rho[m] = rho[m] + rho0;
            u[m] = (((f[5] - f[7]) + (f[8] - f[6])) + (f[1] - f[3])) / rho[m];
            v[m] = (((f[5] - f[7]) + (f[6] - f[8])) + (f[2] - f[4])) / rho[m];
        }
}

/* sim/sim_lw.F90:24-85 lw_stream: second-order Lax-Wendroff streaming on the haloed (0:nx+1,0:ny+1,0:8)
 * arrays (interior only; the halo must have been filled by lw_bc). */
void SFX(orc_lw_stream)(int nx, int ny, const REAL *fsrc, REAL *fdst, REAL dt)
{
    static const int cx[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
    static const int cy[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
    for (int j = 1; j <= ny; ++j)
        for (int i = 1; i <= nx; ++i) fdst[SIDX(i, j, 0)] = fsrc[SIDX(i, j, 0)];
    for (int k = 1; k < 9; ++k) {
        REAL vx = dt * R(cx[k]), vy = dt * R(cy[k]);
        REAL vxx = R(0.5) * vx * vx, vyy = R(0.5) * vy * vy, vxy = vx * vy;
        for (int j = 1; j <= ny; ++j)
            for (int i = 1; i <= nx; ++i) {
                REAL dfx = R(0.5) * (fsrc[SIDX(i + 1, j, k)] - fsrc[SIDX(i - 1, j, k)]);
                REAL dfy = R(0.5) * (fsrc[SIDX(i, j + 1, k)] - fsrc[SIDX(i, j - 1, k)]);
                REAL dfxx = fsrc[SIDX(i + 1, j, k)] - R(2.0) * fsrc[SIDX(i, j, k)] + fsrc[SIDX(i - 1, j, k)];
                REAL dfyy = fsrc[SIDX(i, j + 1, k)] - R(2.0) * fsrc[SIDX(i, j, k)] + fsrc[SIDX(i, j - 1, k)];
                REAL dfxy = R(0.25) * (fsrc[SIDX(i + 1, j + 1, k)] - fsrc[SIDX(i - 1, j + 1, k)] + fsrc[SIDX(i - 1, j - 1, k)] -
                                       fsrc[SIDX(i + 1, j - 1, k)]);
                fdst[SIDX(i, j, k)] = fsrc[SIDX(i, j, k)] - vx * dfx - vy * dfy + (vxx * dfxx + vxy * dfxy + vyy * dfyy);
            }
    }
}

/* sim/sim_lw.F90:87-166 lw_collision: DDF-shifted BGK, velocity by reciprocal multiplication */
void SFX(orc_lw_collision)(int nx, int ny, REAL *pdf, REAL omega)
{
    const REAL rho0 = R(1.0);
    const REAL w[9] = {W0, WS, WS, WS, WS, WD, WD, WD, WD};
    for (int j = 1; j <= ny; ++j)
        for (int i = 1; i <= nx; ++i) {
            REAL f[9], feq[9];
            for (int k = 0; k < 9; ++k) f[k] = pdf[SIDX(i, j, k)];
            REAL rho = f[0] + (((f[5] + f[7]) + (f[6] + f[8])) + ((f[1] + f[3]) + (f[2] + f[4]))) + rho0;
            REAL irho = R(1.0) / rho;
            REAL ux = (((f[5] - f[7]) + (f[8] - f[6])) + (f[1] - f[3])) * irho;
            REAL uy = (((f[5] - f[7]) + (f[6] - f[8])) + (f[2] - f[4])) * irho;
            REAL uxx = ux * ux, uyy = uy * uy;
            REAL indp = -R(1.5) * (uxx + uyy);
            feq[0] = W0 * rho * (indp);
            feq[1] = WS * rho * (indp + R(3.0) * ux + R(4.5) * uxx);
            feq[2] = WS * rho * (indp + R(3.0) * uy + R(4.5) * uyy);
            feq[3] = WS * rho * (indp - R(3.0) * ux + R(4.5) * uxx);
            feq[4] = WS * rho * (indp - R(3.0) * uy + R(4.5) * uyy);
            REAL uxpy = ux + uy;
            feq[5] = WD * rho * (indp + R(3.0) * uxpy + R(4.5) * uxpy * uxpy);
            feq[7] = WD * rho * (indp - R(3.0) * uxpy + R(4.5) * uxpy * uxpy);
            REAL uxmy = ux - uy;
            feq[6] = WD * rho * (indp - R(3.0) * uxmy + R(4.5) * uxmy * uxmy);
            feq[8] = WD * rho * (indp + R(3.0) * uxmy + R(4.5) * uxmy * uxmy);
            for (int k = 0; k < 9; ++k) feq[k] = feq[k] + w[k] * (rho - rho0);
            for (int k = 0; k < 9; ++k) pdf[SIDX(i, j, k)] = f[k] + omega * (feq[k] - f[k]);
        }
}

/* sim/sim_lw.F90:169-197 lw_bc: periodic halo of ALL populations, edges then corners */
void SFX(orc_lw_bc)(int nx, int ny, REAL *f)
{
    for (int k = 0; k < 9; ++k) {
        for (int i = 1; i <= nx; ++i) {
            f[SIDX(i, 0, k)] = f[SIDX(i, ny, k)];
            f[SIDX(i, ny + 1, k)] = f[SIDX(i, 1, k)];
        }
        for (int j = 1; j <= ny; ++j) {
            f[SIDX(0, j, k)] = f[SIDX(nx, j, k)];
            f[SIDX(nx + 1, j, k)] = f[SIDX(1, j, k)];
        }
        f[SIDX(0, 0, k)] = f[SIDX(nx, ny, k)];
        f[SIDX(nx + 1, ny + 1, k)] = f[SIDX(1, 1, k)];
        f[SIDX(0, ny + 1, k)] = f[SIDX(nx, 1, k)];
        f[SIDX(nx + 1, 0, k)] = f[SIDX(1, ny, k)];
    }
}

/* ------------------------------------------------------------------ */
/* Higher-order Lax-Wendroff plugins lw4 / lw6 (sim/sim_lw4.F90, sim/sim_lw6.F90): same DDF-shifted
 * collision as lw, wider streaming stencils, arrays (1-H:nx+H, 1-H:ny+H, 0:8) with H = 2 / 3. */
#define HIDX(i, j, k) ((size_t)((i) + H - 1) + (size_t)(nx + 2 * H) * ((size_t)((j) + H - 1) + (size_t)(ny + 2 * H) * (size_t)(k)))

/* lbm_eqinit_fields, sim/sim.F90:181-199, halo width H */
void SFX(orc_simh_eqinit)(int nx, int ny, int H, REAL *f, const REAL *p, const REAL *u, const REAL *v)
{
    const REAL rho0 = R(1.0);
    for (size_t i = 0; i < (size_t)(nx + 2 * H) * (ny + 2 * H) * 9; ++i) f[i] = R(0.0);
    for (int j = 1; j <= ny; ++j)
        for (int i = 1; i <= nx; ++i) {
            size_t m = (size_t)(i - 1) + (size_t)nx * (j - 1);
            REAL rho = rho0 + p[m] / CSQR;
            REAL feq[9];
            SFX(orc_sim_equilibrium)(rho, u[m], v[m], feq);
            for (int k = 0; k < 9; ++k) f[HIDX(i, j, k)] = feq[k];
        }
}

/* lbm_macros, sim/sim.F90:148-179, halo width H */
void SFX(orc_simh_macros)(int nx, int ny, int H, const REAL *fsrc, REAL *rho, REAL *u, REAL *v)
{
    const REAL rho0 = R(1.0);
    for (int j = 1; j <= ny; ++j)
        for (int i = 1; i <= nx; ++i) {
            REAL f[9];
            for (int k = 0; k < 9; ++k) f[k] = fsrc[HIDX(i, j, k)];
            size_t m = (size_t)(i - 1) + (size_t)nx * (j - 1);
            rho[m] = (((f[5] + f[7]) + (f[6] + f[8])) + ((f[1] + f[3]) + (f[2] + f[4]))) + f[0];
            rho[m] = rho[m] + rho0;
            u[m] = (((f[5] - f[7]) + (f[8] - f[6])) + (f[1] - f[3])) / rho[m];
            v[m] = (((f[5] - f[7]) + (f[6] - f[8])) + (f[2] - f[4])) / rho[m];
        }
}

/* sim/sim_lw4.F90:26-115 lw4_stream (order = 4, H = 2) and sim/sim_lw6.F90:26-127 lw6_stream (order = 6, H = 3) */
void SFX(orc_lwh_stream)(int order, int nx, int ny, const REAL *fsrc, REAL *fdst, REAL dt)
{
    static const int cx[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
    static const int cy[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
    const int H = order / 2;
    for (int j = 1; j <= ny; ++j)
        for (int i = 1; i <= nx; ++i) fdst[HIDX(i, j, 0)] = fsrc[HIDX(i, j, 0)];
    for (int k = 1; k < 9; ++k) {
        REAL vx = dt * R(cx[k]), vy = dt * R(cy[k]);
        REAL vxx = R(0.5) * vx * vx, vyy = R(0.5) * vy * vy, vxy = vx * vy;
#define FS(di, dj) fsrc[HIDX(i + (di), j + (dj), k)]
        for (int j = 1; j <= ny; ++j)
            for (int i = 1; i <= nx; ++i) {
                REAL dfx, dfy, dfxx, dfyy, dfxy;
                if (order == 4) {
                    REAL fc = FS(0, 0), fe = FS(1, 0), fw = FS(-1, 0), fee = FS(2, 0), fww = FS(-2, 0);
                    REAL fn = FS(0, 1), fs = FS(0, -1), fnn = FS(0, 2), fss = FS(0, -2);
                    REAL fne = FS(1, 1), fnw = FS(-1, 1), fsw = FS(-1, -1), fse = FS(1, -1);
                    REAL fne2 = FS(2, 2), fnw2 = FS(-2, 2), fsw2 = FS(-2, -2), fse2 = FS(2, -2);
                    dfx = (R(1.0) / R(12.0)) * (fww - fee) + (R(2.0) / R(3.0)) * (fe - fw);
                    dfy = (R(1.0) / R(12.0)) * (fss - fnn) + (R(2.0) / R(3.0)) * (fn - fs);
                    dfxx = -(R(1.0) / R(12.0)) * fww + (R(4.0) / R(3.0)) * fw - (R(5.0) / R(2.0)) * fc + (R(4.0) / R(3.0)) * fe -
                           (R(1.0) / R(12.0)) * fee;
                    dfyy = -(R(1.0) / R(12.0)) * fss + (R(4.0) / R(3.0)) * fs - (R(5.0) / R(2.0)) * fc + (R(4.0) / R(3.0)) * fn -
                           (R(1.0) / R(12.0)) * fnn;
                    dfxy = (R(1.0) / R(3.0)) * (fne - fnw + fsw - fse) - (R(1.0) / R(48.0)) * (fne2 - fnw2 + fsw2 - fse2);
                } else {
                    dfx = (R(1.0) / R(60.0)) * (FS(3, 0) - FS(-3, 0)) - (R(3.0) / R(20.0)) * (FS(2, 0) - FS(-2, 0)) +
                          (R(3.0) / R(4.0)) * (FS(1, 0) - FS(-1, 0));
                    dfy = (R(1.0) / R(60.0)) * (FS(0, 3) - FS(0, -3)) - (R(3.0) / R(20.0)) * (FS(0, 2) - FS(0, -2)) +
                          (R(3.0) / R(4.0)) * (FS(0, 1) - FS(0, -1));
                    /* `1.0_wp/90_wp`: the integer literal 90 of kind wp is promoted to real */
                    dfxx = (R(1.0) / R(90.0)) * (FS(-3, 0) + FS(3, 0)) - (R(3.0) / R(20.0)) * (FS(-2, 0) + FS(2, 0)) +
                           (R(3.0) / R(2.0)) * (FS(-1, 0) + FS(1, 0)) - (R(49.0) / R(18.0)) * (FS(0, 0));
                    dfyy = (R(1.0) / R(90.0)) * (FS(0, -3) + FS(0, 3)) - (R(3.0) / R(20.0)) * (FS(0, -2) + FS(0, 2)) +
                           (R(3.0) / R(2.0)) * (FS(0, -1) + FS(0, 1)) - (R(49.0) / R(18.0)) * (FS(0, 0));
                    dfxy = (R(3.0) / R(8.0)) * (FS(1, 1) - FS(-1, 1) + FS(-1, -1) - FS(1, -1)) -
                           (R(3.0) / R(80.0)) * (FS(2, 2) - FS(-2, 2) + FS(-2, -2) - FS(2, -2)) +
                           (R(1.0) / R(360.0)) * (FS(3, 3) - FS(-3, 3) + FS(-3, -3) - FS(3, -3));
                }
                fdst[HIDX(i, j, k)] = FS(0, 0) - vx * dfx - vy * dfy + (vxx * dfxx + vxy * dfxy + vyy * dfyy);
            }
#undef FS
    }
}

/* lw4_collision / lw6_collision (sim/sim_lw4.F90:118-195): identical to lw_collision, halo width H */
void SFX(orc_lwh_collision)(int nx, int ny, int H, REAL *pdf, REAL omega)
{
    const REAL rho0 = R(1.0);
    const REAL w[9] = {W0, WS, WS, WS, WS, WD, WD, WD, WD};
    for (int j = 1; j <= ny; ++j)
        for (int i = 1; i <= nx; ++i) {
            REAL f[9], feq[9];
            for (int k = 0; k < 9; ++k) f[k] = pdf[HIDX(i, j, k)];
            REAL rho = f[0] + (((f[5] + f[7]) + (f[6] + f[8])) + ((f[1] + f[3]) + (f[2] + f[4]))) + rho0;
            REAL irho = R(1.0) / rho;
            REAL ux = (((f[5] - f[7]) + (f[8] - f[6])) + (f[1] - f[3])) * irho;
            REAL uy = (((f[5] - f[7]) + (f[6] - f[8])) + (f[2] - f[4])) * irho;
            REAL uxx = ux * ux, uyy = uy * uy;
            REAL indp = -R(1.5) * (uxx + uyy);
            feq[0] = W0 * rho * (indp);
            feq[1] = WS * rho * (indp + R(3.0) * ux + R(4.5) * uxx);
            feq[2] = WS * rho * (indp + R(3.0) * uy + R(4.5) * uyy);
            feq[3] = WS * rho * (indp - R(3.0) * ux + R(4.5) * uxx);
            feq[4] = WS * rho * (indp - R(3.0) * uy + R(4.5) * uyy);
            REAL uxpy = ux + uy;
            feq[5] = WD * rho * (indp + R(3.0) * uxpy + R(4.5) * uxpy * uxpy);
            feq[7] = WD * rho * (indp - R(3.0) * uxpy + R(4.5) * uxpy * uxpy);
            REAL uxmy = ux - uy;
            feq[6] = WD * rho * (indp - R(3.0) * uxmy + R(4.5) * uxmy * uxmy);
            feq[8] = WD * rho * (indp + R(3.0) * uxmy + R(4.5) * uxmy * uxmy);
            for (int k = 0; k < 9; ++k) feq[k] = feq[k] + w[k] * (rho - rho0);
            for (int k = 0; k < 9; ++k) pdf[HIDX(i, j, k)] = f[k] + omega * (feq[k] - f[k]);
        }
}

/* lw4_bc (sim/sim_lw4.F90:198-245, the `#if 1` branch) / lw6_bc (sim/sim_lw6.F90): south and north halo
 * rows of the interior columns, then west and east halo columns over ALL rows (corners included) */
void SFX(orc_lwh_bc)(int nx, int ny, int H, REAL *f)
{
    for (int k = 0; k < 9; ++k) {
        for (int h = 0; h < H; ++h)
            for (int i = 1; i <= nx; ++i) {
                f[HIDX(i, -h, k)] = f[HIDX(i, ny - h, k)];
                f[HIDX(i, ny + 1 + h, k)] = f[HIDX(i, 1 + h, k)];
            }
        for (int j = 1 - H; j <= ny + H; ++j)
            for (int h = 0; h < H; ++h) {
                f[HIDX(-h, j, k)] = f[HIDX(nx - h, j, k)];
                f[HIDX(nx + 1 + h, j, k)] = f[HIDX(1 + h, j, k)];
            }
    }
}
/* ---- Heun finite-volume plugin `fvm` (sim/sim_fvm.F90), arrays (-1:nx+2, -1:ny+2, 0:8), H = 2 ---- */
/* fvm_collision, sim/sim_fvm.F90:139-188: the halo layer is collided too; velocity by division */
void SFX(orc_simfvm_collision)(int nx, int ny, REAL *pdf, REAL omega)
{
    const int H = 2;
    const REAL rho0 = R(1.0);
    for (int j = -1; j <= ny + 2; ++j)
        for (int i = -1; i <= nx + 2; ++i) {
            REAL f[9], feq[9];
            for (int k = 0; k < 9; ++k) f[k] = pdf[HIDX(i, j, k)];
            REAL rho = f[0] + (((f[5] + f[7]) + (f[6] + f[8])) + ((f[1] + f[3]) + (f[2] + f[4]))) + rho0;
            REAL ux = (((f[5] - f[7]) + (f[8] - f[6])) + (f[1] - f[3])) / rho;
            REAL uy = (((f[5] - f[7]) + (f[6] - f[8])) + (f[2] - f[4])) / rho;
            SFX(orc_sim_equilibrium)(rho, ux, uy, feq);
            for (int k = 0; k < 9; ++k) pdf[HIDX(i, j, k)] = f[k] + omega * (feq[k] - f[k]);
        }
}

#define FVM_FLUX(a, i, j, k)                                                                                              \
    (cx[k] * (R(0.5) * (a[HIDX((i) + 1, j, k)] + a[HIDX(i, j, k)]) - R(0.5) * (a[HIDX(i, j, k)] + a[HIDX((i)-1, j, k)])) + \
     cy[k] * (R(0.5) * (a[HIDX(i, (j) + 1, k)] + a[HIDX(i, j, k)]) - R(0.5) * (a[HIDX(i, j, k)] + a[HIDX(i, (j)-1, k)])))

/* fvm_predict_hc, sim/sim_fvm.F90:67-100 */
void SFX(orc_simfvm_predict)(int nx, int ny, const REAL *fsrc, REAL *fdst, REAL dt)
{
    const int H = 2;
    const REAL cx[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
    const REAL cy[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
    for (int j = -1; j <= ny + 2; ++j)
        for (int i = -1; i <= nx + 2; ++i) fdst[HIDX(i, j, 0)] = fsrc[HIDX(i, j, 0)];
    for (int k = 1; k < 9; ++k)
        for (int j = 1; j <= ny; ++j)
            for (int i = 1; i <= nx; ++i) {
                REAL flux = FVM_FLUX(fsrc, i, j, k);
                fdst[HIDX(i, j, k)] = fdst[HIDX(i, j, k)] - dt * flux;
            }
}

/* fvm_correct_hc, sim/sim_fvm.F90:103-136 */
void SFX(orc_simfvm_correct)(int nx, int ny, const REAL *fp, const REAL *fsrc, REAL *fdst, REAL dt)
{
    const int H = 2;
    const REAL cx[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
    const REAL cy[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
    for (int k = 1; k < 9; ++k)
        for (int j = 1; j <= ny; ++j)
            for (int i = 1; i <= nx; ++i) {
                REAL flux = FVM_FLUX(fsrc, i, j, k);
                REAL fluxp = FVM_FLUX(fp, i, j, k);
                fdst[HIDX(i, j, k)] = fdst[HIDX(i, j, k)] - dt * R(0.5) * (flux + fluxp);
            }
}
#undef FVM_FLUX

/* sim_fvm%step, sim/sim_fvm.F90:282-322 (fvm_bc :191-218 is orc_lwh_bc with H = 2).  f1, f2, fc are the three
 * haloed buffers; on return f1 holds the new state and fc the half-collided old one (the move_alloc swap is
 * done by exchanging contents through the caller's pointers: `swapped` tells the caller to swap f1 and fc). */
void SFX(orc_simfvm_step)(int nx, int ny, REAL *f1, REAL *f2, REAL *fc, REAL dt, REAL omega)
{
    const size_t n = (size_t)(nx + 4) * (ny + 4) * 9;
    for (size_t i = 0; i < n; ++i) fc[i] = f1[i];
    SFX(orc_simfvm_collision)(nx, ny, fc, omega);
    for (size_t i = 0; i < n; ++i) f2[i] = fc[i];
    SFX(orc_simfvm_collision)(nx, ny, f1, R(0.5) * omega);
    SFX(orc_simfvm_predict)(nx, ny, f1, f2, dt);
    SFX(orc_lwh_bc)(nx, ny, 2, f2);
    SFX(orc_simfvm_collision)(nx, ny, f2, R(0.5) * omega);
    SFX(orc_simfvm_correct)(nx, ny, f2, f1, fc, dt);
    SFX(orc_lwh_bc)(nx, ny, 2, fc);
    /* caller swaps f1 <-> fc */
}
#undef HIDX

#undef SIDX
#undef FIDX
#undef MIDX
#undef R
#undef WRAP_P1
#undef WRAP_M1
#undef W0
#undef WS
#undef WD
#undef CSQR
#undef INVCSQR
#undef ONE_THIRD
