// plbm_math.cuh -- per-node D2Q9 arithmetic (device), sm_100a.
//
// Each function evaluates one node exactly as the reference's Fortran does: same pairing
// of the moment sums, same left-to-right products, one reciprocal + two multiplies for the
// velocity.  The library is compiled with -fmad=false, so every add/mul/div below is an
// individually rounded IEEE operation and results are bit-identical to the non-FMA CPU
// build of the reference (gfortran -O3 on baseline x86-64).
//
// Reference sources (relative to the reference root):
//   equilibrium            src/fvm_bardow.F90:99-126
//   bgk_kernel             src/collision_bgk.F90:35-82
//   kernel_bgk / bgk_cache src/periodic_dugks.F90:80-169, src/collision_bgk.F90:84-176
//   trt_naive              src/collision_trt.F90:13-34, 64-160
//   rr_kernel_naive        src/collision_regularized.F90:40-202
//   update_ew / update_ns  src/periodic_dugks.F90:310-434
//   update_macros_kernel   src/fvm_bardow.F90:356-388
#pragma once
#include <cuda_runtime.h>

#include "plbm_f32x2.cuh"

namespace plbm {

// D2Q9 velocity set, src/fvm_bardow.F90:87-88
__host__ __device__ constexpr int cxi(int q) { return q == 1 || q == 5 || q == 8 ? 1 : (q == 3 || q == 6 || q == 7 ? -1 : 0); }
__host__ __device__ constexpr int cyi(int q) { return q == 2 || q == 5 || q == 6 ? 1 : (q == 4 || q == 7 || q == 8 ? -1 : 0); }

enum Model : int { M_NONE = -1, M_BGK = 0, M_TRT = 1, M_RR = 2, M_BGK_SPLIT = 3, M_TRT_SPLIT = 4, M_BGK_IMPROVED = 5 };

template <typename T> struct K {
    // evaluated in working precision, like the Fortran `parameter`s with _wp literals
    static __host__ __device__ constexpr T w0() { return T(4) / T(9); }
    static __host__ __device__ constexpr T ws() { return T(1) / T(9); }
    static __host__ __device__ constexpr T wd() { return T(1) / T(36); }
    static __host__ __device__ constexpr T csqr() { return T(1) / T(3); }
    static __host__ __device__ constexpr T one_third() { return T(1) / T(3); }
};

// Collision parameters, precomputed on the host in working precision.
template <typename T> struct CollideParams {
    T omega;     // grid%omega              (BGK, RR, split BGK; lambda_even for TRT)
    T lambda_d;  // lambda_d(omega, magic)  (TRT only)
};

// Density of a node in the collision kernels' summation order (f0 added last).
template <typename T> __device__ __forceinline__ T node_rho(const T (&f)[9])
{
    return (((f[5] + f[7]) + (f[6] + f[8])) + ((f[1] + f[3]) + (f[2] + f[4]))) + f[0];
}

// The collisions that divide by the density take an optional `inv` (template parameter PRE = true): the reciprocal
// T(1) / node_rho(f), computed by the caller.
// Same division on the same operand, so nothing changes in the result; a kernel that collides several independent nodes per
// thread uses it to issue all the divisions first -- the IEEE division expands into a fast path plus a branch to a slow path, and
// that branch otherwise ends the basic block, so that the instruction streams of the nodes could not be interleaved.

// PRE is a template parameter, not a run-time test: with PRE = false (every caller but the three-level kernel of plbm_lbm3w.cu)
// each function instantiates to exactly the statements it had before `inv` existed.

// rho, ux, uy with the collision kernels' summation order (f0 added last).
template <typename T, bool PRE = false>
__device__ __forceinline__ void moments(const T (&f)[9], T& rho, T& ux, T& uy, const T* inv = nullptr)
{
    rho = (((f[5] + f[7]) + (f[6] + f[8])) + ((f[1] + f[3]) + (f[2] + f[4]))) + f[0];
    T invrho;
    if constexpr (PRE) invrho = *inv;
    else invrho = T(1) / rho;
    ux = invrho * (((f[5] - f[7]) + (f[8] - f[6])) + (f[1] - f[3]));
    uy = invrho * (((f[5] - f[7]) + (f[6] - f[8])) + (f[2] - f[4]));
}

// update_macros_kernel order: f0 added first (bitwise the same sum, kept literal).
template <typename T> __device__ __forceinline__ void macros(const T (&f)[9], T& rho, T& ux, T& uy)
{
    rho = f[0] + (((f[5] + f[7]) + (f[6] + f[8])) + ((f[1] + f[3]) + (f[2] + f[4])));
    T invrho = T(1) / rho;
    ux = invrho * (((f[5] - f[7]) + (f[8] - f[6])) + (f[1] - f[3]));
    uy = invrho * (((f[5] - f[7]) + (f[6] - f[8])) + (f[2] - f[4]));
}

template <typename T> __device__ __forceinline__ void equilibrium(T rho, T ux, T uy, T (&feq)[9])
{
    const T w0 = K<T>::w0(), ws = K<T>::ws(), wd = K<T>::wd();
    T uxx = ux * ux;
    T uyy = uy * uy;
    T indp = T(1) - T(1.5) * (uxx + uyy);
    feq[0] = w0 * rho * (indp);
    feq[1] = ws * rho * (indp + T(3) * ux + T(4.5) * uxx);
    feq[2] = ws * rho * (indp + T(3) * uy + T(4.5) * uyy);
    feq[3] = ws * rho * (indp - T(3) * ux + T(4.5) * uxx);
    feq[4] = ws * rho * (indp - T(3) * uy + T(4.5) * uyy);
    T uxpy = ux + uy;
    feq[5] = wd * rho * (indp + T(3) * uxpy + T(4.5) * uxpy * uxpy);
    feq[7] = wd * rho * (indp - T(3) * uxpy + T(4.5) * uxpy * uxpy);
    T uxmy = ux - uy;
    feq[6] = wd * rho * (indp - T(3) * uxmy + T(4.5) * uxmy * uxmy);
    feq[8] = wd * rho * (indp + T(3) * uxmy + T(4.5) * uxmy * uxmy);
}

template <typename T, bool PRE = false> __device__ __forceinline__ void collide_bgk(T (&f)[9], T omega, const T* inv = nullptr)
{
    T omegabar = T(1) - omega;
    T rho, ux, uy, feq[9];
    moments<T, PRE>(f, rho, ux, uy, inv);
    equilibrium(rho, ux, uy, feq);
#pragma unroll
    for (int q = 0; q < 9; ++q) f[q] = omegabar * f[q] + omega * feq[q];
}

// Re-associated BGK of kernel_bgk / bgk_kernel_cache (numerically different from collide_bgk).
template <typename T, bool PRE = false> __device__ __forceinline__ void collide_bgk_split(T (&f)[9], T omega, const T* inv = nullptr)
{
    const T w0 = K<T>::w0(), ws = K<T>::ws(), wd = K<T>::wd();
    T omegabar = T(1) - omega;
    T omega_w0 = T(3) * omega * w0;
    T omega_ws = T(3) * omega * ws;
    T omega_wd = T(3) * omega * wd;
    T rho, ux, uy;
    moments<T, PRE>(f, rho, ux, uy, inv);
    T indp = K<T>::one_third() - T(0.5) * (ux * ux + uy * uy);

    f[0] = omegabar * f[0] + omega_w0 * rho * indp;
    T vel_trm_13 = indp + T(1.5) * ux * ux;
    f[1] = omegabar * f[1] + omega_ws * rho * (vel_trm_13 + ux);
    f[3] = omegabar * f[3] + omega_ws * rho * (vel_trm_13 - ux);
    T vel_trm_24 = indp + T(1.5) * uy * uy;
    f[2] = omegabar * f[2] + omega_ws * rho * (vel_trm_24 + uy);
    f[4] = omegabar * f[4] + omega_ws * rho * (vel_trm_24 - uy);
    T velxpy = ux + uy;
    T vel_trm_57 = indp + T(1.5) * velxpy * velxpy;
    f[5] = omegabar * f[5] + omega_wd * rho * (vel_trm_57 + velxpy);
    f[7] = omegabar * f[7] + omega_wd * rho * (vel_trm_57 - velxpy);
    T velxmy = ux - uy;
    T vel_trm_68 = indp + T(1.5) * velxmy * velxmy;
    f[6] = omegabar * f[6] + omega_wd * rho * (vel_trm_68 - velxmy);
    f[8] = omegabar * f[8] + omega_wd * rho * (vel_trm_68 + velxmy);
}

// DUGKS face relaxation: moments of all nine face values, relax only the flux carriers.
// EW = true: update_ew (1,3,5,7,6,8); EW = false: update_ns (2,4,5,7,6,8).
template <typename T, bool EW> __device__ __forceinline__ void face_relax(T (&f)[9], T omega)
{
    const T ws = K<T>::ws(), wd = K<T>::wd();
    T omegabar = T(1) - omega;
    T omega_ws = T(3) * omega * ws;
    T omega_wd = T(3) * omega * wd;
    T rho, ux, uy;
    moments(f, rho, ux, uy);
    T indp = K<T>::one_third() - T(0.5) * (ux * ux + uy * uy);
    if (EW) {
        T vel_trm_13 = indp + T(1.5) * ux * ux;
        f[1] = omegabar * f[1] + omega_ws * rho * (vel_trm_13 + ux);
        f[3] = omegabar * f[3] + omega_ws * rho * (vel_trm_13 - ux);
    } else {
        T vel_trm_24 = indp + T(1.5) * uy * uy;
        f[2] = omegabar * f[2] + omega_ws * rho * (vel_trm_24 + uy);
        f[4] = omegabar * f[4] + omega_ws * rho * (vel_trm_24 - uy);
    }
    T velxpy = ux + uy;
    T vel_trm_57 = indp + T(1.5) * velxpy * velxpy;
    f[5] = omegabar * f[5] + omega_wd * rho * (vel_trm_57 + velxpy);
    f[7] = omegabar * f[7] + omega_wd * rho * (vel_trm_57 - velxpy);
    T velxmy = ux - uy;
    T vel_trm_68 = indp + T(1.5) * velxmy * velxmy;
    f[6] = omegabar * f[6] + omega_wd * rho * (vel_trm_68 - velxmy);
    f[8] = omegabar * f[8] + omega_wd * rho * (vel_trm_68 + velxmy);
}

// Two-relaxation-time, incompressible equilibrium (velocity = momentum).
template <typename T> __device__ __forceinline__ void collide_trt(T (&f)[9], T lambda_e, T lambda_d)
{
    // compile-time constants, evaluated in the scalar working precision (S = T except for the packed pair type F2)
    typedef typename scalar_of<T>::type S;
    const S t1x2_s = (S(1) / S(9)) * S(2);
    const S t2x2_s = (S(1) / S(36)) * S(2);
    const S inv2csq2_s = S(1) / (S(2) * (S(1) / S(3)) * (S(1) / S(3)));
    const T t0 = T(S(4) / S(9));
    const T t1x2 = T(t1x2_s);
    const T t2x2 = T(t2x2_s);
    const T fac1 = T(t1x2_s * inv2csq2_s);
    const T fac2 = T(t2x2_s * inv2csq2_s);
    const T three_t1x2 = T(S(3) * t1x2_s);
    const T three_t2x2 = T(S(3) * t2x2_s);

    T lambda_e_scaled = T(0.5) * lambda_e;
    T lambda_d_scaled = T(0.5) * lambda_d;

    T vC = f[0], vE = f[1], vN = f[2], vW = f[3], vS = f[4];
    T vNE = f[5], vNW = f[6], vSW = f[7], vSE = f[8];

    T rho = (((vNE + vSW) + (vNW + vSE)) + ((vE + vW) + (vN + vS))) + vC;
    T velX = (((vNE - vSW) + (vSE - vNW)) + (vE - vW));
    T velY = (((vNE - vSW) + (vNW - vSE)) + (vN - vS));
    T velX2 = velX * velX;
    T velY2 = velY * velY;
    T feq_common = rho - T(1.5) * (velX2 + velY2);

    f[0] = vC * (T(1) - lambda_e) + lambda_e * t0 * feq_common;

    T velXPY = velX + velY;
    T sym_NE_SW = lambda_e_scaled * (vNE + vSW - fac2 * velXPY * velXPY - t2x2 * feq_common);
    T asym_NE_SW = lambda_d_scaled * (vNE - vSW - three_t2x2 * velXPY);
    f[5] = vNE - sym_NE_SW - asym_NE_SW;
    f[7] = vSW - sym_NE_SW + asym_NE_SW;

    T velXMY = velX - velY;
    T sym_SE_NW = lambda_e_scaled * (vSE + vNW - fac2 * velXMY * velXMY - t2x2 * feq_common);
    T asym_SE_NW = lambda_d_scaled * (vSE - vNW - three_t2x2 * velXMY);
    f[8] = vSE - sym_SE_NW - asym_SE_NW;
    f[6] = vNW - sym_SE_NW + asym_SE_NW;

    T sym_N_S = lambda_e_scaled * (vN + vS - fac1 * velY2 - t1x2 * feq_common);
    T asym_N_S = lambda_d_scaled * (vN - vS - three_t1x2 * velY);
    f[2] = vN - sym_N_S - asym_N_S;
    f[4] = vS - sym_N_S + asym_N_S;

    T sym_E_W = lambda_e_scaled * (vE + vW - fac1 * velX2 - t1x2 * feq_common);
    T asym_E_W = lambda_d_scaled * (vE - vW - three_t1x2 * velX);
    f[1] = vE - sym_E_W - asym_E_W;
    f[3] = vW - sym_E_W + asym_E_W;
}

// Recursive-regularized collision with 3rd/4th-order Hermite equilibrium.
template <typename T, bool PRE = false> __device__ __forceinline__ void collide_rr(T (&f)[9], T omega, const T* inv = nullptr)
{
    const T w0 = K<T>::w0(), ws = K<T>::ws(), wd = K<T>::wd(), csqr = K<T>::csqr();
    T omega_w0 = w0 * (T(1) - omega);
    T omega_ws = ws * (T(1) - omega);
    T omega_wd = wd * (T(1) - omega);

    T vC = f[0], vE = f[1], vN = f[2], vW = f[3], vS = f[4];
    T vNE = f[5], vNW = f[6], vSW = f[7], vSE = f[8];
    T feq[9];

    T rho = (((vNE + vSW) + (vNW + vSE)) + ((vE + vW) + (vN + vS))) + vC;
    T invrho;
    if constexpr (PRE) invrho = *inv;
    else invrho = T(1) / rho;
    T ux = invrho * (((vNE - vSW) + (vSE - vNW)) + (vE - vW));
    T uy = invrho * (((vNE - vSW) + (vNW - vSE)) + (vN - vS));

    T uxx = ux * ux;
    T uyy = uy * uy;
    T uxxy = uxx * uy;
    T uyyx = uyy * ux;
    T uxxyy = uxx * uyy;

    T indp0 = T(1) - T(1.5) * (uxx + uyy);
    T indps = indp0 - T(4.5) * uxxyy;
    T indpd = indp0 + T(9) * uxxyy;
    indp0 = indp0 + T(2.25) * uxxyy;

    feq[0] = w0 * rho * indp0;
    feq[1] = ws * rho * (indps + T(3) * ux + T(4.5) * (uxx - uyyx));
    feq[3] = ws * rho * (indps - T(3) * ux + T(4.5) * (uxx + uyyx));
    feq[2] = ws * rho * (indps + T(3) * uy + T(4.5) * (uyy - uxxy));
    feq[4] = ws * rho * (indps - T(3) * uy + T(4.5) * (uyy + uxxy));

    vC = vC - feq[0];
    vE = vE - feq[1];
    vN = vN - feq[2];
    vW = vW - feq[3];
    vS = vS - feq[4];

    T axx = csqr * (T(2) * (vE + vW) - (vN + vS) - vC);
    T ayy = csqr * (T(2) * (vN + vS) - (vE + vW) - vC);

    T u3p = uxxy + uyyx;
    T uxpy = ux + uy;
    T indp57 = indpd + T(4.5) * uxpy * uxpy;
    feq[5] = wd * rho * (indp57 + T(3) * uxpy + T(9) * u3p);
    feq[7] = wd * rho * (indp57 - T(3) * uxpy - T(9) * u3p);

    T u3m = uxxy - uyyx;
    T uxmy = ux - uy;
    T indp68 = indpd + T(4.5) * uxmy * uxmy;
    feq[6] = wd * rho * (indp68 - T(3) * uxmy + T(9) * u3m);
    feq[8] = wd * rho * (indp68 + T(3) * uxmy - T(9) * u3m);

    vNE = vNE - feq[5];
    vNW = vNW - feq[6];
    vSW = vSW - feq[7];
    vSE = vSE - feq[8];

    T tmp = T(2) * csqr * (vNE + vNW + vSW + vSE);
    axx = axx + tmp;
    ayy = ayy + tmp;
    T axy = ((vNE + vSW) - (vNW + vSE));

    T axxy = T(2) * ux * axy + uy * axx;
    T ayyx = T(2) * uy * axy + ux * ayy;
    T axxyy = T(2) * (ux * ayyx + uy * axxy) - uxx * ayy - uyy * axx - T(4) * ux * uy * axy;

    indp0 = -T(1.5) * (axx + ayy);
    indps = indp0 - T(4.5) * axxyy;
    indpd = T(9) * axxyy - T(2) * indp0;
    indp0 = indp0 + T(2.25) * axxyy;

    vC = indp0;
    vE = indps + T(4.5) * (axx - ayyx);
    vW = indps + T(4.5) * (axx + ayyx);
    vN = indps + T(4.5) * (ayy - axxy);
    vS = indps + T(4.5) * (ayy + axxy);
    vNE = indpd + T(9) * (axxy + ayyx + axy);
    vSW = indpd - T(9) * (axxy + ayyx - axy);
    vNW = indpd + T(9) * (axxy - ayyx - axy);
    vSE = indpd - T(9) * (axxy - ayyx + axy);

    f[0] = feq[0] + omega_w0 * vC;
    f[1] = feq[1] + omega_ws * vE;
    f[2] = feq[2] + omega_ws * vN;
    f[3] = feq[3] + omega_ws * vW;
    f[4] = feq[4] + omega_ws * vS;
    f[5] = feq[5] + omega_wd * vNE;
    f[6] = feq[6] + omega_wd * vNW;
    f[7] = feq[7] + omega_wd * vSW;
    f[8] = feq[8] + omega_wd * vSE;
}

// trt_split (-DSPLIT, src/collision_trt.F90:162-290): the column-blocked variant.  Per node it differs
// from trt_naive only in the axis pairs, where `fac1 * vel * vel` is (fac1*vel)*vel instead of
// fac1*(vel*vel) -- a last-bit difference, reproduced.
template <typename T> __device__ __forceinline__ void collide_trt_split(T (&f)[9], T lambda_e, T lambda_d)
{
    // compile-time constants, evaluated in the scalar working precision (S = T except for the packed pair type F2)
    typedef typename scalar_of<T>::type S;
    const S t1x2_s = (S(1) / S(9)) * S(2);
    const S t2x2_s = (S(1) / S(36)) * S(2);
    const S inv2csq2_s = S(1) / (S(2) * (S(1) / S(3)) * (S(1) / S(3)));
    const T t0 = T(S(4) / S(9));
    const T t1x2 = T(t1x2_s);
    const T t2x2 = T(t2x2_s);
    const T fac1 = T(t1x2_s * inv2csq2_s);
    const T fac2 = T(t2x2_s * inv2csq2_s);
    const T three_t1x2 = T(S(3) * t1x2_s);
    const T three_t2x2 = T(S(3) * t2x2_s);
    T lambda_e_scaled = T(0.5) * lambda_e;
    T lambda_d_scaled = T(0.5) * lambda_d;

    T vC = f[0], vE = f[1], vN = f[2], vW = f[3], vS = f[4];
    T vNE = f[5], vNW = f[6], vSW = f[7], vSE = f[8];
    T rho = (((vNE + vSW) + (vNW + vSE)) + ((vE + vW) + (vN + vS))) + vC;
    T velX = (((vNE - vSW) + (vSE - vNW)) + (vE - vW));
    T velY = (((vNE - vSW) + (vNW - vSE)) + (vN - vS));
    T feq_common = rho - T(1.5) * (velX * velX + velY * velY);
    f[0] = vC * (T(1) - lambda_e) + lambda_e * t0 * feq_common;

    T velXPY = velX + velY;
    T sym_NE_SW = lambda_e_scaled * (vNE + vSW - fac2 * velXPY * velXPY - t2x2 * feq_common);
    T asym_NE_SW = lambda_d_scaled * (vNE - vSW - three_t2x2 * velXPY);
    f[5] = vNE - sym_NE_SW - asym_NE_SW;
    f[7] = vSW - sym_NE_SW + asym_NE_SW;

    T velXMY = velX - velY;
    T sym_SE_NW = lambda_e_scaled * (vSE + vNW - fac2 * velXMY * velXMY - t2x2 * feq_common);
    T asym_SE_NW = lambda_d_scaled * (vSE - vNW - three_t2x2 * velXMY);
    f[8] = vSE - sym_SE_NW - asym_SE_NW;
    f[6] = vNW - sym_SE_NW + asym_SE_NW;

    T sym_N_S = lambda_e_scaled * (vN + vS - fac1 * velY * velY - t1x2 * feq_common);
    T asym_N_S = lambda_d_scaled * (vN - vS - three_t1x2 * velY);
    f[2] = vN - sym_N_S - asym_N_S;
    f[4] = vS - sym_N_S + asym_N_S;

    T sym_E_W = lambda_e_scaled * (vE + vW - fac1 * velX * velX - t1x2 * feq_common);
    T asym_E_W = lambda_d_scaled * (vE - vW - three_t1x2 * velX);
    f[1] = vE - sym_E_W - asym_E_W;
    f[3] = vW - sym_E_W + asym_E_W;
}

// collide_bgk_improved (src/collision_bgk_improved.f90:24-107): product-form BGK with a cubic
// Galilean-invariance correction.  The reference kernel ignores the padded leading dimension
// (identical when ny is a multiple of 16); the intended per-node arithmetic is kept.
template <typename T, bool PRE = false> __device__ __forceinline__ void collide_bgk_improved(T (&f)[9], T omega, const T* inv = nullptr)
{
    const T one_third = T(1) / T(3), two_thirds = T(2) / T(3);
    T fac = T(4.5) - T(2.25) * omega;
    T omegabar = T(1) - omega;
    T vC = f[0], vE = f[1], vN = f[2], vW = f[3], vS = f[4];
    T vNE = f[5], vNW = f[6], vSW = f[7], vSE = f[8];

    T rho = (((vNE + vSW) + (vNW + vSE)) + ((vE + vW) + (vN + vS))) + vC;
    T invrho;
    if constexpr (PRE) invrho = *inv;
    else invrho = T(1) / rho;
    T sumX1 = vE + vNE + vSE;
    T sumXN = vW + vNW + vSW;
    T sumY1 = vN + vNE + vNW;
    T sumYN = vS + vSE + vSW;
    T m10 = invrho * (sumX1 - sumXN);
    T m01 = invrho * (sumY1 - sumYN);
    T u2 = m10 * m10;
    T v2 = m01 * m01;
    T m20 = invrho * (sumX1 + sumXN);
    T m02 = invrho * (sumY1 + sumYN);
    T Gx = fac * u2 * (m20 - one_third - u2);
    T Gy = fac * v2 * (m02 - one_third - v2);
    T X0 = -two_thirds + u2 + Gx;
    T X1 = -(X0 + T(1) + m10) * T(0.5);
    T XN = X1 + m10;
    T Y0 = -two_thirds + v2 + Gy;
    T Y1 = -(Y0 + T(1) + m01) * T(0.5);
    T YN = Y1 + m01;
    T rho_omega = rho * omega;
    X0 = X0 * rho_omega;
    X1 = X1 * rho_omega;
    XN = XN * rho_omega;
    f[0] = omegabar * vC + X0 * Y0;
    f[1] = omegabar * vE + X1 * Y0;
    f[2] = omegabar * vN + X0 * Y1;
    f[3] = omegabar * vW + XN * Y0;
    f[4] = omegabar * vS + X0 * YN;
    f[5] = omegabar * vNE + X1 * Y1;
    f[6] = omegabar * vNW + XN * Y1;
    f[7] = omegabar * vSW + XN * YN;
    f[8] = omegabar * vSE + X1 * YN;
}

template <typename T, int MODEL, bool PRE = false>
__device__ __forceinline__ void collide(T (&f)[9], const CollideParams<T>& p, const T* inv = nullptr)
{
    if (MODEL == M_BGK) collide_bgk<T, PRE>(f, p.omega, inv);
    else if (MODEL == M_TRT) collide_trt(f, p.omega, p.lambda_d);
    else if (MODEL == M_RR) collide_rr<T, PRE>(f, p.omega, inv);
    else if (MODEL == M_BGK_SPLIT) collide_bgk_split<T, PRE>(f, p.omega, inv);
    else if (MODEL == M_TRT_SPLIT) collide_trt_split(f, p.omega, p.lambda_d);
    else if (MODEL == M_BGK_IMPROVED) collide_bgk_improved<T, PRE>(f, p.omega, inv);
}
// does collide<T, MODEL> divide by the density (i.e. does it take `inv`)
__host__ __device__ constexpr bool model_divides(int model) { return model != M_TRT && model != M_TRT_SPLIT && model != M_NONE; }

// Collision of the V vertically adjacent nodes a thread owns.  PACKED (fp32 only, V even): two nodes per instruction
// through the packed pair type F2 (plbm_f32x2.cuh) -- the same individually rounded operations on the same operands, so
// bit-identical to the scalar loop.
template <typename T, int MODEL, int V, bool PACKED> struct CollideNodes {
    static __device__ __forceinline__ void run(T (&n)[V][9], const CollideParams<T>& p)
    {
#pragma unroll
        for (int v = 0; v < V; ++v) collide<T, MODEL>(n[v], p);
    }
};
template <int MODEL, int V> struct CollideNodes<float, MODEL, V, true> {
    static __device__ __forceinline__ void run(float (&n)[V][9], const CollideParams<float>& p)
    {
        static_assert(V % 2 == 0, "packed fp32 collisions take the nodes in pairs");
        CollideParams<F2> p2;
        p2.omega = F2(p.omega);
        p2.lambda_d = F2(p.lambda_d);
#pragma unroll
        for (int v = 0; v < V; v += 2) {
            F2 f[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) f[q] = F2(n[v][q], n[v + 1][q]);
            collide<F2, MODEL>(f, p2);
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                n[v][q] = f[q].lo();
                n[v + 1][q] = f[q].hi();
            }
        }
    }
};
template <typename T, int MODEL, int V, bool PACKED> __device__ __forceinline__ void collide_nodes(T (&n)[V][9], const CollideParams<T>& p)
{
    CollideNodes<T, MODEL, V, PACKED>::run(n, p);
}

// The same in two parts: node_reciprocals() issues the divisions of the V nodes (nothing for a collision that does not divide),
// collide_nodes(n, p, inv) does the rest -- bit-identical to collide_nodes(n, p), see the note at node_rho.
template <typename T, int MODEL, int V> __device__ __forceinline__ void node_reciprocals(const T (&n)[V][9], T (&inv)[V])
{
#pragma unroll
    for (int v = 0; v < V; ++v) inv[v] = model_divides(MODEL) ? T(1) / node_rho(n[v]) : T(0);
}
template <typename T, int MODEL, int V, bool PACKED>
__device__ __forceinline__ void collide_nodes(T (&n)[V][9], const CollideParams<T>& p, const T (&inv)[V])
{
    if constexpr (PACKED && sizeof(T) == 4 && (V % 2) == 0) {
        CollideParams<F2> p2;
        p2.omega = F2(p.omega);
        p2.lambda_d = F2(p.lambda_d);
#pragma unroll
        for (int v = 0; v < V; v += 2) {
            F2 f[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) f[q] = F2(n[v][q], n[v + 1][q]);
            const F2 i2(inv[v], inv[v + 1]);
            collide<F2, MODEL, true>(f, p2, &i2);
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                n[v][q] = f[q].lo();
                n[v + 1][q] = f[q].hi();
            }
        }
    } else {
#pragma unroll
        for (int v = 0; v < V; ++v) collide<T, MODEL, true>(n[v], p, &inv[v]);
    }
}

// 2nd-order, half-step back-traced face reconstruction of one population from its 3x3
// neighbourhood: src/fvm_bardow.F90:449-473 == src/periodic_dugks.F90:238-263.
template <typename T>
__device__ __forceinline__ void faces(T fc, T fe, T fn, T fw, T fs, T fne, T fnw, T fsw, T fse, T cxq, T cyq,
                                      T& cfw, T& cfn, T& cfe, T& cfs)
{
    const T p2 = T(0.5), p8 = T(0.125);
    cfw = p2 * (fc + fw) - p2 * cxq * (fc - fw) - p8 * cyq * (fnw + fn - fsw - fs);
    cfn = p2 * (fc + fn) - p2 * cyq * (fn - fc) - p8 * cxq * (fne + fe - fnw - fw);
    cfe = p2 * (fc + fe) - p2 * cxq * (fe - fc) - p8 * cyq * (fne + fn - fse - fs);
    cfs = p2 * (fc + fs) - p2 * cyq * (fc - fs) - p8 * cxq * (fse + fe - fsw - fw);
}

}  // namespace plbm
