// plbm_fv.cuh -- per-node finite-volume flux update from a shared-memory tile (device), shared by
// the tile kernels of plbm_fvm.cu (plain loads) and plbm_fvm_tma.cu (TMA + mbarrier pipeline).
//   face reconstruction   src/fvm_bardow.F90:449-473 == src/periodic_dugks.F90:238-263
//   update_ew / update_ns src/periodic_dugks.F90:310-434
//   flux update           src/periodic_dugks.F90:282-300, src/fvm_bardow.F90:497
// Tile layout: t[q * PLANE + sx * PITCH + sy]; the centre pointer c0 addresses (q = 0, own node).
#pragma once
#include "plbm_math.cuh"

namespace plbm {

// West/East faces of population Q from the shared tile; c points at the centre node, neighbours
// are at c[dx * PITCH + dy].  hx = p2*cxq, ey = p8*cyq (products formed left to right like the
// Fortran `p2*cxq*(...)`).
template <typename T, int Q, int PITCH> __device__ __forceinline__ void faces_ew(const T* c, T hx, T ey, T& cfw, T& cfe)
{
    constexpr int CX = cxi(Q), CY = cyi(Q);
    const T p2 = T(0.5);
    const T fc = c[0], fw = c[-PITCH], fe = c[PITCH];
    cfw = p2 * (fc + fw);
    cfe = p2 * (fc + fe);
    if (CX != 0) {
        cfw = cfw - hx * (fc - fw);
        cfe = cfe - hx * (fe - fc);
    }
    if (CY != 0) {
        const T fn = c[1], fs = c[-1];
        const T fnw = c[-PITCH + 1], fsw = c[-PITCH - 1], fne = c[PITCH + 1], fse = c[PITCH - 1];
        cfw = cfw - ey * (fnw + fn - fsw - fs);
        cfe = cfe - ey * (fne + fn - fse - fs);
    }
}

// North/South faces; hy = p2*cyq, ex = p8*cxq.
template <typename T, int Q, int PITCH> __device__ __forceinline__ void faces_ns(const T* c, T hy, T ex, T& cfn, T& cfs)
{
    constexpr int CX = cxi(Q), CY = cyi(Q);
    const T p2 = T(0.5);
    const T fc = c[0], fn = c[1], fs = c[-1];
    cfn = p2 * (fc + fn);
    cfs = p2 * (fc + fs);
    if (CY != 0) {
        cfn = cfn - hy * (fn - fc);
        cfs = cfs - hy * (fc - fs);
    }
    if (CX != 0) {
        const T fe = c[PITCH], fw = c[-PITCH];
        const T fne = c[PITCH + 1], fnw = c[-PITCH + 1], fse = c[PITCH - 1], fsw = c[-PITCH - 1];
        cfn = cfn - ex * (fne + fe - fnw - fw);
        cfs = cfs - ex * (fse + fe - fsw - fw);
    }
}

template <typename T, int Q, int PITCH, int PLANE>
__device__ __forceinline__ void ew_pop(const T* c0, T dt, T (&cfw)[9], T (&cfe)[9])
{
    const T cxq = dt * T(cxi(Q)), cyq = dt * T(cyi(Q));
    faces_ew<T, Q, PITCH>(c0 + Q * PLANE, T(0.5) * cxq, T(0.125) * cyq, cfw[Q], cfe[Q]);
}
template <typename T, int Q, int PITCH, int PLANE>
__device__ __forceinline__ void ns_pop(const T* c0, T dt, T (&cfn)[9], T (&cfs)[9])
{
    const T cxq = dt * T(cxi(Q)), cyq = dt * T(cyi(Q));
    faces_ns<T, Q, PITCH>(c0 + Q * PLANE, T(0.5) * cyq, T(0.125) * cxq, cfn[Q], cfs[Q]);
}

// Flux update of one node from the shared tile of fbar (DUGKS) or f^n (Bardow):
//   fp(q) = fp(q) - cxq*(cfe - cfw) - cyq*(cfn - cfs)        (src/periodic_dugks.F90:297)
// evaluated as two passes (east/west, then north/south) so only two face sets are live.
template <typename T, bool DUGKS, int PITCH, int PLANE> __device__ __forceinline__ void flux_update(const T* c0, T dt, T omega_face, T (&fp)[9])
{
    {
        T cfw[9], cfe[9];
        if (DUGKS) ew_pop<T, 0, PITCH, PLANE>(c0, dt, cfw, cfe);  // rest population: faces only feed the moments
        ew_pop<T, 1, PITCH, PLANE>(c0, dt, cfw, cfe);
        ew_pop<T, 3, PITCH, PLANE>(c0, dt, cfw, cfe);
        ew_pop<T, 5, PITCH, PLANE>(c0, dt, cfw, cfe);
        ew_pop<T, 6, PITCH, PLANE>(c0, dt, cfw, cfe);
        ew_pop<T, 7, PITCH, PLANE>(c0, dt, cfw, cfe);
        ew_pop<T, 8, PITCH, PLANE>(c0, dt, cfw, cfe);
        if (DUGKS) {
            ew_pop<T, 2, PITCH, PLANE>(c0, dt, cfw, cfe);
            ew_pop<T, 4, PITCH, PLANE>(c0, dt, cfw, cfe);
            face_relax<T, true>(cfw, omega_face);
            face_relax<T, true>(cfe, omega_face);
        }
#pragma unroll
        for (int q = 1; q < 9; ++q)
            if (cxi(q) != 0) fp[q] = fp[q] - (dt * T(cxi(q))) * (cfe[q] - cfw[q]);
    }
    {
        T cfn[9], cfs[9];
        if (DUGKS) ns_pop<T, 0, PITCH, PLANE>(c0, dt, cfn, cfs);
        ns_pop<T, 2, PITCH, PLANE>(c0, dt, cfn, cfs);
        ns_pop<T, 4, PITCH, PLANE>(c0, dt, cfn, cfs);
        ns_pop<T, 5, PITCH, PLANE>(c0, dt, cfn, cfs);
        ns_pop<T, 6, PITCH, PLANE>(c0, dt, cfn, cfs);
        ns_pop<T, 7, PITCH, PLANE>(c0, dt, cfn, cfs);
        ns_pop<T, 8, PITCH, PLANE>(c0, dt, cfn, cfs);
        if (DUGKS) {
            ns_pop<T, 1, PITCH, PLANE>(c0, dt, cfn, cfs);
            ns_pop<T, 3, PITCH, PLANE>(c0, dt, cfn, cfs);
            face_relax<T, false>(cfn, omega_face);
            face_relax<T, false>(cfs, omega_face);
        }
#pragma unroll
        for (int q = 1; q < 9; ++q)
            if (cyi(q) != 0) fp[q] = fp[q] - (dt * T(cyi(q))) * (cfn[q] - cfs[q]);
    }
}

// stream_fdm_bardow, default build (src/fvm_bardow.F90:511-685): second-order Lax-Wendroff update with
// plain central differences,  fnew = fc - cxq*dfx - cyq*dfy + (cxxq*dfxx + cxyq*dfxy + cyyq*dfyy).
// Terms that carry an exact-zero factor (cx = 0 or cy = 0) are skipped.
template <typename T, int Q, int PITCH> __device__ __forceinline__ T fdm_bardow_pop(const T* c, T dt)
{
    constexpr int CX = cxi(Q), CY = cyi(Q);
    const T cxq = dt * T(CX), cyq = dt * T(CY);
    const T fc = c[0];
    T r = fc;
    if (CX != 0) r = r - cxq * (T(0.5) * (c[PITCH] - c[-PITCH]));
    if (CY != 0) r = r - cyq * (T(0.5) * (c[1] - c[-1]));
    T paren;
    if (CX != 0 && CY != 0) {
        const T cxxq = T(0.5) * cxq * cxq, cyyq = T(0.5) * cyq * cyq, cxyq = cxq * cyq;
        const T dfxx = c[PITCH] - T(2) * fc + c[-PITCH];
        const T dfyy = c[1] - T(2) * fc + c[-1];
        const T dfxy = T(0.25) * (c[PITCH + 1] - c[PITCH - 1] - c[-PITCH + 1] + c[-PITCH - 1]);
        paren = cxxq * dfxx + cxyq * dfxy + cyyq * dfyy;
    } else if (CX != 0) {
        paren = (T(0.5) * cxq * cxq) * (c[PITCH] - T(2) * fc + c[-PITCH]);
    } else {
        paren = (T(0.5) * cyq * cyq) * (c[1] - T(2) * fc + c[-1]);
    }
    return r + paren;
}

// stream_fdm_sofonea (src/fvm_bardow.F90:688-893): one-dimensional Lax-Wendroff along each
// characteristic, fu = f(x + c), fd = f(x - c); the diagonals carry the 1/sqrt(2) grid spacing.
template <typename T, int Q, int PITCH> __device__ __forceinline__ T fdm_sofonea_pop(const T* c, T dt)
{
    constexpr int CX = cxi(Q), CY = cyi(Q);
    const T fc = c[0], fu = c[CX * PITCH + CY], fd = c[-CX * PITCH - CY];
    T du1, du2;
    if (CX == 0 || CY == 0) {
        du1 = T(0.5) * (fu - fd);
        du2 = fu - T(2) * fc + fd;
    } else {
        const T p2 = T(0.5) / sqrt(T(2));
        du1 = p2 * (fu - fd);
        du2 = T(0.5) * (fu - T(2) * fc + fd);
    }
    return fc + dt * (T(0.5) * dt * du2 - du1);
}

template <typename T, bool SOFONEA, int PITCH, int PLANE> __device__ __forceinline__ void fdm_update(const T* c0, T dt, T (&fp)[9])
{
#define PLBM_FDM_Q(Q) fp[Q] = SOFONEA ? fdm_sofonea_pop<T, Q, PITCH>(c0 + Q * PLANE, dt) : fdm_bardow_pop<T, Q, PITCH>(c0 + Q * PLANE, dt)
    PLBM_FDM_Q(1);
    PLBM_FDM_Q(2);
    PLBM_FDM_Q(3);
    PLBM_FDM_Q(4);
    PLBM_FDM_Q(5);
    PLBM_FDM_Q(6);
    PLBM_FDM_Q(7);
    PLBM_FDM_Q(8);
#undef PLBM_FDM_Q
}

// stream_fdm_bardow built with -DFDM_WLS (STENCIL 1), -DFDM_WLS_GAUSS_V1 (2), -DFDM_WLS_GAUSS_V2 (3) or -DFDM_ISO (4):
// the cpp alternatives of the derivative stencils, src/fvm_bardow.F90:591-660, evaluated in full in the
// reference's operation order (every population uses all eight neighbours here).
template <typename T, int Q, int PITCH, int STENCIL> __device__ __forceinline__ T fdm_bardow_stencil_pop(const T* c, T dt)
{
    const T cxq = dt * T(cxi(Q)), cyq = dt * T(cyi(Q));
    const T cxxq = T(0.5) * cxq * cxq, cyyq = T(0.5) * cyq * cyq, cxyq = cxq * cyq;
    const T fc = c[0], fe = c[PITCH], fw = c[-PITCH], fn = c[1], fs = c[-1];
    const T fne = c[PITCH + 1], fnw = c[-PITCH + 1], fse = c[PITCH - 1], fsw = c[-PITCH - 1];
    const T one_sixth = T(1) / T(6), two_thirds = T(2) / T(3), five_sixths = T(10) / T(12), one_twelth = T(1) / T(12);
    const T one_third = T(1) / T(3);
    T dfx, dfy, dfxx, dfyy, dfxy;
    if (STENCIL == 1) {
        dfx = one_sixth * ((fne - fnw) + (fe - fw) + (fse - fsw));
        dfy = one_sixth * ((fne - fse) + (fn - fs) + (fnw - fsw));
        dfxx = one_third * (fne - T(2) * fn + fnw) + one_third * (fe - T(2) * fc + fw) + one_third * (fse - T(2) * fs + fsw);
        dfyy = one_third * (fne - T(2) * fe + fse) + one_third * (fn - T(2) * fc + fs) + one_third * (fnw - T(2) * fw + fsw);
        dfxy = T(0.25) * (fne - fnw + fsw - fse);
    } else if (STENCIL == 2 || STENCIL == 3) {
        const T p1s = STENCIL == 2 ? T(0.2880584423829145035434) : T(0.3934930210807994210853);
        const T p1d = STENCIL == 2 ? T(0.1059707788085427065949) : T(0.05325348945960039354075);
        const T p2c = STENCIL == 2 ? T(-1.152233769531658458263) : T(-1.573972084323197018207);
        const T p2d1 = STENCIL == 2 ? T(0.5761168847658292291314) : T(0.7869860421615988421706);
        const T p2d2 = STENCIL == 2 ? T(-0.4238831152341712149578) : T(-0.2130139578384016019186);
        const T p2d = STENCIL == 2 ? T(0.2119415576170855242122) : T(0.1065069789192007732037);
        dfx = p1s * (fe - fw) + p1d * (fne - fnw) + p1d * (fse - fsw);
        dfy = p1s * (fn - fs) + p1d * (fne - fse) + p1d * (fnw - fsw);
        dfxx = p2c * fc + p2d1 * (fe + fw) + p2d2 * (fn + fs) + p2d * (fne + fnw + fsw + fse);
        dfyy = p2c * fc + p2d2 * (fe + fw) + p2d1 * (fn + fs) + p2d * (fne + fnw + fsw + fse);
        dfxy = T(0.25) * (fne - fnw + fsw - fse);
    } else {
        dfx = T(0.5) * (one_sixth * (fne - fnw) + two_thirds * (fe - fw) + one_sixth * (fse - fsw));
        dfy = T(0.5) * (one_sixth * (fne - fse) + two_thirds * (fn - fs) + one_sixth * (fnw - fsw));
        dfxx = one_twelth * (fne - T(2) * fn + fnw) + five_sixths * (fe - T(2) * fc + fw) + one_twelth * (fse - T(2) * fs + fsw);
        dfyy = one_twelth * (fne - T(2) * fe + fse) + five_sixths * (fn - T(2) * fc + fs) + one_twelth * (fnw - T(2) * fw + fsw);
        dfxy = T(0.25) * (fne - fse - fnw + fsw);
    }
    return fc - cxq * dfx - cyq * dfy + (cxxq * dfxx + cxyq * dfxy + cyyq * dfyy);
}

template <typename T, int STENCIL, int PITCH, int PLANE> __device__ __forceinline__ void fdm_stencil_update(const T* c0, T dt, T (&fp)[9])
{
#define PLBM_FDM_Q(Q) fp[Q] = fdm_bardow_stencil_pop<T, Q, PITCH, STENCIL>(c0 + Q * PLANE, dt)
    PLBM_FDM_Q(1);
    PLBM_FDM_Q(2);
    PLBM_FDM_Q(3);
    PLBM_FDM_Q(4);
    PLBM_FDM_Q(5);
    PLBM_FDM_Q(6);
    PLBM_FDM_Q(7);
    PLBM_FDM_Q(8);
#undef PLBM_FDM_Q
}

}  // namespace plbm
