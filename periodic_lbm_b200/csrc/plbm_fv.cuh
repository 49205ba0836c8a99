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

}  // namespace plbm
