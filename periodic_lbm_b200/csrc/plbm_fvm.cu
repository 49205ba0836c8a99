// plbm_fvm.cu -- finite-volume streaming on the D2Q9 lattice: Bardow's scheme and DUGKS (sm_100a).
//
//   fvm_bardow_kernel             src/fvm_bardow.F90:410-507
//   dugks_collide (+copy_field)   src/periodic_dugks.F90:46-77, 441-486
//   kernel_bgk                    src/periodic_dugks.F90:80-169
//   kernel_stream (+update_ew/ns) src/periodic_dugks.F90:190-438
//
// Every population reads its own 3x3 neighbourhood (9x reuse), so the step kernels stage a
// (FY+2) x (FX+2) tile of all nine populations in shared memory with a one-cell periodic halo.
// The fused DUGKS kernel additionally recomputes the half-step collision on the halo, so a whole
// step costs one read and one write of the state (144 B fp64 per node) instead of the reference's
// three passes (432 B by the author's own accounting, sim/standard_lbm.F90:331).
//
// Bit parity forbids sharing a face value between the two cells it separates: the reference
// evaluates it twice with a different order of the two subtractions (cfe of x vs cfw of x+1),
// so each node computes its own four faces exactly like the Fortran.  Terms multiplied by
// cx = 0 or cy = 0 are exact zeros and are skipped (x - 0 == x).
#include "plbm_internal.h"
#include "plbm_fv.cuh"

namespace plbm {

__device__ __forceinline__ int wrap_p1(int i, int n) { return i + 1 == n ? 0 : i + 1; }
__device__ __forceinline__ int wrap_m1(int i, int n) { return i == 0 ? n - 1 : i - 1; }
__device__ __forceinline__ int pmod(int i, int n)
{
    i %= n;
    return i < 0 ? i + n : i;
}

// ---------------------------------------------------------------------------------------
// Tile geometry of the fused kernels: FY rows (unit stride, threadIdx.x) x FX lines (threadIdx.y).
constexpr int FY = 32, FX = 8;
constexpr int GY = FY + 2, GX = FX + 2;  // with halo: 34 x 10 = 340 nodes
constexpr int PITCH = GY + 1;            // odd pitch: neighbouring lines fall in different banks
constexpr int NHALO = GX * GY - FX * FY; // 84 ring nodes

// MODE_DUGKS : fin = ftilde^n -> fout = ftilde^{n+1}   (perform_dugks_step, -DDUGKS or not)
// MODE_BARDOW: fin = f^n      -> fout = collide(stream_fvm_bardow(f^n))   (perform_step)
//   stage 1: every thread parks its own node in the shared tile (DUGKS: after the half-step
//            collision fbar^+, keeping the full-step collision ftilde^+ in registers); the first
//            84 threads also fill the periodic halo ring (DUGKS: recomputing fbar^+ there).
//   stage 2: faces from the tile, face relaxation (DUGKS), flux update, collision (Bardow), store.
// fbar^+ of the previous step (what the reference leaves in lattice `inew`, read by the lagged
// update_macros) is not stored: it is recomputed on demand from fin, bit-identically.
// MODE_FDM_STENCIL + k (k = 1..4): stream_fdm_bardow with the reference's -DFDM_WLS / _GAUSS_V1 / _GAUSS_V2 / -DFDM_ISO stencils
enum { MODE_DUGKS = 0, MODE_DUGKS_OFF = 1, MODE_BARDOW = 2, MODE_FDM_BARDOW = 4, MODE_FDM_SOFONEA = 5, MODE_FDM_STENCIL = 5 };

template <typename T, int MODE, int MODEL>
__global__ void __launch_bounds__(FY* FX, 2)
    k_fv_fused(const T* __restrict__ fin, T* __restrict__ fout, int nx, int ny, int ld, T dt, T omega_full, T omega_half,
               T omega_face, CollideParams<T> cp)
{
    extern __shared__ unsigned char smem_raw[];
    T* sm = reinterpret_cast<T*>(smem_raw);  // sm[q][sx][sy], pitch PITCH
    const int x0 = blockIdx.y * FX, y0 = blockIdx.x * FY;
    const int tx = threadIdx.y, ty = threadIdx.x;
    const int tid = tx * FY + ty;
    constexpr bool IS_DUGKS = MODE == MODE_DUGKS || MODE == MODE_DUGKS_OFF;

    // own node (tiles may overhang the grid: wrap, the result is discarded)
    const int x = x0 + tx, y = y0 + ty;
    const bool active = x < nx && y < ny;
    T fp[9];
    {
        const int xs = active ? x : pmod(x, nx), ys = active ? y : pmod(y, ny);
        T b[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) fp[q] = b[q] = fin[((size_t)q * nx + xs) * (size_t)ld + ys];
        if (IS_DUGKS) {
            collide_bgk_split(b, omega_half);
            collide_bgk_split(fp, omega_full);
        }
        T* c = sm + (tx + 1) * PITCH + (ty + 1);
#pragma unroll
        for (int q = 0; q < 9; ++q) c[q * (GX * PITCH)] = b[q];
    }
    // halo ring: 2 full lines (sx = 0, GX-1) and 2 x FX edge cells (sy = 0, GY-1)
    if (tid < NHALO) {
        int sx, sy;
        if (tid < 2 * GY) {
            sx = tid < GY ? 0 : GX - 1;
            sy = tid < GY ? tid : tid - GY;
        } else {
            const int k = tid - 2 * GY;
            sx = 1 + (k >> 1);
            sy = (k & 1) ? GY - 1 : 0;
        }
        const int xs = pmod(x0 + sx - 1, nx), ys = pmod(y0 + sy - 1, ny);
        T b[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) b[q] = fin[((size_t)q * nx + xs) * (size_t)ld + ys];
        if (IS_DUGKS) collide_bgk_split(b, omega_half);
        T* c = sm + sx * PITCH + sy;
#pragma unroll
        for (int q = 0; q < 9; ++q) c[q * (GX * PITCH)] = b[q];
    }
    __syncthreads();
    if (!active) return;

    const T* c0 = sm + (tx + 1) * PITCH + (ty + 1);
    if (MODE > MODE_FDM_STENCIL)
        fdm_stencil_update<T, MODE - MODE_FDM_STENCIL, PITCH, GX * PITCH>(c0, dt, fp);
    else if (MODE == MODE_FDM_BARDOW || MODE == MODE_FDM_SOFONEA)
        fdm_update<T, MODE == MODE_FDM_SOFONEA, PITCH, GX * PITCH>(c0, dt, fp);
    else
        flux_update<T, MODE == MODE_DUGKS, PITCH, GX * PITCH>(c0, dt, omega_face, fp);
    if (!IS_DUGKS && MODEL != M_NONE) collide<T, MODEL>(fp, cp);
#pragma unroll
    for (int q = 0; q < 9; ++q) fout[((size_t)q * nx + x) * (size_t)ld + y] = fp[q];
}

template <typename T, int MODE, int MODEL>
static int launch_fv(const Grid& g, const T* fin, T* fout, T dt, T of, T oh, T oc, const CollideParams<T>& cp, cudaStream_t s)
{
    dim3 block(FY, FX), grid((g.ny + FY - 1) / FY, (g.nx + FX - 1) / FX);
    const size_t smem = sizeof(T) * 9 * GX * PITCH;  // 25.2 KB fp64: below the 48 KB default, no opt-in needed
    k_fv_fused<T, MODE, MODEL><<<grid, block, smem, s>>>(fin, fout, g.nx, g.ny, g.ld, dt, of, oh, oc, cp);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PLBM_CUDA(cudaGetLastError());
    return PLBM_OK;
}

template <typename T>
int launch_fvm_bardow(const Grid& g, const T* fold, T* fnew, T dt, int model, const CollideParams<T>& cp, cudaStream_t s, int mode)
{
    if (mode > MODE_FDM_STENCIL) {  // alternative derivative stencils of stream_fdm_bardow: streaming only
        if (model != M_NONE) {
            set_error("fdm stencil variants: the collision is a separate launch");
            return PLBM_ERR_ARG;
        }
        switch (mode - MODE_FDM_STENCIL) {
        case 1: return launch_fv<T, MODE_FDM_STENCIL + 1, M_NONE>(g, fold, fnew, dt, T(0), T(0), T(0), cp, s);
        case 2: return launch_fv<T, MODE_FDM_STENCIL + 2, M_NONE>(g, fold, fnew, dt, T(0), T(0), T(0), cp, s);
        case 3: return launch_fv<T, MODE_FDM_STENCIL + 3, M_NONE>(g, fold, fnew, dt, T(0), T(0), T(0), cp, s);
        case 4: return launch_fv<T, MODE_FDM_STENCIL + 4, M_NONE>(g, fold, fnew, dt, T(0), T(0), T(0), cp, s);
        }
        set_error("fdm stencil variants: unknown stencil");
        return PLBM_ERR_ARG;
    }
    if (mode == MODE_FDM_BARDOW || mode == MODE_FDM_SOFONEA) {  // plain-load fallback of the FDM schemes
        const bool sof = mode == MODE_FDM_SOFONEA;
        switch (model) {
        case M_NONE: return sof ? launch_fv<T, MODE_FDM_SOFONEA, M_NONE>(g, fold, fnew, dt, T(0), T(0), T(0), cp, s)
                                : launch_fv<T, MODE_FDM_BARDOW, M_NONE>(g, fold, fnew, dt, T(0), T(0), T(0), cp, s);
        case M_BGK: return sof ? launch_fv<T, MODE_FDM_SOFONEA, M_BGK>(g, fold, fnew, dt, T(0), T(0), T(0), cp, s)
                               : launch_fv<T, MODE_FDM_BARDOW, M_BGK>(g, fold, fnew, dt, T(0), T(0), T(0), cp, s);
        case M_TRT: return sof ? launch_fv<T, MODE_FDM_SOFONEA, M_TRT>(g, fold, fnew, dt, T(0), T(0), T(0), cp, s)
                               : launch_fv<T, MODE_FDM_BARDOW, M_TRT>(g, fold, fnew, dt, T(0), T(0), T(0), cp, s);
        case M_RR: return sof ? launch_fv<T, MODE_FDM_SOFONEA, M_RR>(g, fold, fnew, dt, T(0), T(0), T(0), cp, s)
                              : launch_fv<T, MODE_FDM_BARDOW, M_RR>(g, fold, fnew, dt, T(0), T(0), T(0), cp, s);
        default: set_error("fdm streaming: only none/bgk/trt/rr are fused"); return PLBM_ERR_ARG;
        }
    }
    switch (model) {
    case M_NONE: return launch_fv<T, MODE_BARDOW, M_NONE>(g, fold, fnew, dt, T(0), T(0), T(0), cp, s);
    case M_BGK: return launch_fv<T, MODE_BARDOW, M_BGK>(g, fold, fnew, dt, T(0), T(0), T(0), cp, s);
    case M_TRT: return launch_fv<T, MODE_BARDOW, M_TRT>(g, fold, fnew, dt, T(0), T(0), T(0), cp, s);
    case M_RR: return launch_fv<T, MODE_BARDOW, M_RR>(g, fold, fnew, dt, T(0), T(0), T(0), cp, s);
    case M_BGK_SPLIT: return launch_fv<T, MODE_BARDOW, M_BGK_SPLIT>(g, fold, fnew, dt, T(0), T(0), T(0), cp, s);
    case M_TRT_SPLIT: return launch_fv<T, MODE_BARDOW, M_TRT_SPLIT>(g, fold, fnew, dt, T(0), T(0), T(0), cp, s);
    case M_BGK_IMPROVED: return launch_fv<T, MODE_BARDOW, M_BGK_IMPROVED>(g, fold, fnew, dt, T(0), T(0), T(0), cp, s);
    }
    set_error("fvm_bardow: unknown collision model");
    return PLBM_ERR_ARG;
}

template <typename T>
int launch_dugks_fused(const Grid& g, const T* fin, T* fout, T dt, T omega_full, T omega_half, T omega_face, bool dugks,
                       cudaStream_t s)
{
    const CollideParams<T> cp{T(0), T(0)};
    if (dugks) return launch_fv<T, MODE_DUGKS, M_NONE>(g, fin, fout, dt, omega_full, omega_half, omega_face, cp, s);
    return launch_fv<T, MODE_DUGKS_OFF, M_NONE>(g, fin, fout, dt, omega_full, omega_half, omega_face, cp, s);
}

// ---------------------------------------------------------------------------------------
// The reference's separately public DUGKS procedures (two passes over global memory).
// dugks_collide: fnew = BGK(fold, omega_full) ; fold = BGK(fold, omega_half)
template <typename T>
__global__ void __launch_bounds__(256) k_dugks_collide(T* __restrict__ fold, T* __restrict__ fnew, int nx, int ny, int ld,
                                                       T omega_full, T omega_half)
{
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int x = (int)(g / (size_t)ld);
    const int y = (int)(g - (size_t)x * ld);
    if (x >= nx || y >= ny) return;
    T a[9], b[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) a[q] = b[q] = fold[((size_t)q * nx + x) * (size_t)ld + y];
    collide_bgk_split(a, omega_full);
    collide_bgk_split(b, omega_half);
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        fnew[((size_t)q * nx + x) * (size_t)ld + y] = a[q];
        fold[((size_t)q * nx + x) * (size_t)ld + y] = b[q];
    }
}

template <typename T> int launch_dugks_collide(const Grid& g, T* fold, T* fnew, T omega_full, T omega_half, cudaStream_t s)
{
    const size_t n = (size_t)g.nx * g.ld;
    k_dugks_collide<T><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(fold, fnew, g.nx, g.ny, g.ld, omega_full, omega_half);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PLBM_CUDA(cudaGetLastError());
    return PLBM_OK;
}

// dugks_stream: ft = fbar^+ (global memory), fp updated in place for q = 1..8.
template <typename T, bool DUGKS>
__global__ void __launch_bounds__(256) k_dugks_stream(const T* __restrict__ ft, T* __restrict__ fp_, int nx, int ny, int ld,
                                                      T dt, T omega_face)
{
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int x = (int)(g / (size_t)ld);
    const int y = (int)(g - (size_t)x * ld);
    if (x >= nx || y >= ny) return;
    const int xs[3] = {wrap_m1(x, nx), x, wrap_p1(x, nx)};
    const int ys[3] = {wrap_m1(y, ny), y, wrap_p1(y, ny)};
    T cfw[9], cfn[9], cfe[9], cfs[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const T cxq = dt * T(cxi(q)), cyq = dt * T(cyi(q));
        auto nb = [&](int dx, int dy) -> T { return ft[((size_t)q * nx + xs[dx + 1]) * (size_t)ld + ys[dy + 1]]; };
        faces(nb(0, 0), nb(1, 0), nb(0, 1), nb(-1, 0), nb(0, -1), nb(1, 1), nb(-1, 1), nb(-1, -1), nb(1, -1), cxq, cyq, cfw[q],
              cfn[q], cfe[q], cfs[q]);
    }
    if (DUGKS) {
        face_relax<T, true>(cfw, omega_face);
        face_relax<T, true>(cfe, omega_face);
        face_relax<T, false>(cfn, omega_face);
        face_relax<T, false>(cfs, omega_face);
    }
#pragma unroll
    for (int q = 1; q < 9; ++q) {
        const T cxq = dt * T(cxi(q)), cyq = dt * T(cyi(q));
        const size_t i = ((size_t)q * nx + x) * (size_t)ld + y;
        fp_[i] = fp_[i] - cxq * (cfe[q] - cfw[q]) - cyq * (cfn[q] - cfs[q]);
    }
}

template <typename T>
int launch_dugks_stream(const Grid& g, const T* ft, T* fp, T dt, T omega_face, bool dugks, cudaStream_t s)
{
    const size_t n = (size_t)g.nx * g.ld;
    const unsigned nb = (unsigned)((n + 255) / 256);
    if (dugks)
        k_dugks_stream<T, true><<<nb, 256, 0, s>>>(ft, fp, g.nx, g.ny, g.ld, dt, omega_face);
    else
        k_dugks_stream<T, false><<<nb, 256, 0, s>>>(ft, fp, g.nx, g.ny, g.ld, dt, omega_face);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PLBM_CUDA(cudaGetLastError());
    return PLBM_OK;
}

#define INST(T)                                                                                                    \
    template int launch_fvm_bardow<T>(const Grid&, const T*, T*, T, int, const CollideParams<T>&, cudaStream_t, int); \
    template int launch_dugks_collide<T>(const Grid&, T*, T*, T, T, cudaStream_t);                                 \
    template int launch_dugks_stream<T>(const Grid&, const T*, T*, T, T, bool, cudaStream_t);                      \
    template int launch_dugks_fused<T>(const Grid&, const T*, T*, T, T, T, T, bool, cudaStream_t);
INST(double)
INST(float)
#undef INST

}  // namespace plbm
