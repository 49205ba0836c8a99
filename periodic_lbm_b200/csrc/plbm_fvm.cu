// plbm_fvm.cu -- finite-volume streaming on the D2Q9 lattice: Bardow's scheme and DUGKS (sm_100a).
//
//   fvm_bardow_kernel            src/fvm_bardow.F90:410-507
//   dugks_collide (+copy_field)  src/periodic_dugks.F90:46-77, 441-486
//   kernel_bgk                   src/periodic_dugks.F90:80-169
//   kernel_stream (+update_ew/ns) src/periodic_dugks.F90:190-438
//
// Every population reads its own 3x3 neighbourhood (9x reuse), so these kernels stage a
// (TY+2) x (TX+2) tile of each population in shared memory with a one-cell periodic halo;
// the fused DUGKS kernel additionally recomputes the half-step collision on the halo so a
// whole step costs one read and one write of the state (144 B fp64 per node) instead of the
// reference's three passes.
#include "plbm_internal.h"

namespace plbm {

__device__ __forceinline__ int wrap_p1(int i, int n) { return i + 1 == n ? 0 : i + 1; }
__device__ __forceinline__ int wrap_m1(int i, int n) { return i == 0 ? n - 1 : i - 1; }
__device__ __forceinline__ int pmod(int i, int n) { i %= n; return i < 0 ? i + n : i; }

// ---------------------------------------------------------------------------------------
// Tile geometry: TY rows (unit stride, threadIdx.x) x TX lines (threadIdx.y).
constexpr int TY = 64;
constexpr int TX = 4;
constexpr int SY = TY + 2;  // tile + halo
constexpr int SX = TX + 2;

// Load the (SY x SX) halo tile of population q of `f` into sm[sx][sy]; periodic in x and y.
// With a slab decomposition the x-neighbours beyond the slab are not available here: the
// FVM/DUGKS kernels are single-GPU (nx == nx_global).
template <typename T, typename F>
__device__ __forceinline__ void for_tile(int x0, int y0, int nx, int ny, F&& fn)
{
    for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < SX * SY; i += blockDim.x * blockDim.y) {
        const int sx = i / SY, sy = i - sx * SY;
        // tiles may overhang the grid when nx % TX or ny % TY != 0: wrap whatever is outside
        const int x = pmod(x0 + sx - 1, nx), y = pmod(y0 + sy - 1, ny);
        fn(sx, sy, x, y);
    }
}

// ---------------------------------------------------------------------------------------
// stream_fvm_bardow fused with the collision that perform_step applies right after it.
template <typename T, int MODEL>
__global__ void __launch_bounds__(TY* TX) k_fvm_bardow(const T* __restrict__ fold, T* __restrict__ fnew, int nx, int ny,
                                                        int ld, T dt, CollideParams<T> cp)
{
    __shared__ T sm[SX][SY + 1];
    const int x0 = blockIdx.y * TX, y0 = blockIdx.x * TY;
    const int tx = threadIdx.y, ty = threadIdx.x;
    const int x = x0 + tx, y = y0 + ty;
    const bool active = x < nx && y < ny;
    T f[9];

    if (active) f[0] = fold[(size_t)x * ld + y];  // rest population: plain copy (:429)
#pragma unroll
    for (int q = 1; q < 9; ++q) {
        const T* fq = fold + (size_t)q * nx * (size_t)ld;
        __syncthreads();
        for_tile<T>(x0, y0, nx, ny, [&](int sx, int sy, int gx, int gy) { sm[sx][sy] = fq[(size_t)gx * ld + gy]; });
        __syncthreads();
        const T cxq = dt * T(cxi(q)), cyq = dt * T(cyi(q));
        const int cxs = tx + 1, cys = ty + 1;
        const T fc = sm[cxs][cys], fe = sm[cxs + 1][cys], fw = sm[cxs - 1][cys];
        const T fn = sm[cxs][cys + 1], fs = sm[cxs][cys - 1];
        const T fne = sm[cxs + 1][cys + 1], fnw = sm[cxs - 1][cys + 1];
        const T fsw = sm[cxs - 1][cys - 1], fse = sm[cxs + 1][cys - 1];
        T cfw, cfn, cfe, cfs;
        faces(fc, fe, fn, fw, fs, fne, fnw, fsw, fse, cxq, cyq, cfw, cfn, cfe, cfs);
        f[q] = fc - cxq * (cfe - cfw) - cyq * (cfn - cfs);
    }
    if (!active) return;
    if (MODEL != M_NONE) collide<T, MODEL>(f, cp);
#pragma unroll
    for (int q = 0; q < 9; ++q) fnew[((size_t)q * nx + x) * (size_t)ld + y] = f[q];
}

template <typename T>
int launch_fvm_bardow(const Grid& g, const T* fold, T* fnew, T dt, int model, const CollideParams<T>& cp, cudaStream_t s)
{
    dim3 block(TY, TX), grid((g.ny + TY - 1) / TY, (g.nx + TX - 1) / TX);
    switch (model) {
    case M_NONE: k_fvm_bardow<T, M_NONE><<<grid, block, 0, s>>>(fold, fnew, g.nx, g.ny, g.ld, dt, cp); break;
    case M_BGK: k_fvm_bardow<T, M_BGK><<<grid, block, 0, s>>>(fold, fnew, g.nx, g.ny, g.ld, dt, cp); break;
    case M_TRT: k_fvm_bardow<T, M_TRT><<<grid, block, 0, s>>>(fold, fnew, g.nx, g.ny, g.ld, dt, cp); break;
    case M_RR: k_fvm_bardow<T, M_RR><<<grid, block, 0, s>>>(fold, fnew, g.nx, g.ny, g.ld, dt, cp); break;
    case M_BGK_SPLIT: k_fvm_bardow<T, M_BGK_SPLIT><<<grid, block, 0, s>>>(fold, fnew, g.nx, g.ny, g.ld, dt, cp); break;
    default: set_error("fvm_bardow: unknown collision model"); return PLBM_ERR_ARG;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PLBM_CUDA(cudaGetLastError());
    return PLBM_OK;
}

// ---------------------------------------------------------------------------------------
// dugks_collide: fnew = BGK(fold, omega_full) ; fold = BGK(fold, omega_half)   (one pass)
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

// ---------------------------------------------------------------------------------------
// Shared body of kernel_stream for one node: given a functor nb(q, dx, dy) returning
// fbar^+ of population q at the (dx,dy) neighbour, compute the four faces of all nine
// populations, relax them (DUGKS) and return the flux update of fp[1..8].
template <typename T, bool DUGKS, typename NB>
__device__ __forceinline__ void dugks_node(NB&& nb, T dt, T omega_face, T (&fp)[9])
{
    T cfw[9], cfn[9], cfe[9], cfs[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const T cxq = dt * T(cxi(q)), cyq = dt * T(cyi(q));
        const T fc = nb(q, 0, 0), fe = nb(q, 1, 0), fn = nb(q, 0, 1), fw = nb(q, -1, 0), fs = nb(q, 0, -1);
        const T fne = nb(q, 1, 1), fnw = nb(q, -1, 1), fsw = nb(q, -1, -1), fse = nb(q, 1, -1);
        faces(fc, fe, fn, fw, fs, fne, fnw, fsw, fse, cxq, cyq, cfw[q], cfn[q], cfe[q], cfs[q]);
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
        fp[q] = fp[q] - cxq * (cfe[q] - cfw[q]) - cyq * (cfn[q] - cfs[q]);
    }
}

// dugks_stream as a separate entry (ft = fbar^+ in global memory, fp updated in place).
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
    T fp[9];
#pragma unroll
    for (int q = 1; q < 9; ++q) fp[q] = fp_[((size_t)q * nx + x) * (size_t)ld + y];
    auto nb = [&](int q, int dx, int dy) -> T { return ft[((size_t)q * nx + xs[dx + 1]) * (size_t)ld + ys[dy + 1]]; };
    dugks_node<T, DUGKS>(nb, dt, omega_face, fp);
#pragma unroll
    for (int q = 1; q < 9; ++q) fp_[((size_t)q * nx + x) * (size_t)ld + y] = fp[q];
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

// ---------------------------------------------------------------------------------------
// Fused DUGKS step: fin = ftilde^n  ->  fout = ftilde^{n+1}.
//   stage 1: every thread of the (SY x SX) halo tile reads the nine populations of one node,
//            applies the half-step collision (fbar^+) and parks it in shared memory; the
//            interior threads also keep the full-step collision (ftilde^+) in registers.
//   stage 2: interior threads reconstruct faces from the shared tile, relax, update, store.
// fbar^+ of the previous step (what the reference leaves in lattice `inew`, read by the
// lagged update_macros) is not stored: it is recomputed on demand from fin, bit-identically.
constexpr int FY = 32, FX = 8;         // interior tile of the fused kernel
constexpr int GY = FY + 2, GX = FX + 2;  // with halo: 34 x 10 = 340 nodes

template <typename T, bool DUGKS>
__global__ void __launch_bounds__(FY* FX) k_dugks_fused(const T* __restrict__ fin, T* __restrict__ fout, int nx, int ny,
                                                         int ld, T dt, T omega_full, T omega_half, T omega_face)
{
    extern __shared__ unsigned char smem_raw[];
    T(*sm)[GX][GY + 1] = reinterpret_cast<T(*)[GX][GY + 1]>(smem_raw);  // sm[q][sx][sy]
    const int x0 = blockIdx.y * FX, y0 = blockIdx.x * FY;
    const int tid = threadIdx.y * FY + threadIdx.x;

    for (int i = tid; i < GX * GY; i += FX * FY) {
        const int sx = i / GY, sy = i - sx * GY;
        const int x = pmod(x0 + sx - 1, nx), y = pmod(y0 + sy - 1, ny);
        T b[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) b[q] = fin[((size_t)q * nx + x) * (size_t)ld + y];
        collide_bgk_split(b, omega_half);
#pragma unroll
        for (int q = 0; q < 9; ++q) sm[q][sx][sy] = b[q];
    }
    __syncthreads();

    const int x = x0 + threadIdx.y, y = y0 + threadIdx.x;
    if (x >= nx || y >= ny) return;
    T fp[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) fp[q] = fin[((size_t)q * nx + x) * (size_t)ld + y];
    collide_bgk_split(fp, omega_full);
    const int cxs = threadIdx.y + 1, cys = threadIdx.x + 1;
    auto nb = [&](int q, int dx, int dy) -> T { return sm[q][cxs + dx][cys + dy]; };
    dugks_node<T, DUGKS>(nb, dt, omega_face, fp);
#pragma unroll
    for (int q = 0; q < 9; ++q) fout[((size_t)q * nx + x) * (size_t)ld + y] = fp[q];
}

template <typename T>
int launch_dugks_fused(const Grid& g, const T* fin, T* fout, T dt, T omega_full, T omega_half, T omega_face, bool dugks,
                       cudaStream_t s)
{
    dim3 block(FY, FX), grid((g.ny + FY - 1) / FY, (g.nx + FX - 1) / FX);
    const size_t smem = sizeof(T) * 9 * GX * (GY + 1);
    if (dugks) {
        PLBM_CUDA(cudaFuncSetAttribute(k_dugks_fused<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_dugks_fused<T, true><<<grid, block, smem, s>>>(fin, fout, g.nx, g.ny, g.ld, dt, omega_full, omega_half, omega_face);
    } else {
        PLBM_CUDA(cudaFuncSetAttribute(k_dugks_fused<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_dugks_fused<T, false><<<grid, block, smem, s>>>(fin, fout, g.nx, g.ny, g.ld, dt, omega_full, omega_half, omega_face);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PLBM_CUDA(cudaGetLastError());
    return PLBM_OK;
}

#define INST(T)                                                                                                    \
    template int launch_fvm_bardow<T>(const Grid&, const T*, T*, T, int, const CollideParams<T>&, cudaStream_t);    \
    template int launch_dugks_collide<T>(const Grid&, T*, T*, T, T, cudaStream_t);                                 \
    template int launch_dugks_stream<T>(const Grid&, const T*, T*, T, T, bool, cudaStream_t);                      \
    template int launch_dugks_fused<T>(const Grid&, const T*, T*, T, T, T, T, bool, cudaStream_t);
INST(double)
INST(float)
#undef INST

}  // namespace plbm
