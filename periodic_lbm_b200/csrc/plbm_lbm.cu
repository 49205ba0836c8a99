// plbm_lbm.cu -- fused pull-scheme stream + collide for periodic D2Q9 (sm_100a).
//
// One kernel replaces the reference's two sweeps per step
//   lbm_stream_kernel   src/periodic_lbm.f90:45-127   fdst(y,x,q) = fsrc(y-cy_q, x-cx_q, q)
//   <collision kernel>  src/collision_*.F90           in place on fdst
// Layout is the reference's f(ld,nx,0:8): y is the unit-stride index, so threads map to y,
// stores are always 16-byte aligned vectors, populations with cy = 0 (q = 0,1,3) load
// aligned vectors from line x -/+ 1, and the six populations with cy = +-1 need the same
// vector shifted by one element (done either by element loads or by a warp shuffle of the
// aligned vector).  Periodic wrap is index arithmetic: a per-line select in x, a select on
// the first/last row in y (to ny-1 / 0, not into the padding rows ny..ld-1).
//
// Algorithmic traffic: 9 reads + 9 writes per node = 144 B (fp64) / 72 B (fp32).
#include "plbm_internal.h"

namespace plbm {

template <typename T, int V> struct alignas(sizeof(T) * V) Pack {
    T v[V];
};

// 128-bit accesses with a streaming (evict-first) hint: LM = 2 measurement variant
template <typename T, int V> __device__ __forceinline__ Pack<T, V> ld_stream(const T* p)
{
    Pack<T, V> r;
    if constexpr (V == 2 && sizeof(T) == 8) {
        double2 v = __ldcs(reinterpret_cast<const double2*>(p));
        r.v[0] = v.x;
        r.v[1] = v.y;
    } else if constexpr (V == 4 && sizeof(T) == 4) {
        float4 v = __ldcs(reinterpret_cast<const float4*>(p));
        r.v[0] = v.x;
        r.v[1] = v.y;
        r.v[2] = v.z;
        r.v[3] = v.w;
    } else {
        r = *reinterpret_cast<const Pack<T, V>*>(p);
    }
    return r;
}
template <typename T, int V> __device__ __forceinline__ void st_stream(T* p, const Pack<T, V>& r)
{
    if constexpr (V == 2 && sizeof(T) == 8) {
        __stcs(reinterpret_cast<double2*>(p), make_double2(r.v[0], r.v[1]));
    } else if constexpr (V == 4 && sizeof(T) == 4) {
        __stcs(reinterpret_cast<float4*>(p), make_float4(r.v[0], r.v[1], r.v[2], r.v[3]));
    } else {
        *reinterpret_cast<Pack<T, V>*>(p) = r;
    }
}

// Halo layout (shared with the two-step kernel, plbm_lbm2.cu): all nine populations of the neighbours' two
// nearest lines, [2][9][ld]; halo_lo holds lines -2, -1 and halo_hi lines nx, nx+1.

// Pointer to row 0 of the source line of population q for destination line x.
template <typename T, int Q> __device__ __forceinline__ const T* src_line(const LbmArgs<T>& a, int x)
{
    constexpr int cx = cxi(Q);
    int xs = x - cx;
    if (cx == 1 && xs < 0) {
        if (a.halo_lo) return a.halo_lo + (size_t)(9 + Q) * a.ld;  // line -1
        xs = a.nx - 1;
    }
    if (cx == -1 && xs >= a.nx) {
        if (a.halo_hi) return a.halo_hi + (size_t)Q * a.ld;  // line nx
        xs = 0;
    }
    return a.src + ((size_t)Q * a.nx + xs) * (size_t)a.ld;
}

// LM = 0: element loads for the shifted populations.  LM = 1: aligned vector + shuffle.
template <typename T, int V, int LM, int Q>
__device__ __forceinline__ void load_pop(const LbmArgs<T>& a, int x, int y0, bool active, T (&f)[V][9])
{
    constexpr int cy = cyi(Q);
    const T* line = src_line<T, Q>(a, x);
    if (cy == 0) {
        Pack<T, V> p = LM == 2 ? ld_stream<T, V>(line + y0) : *reinterpret_cast<const Pack<T, V>*>(line + y0);
#pragma unroll
        for (int v = 0; v < V; ++v) f[v][Q] = p.v[v];
    } else if (LM == 0 || LM == 2 || V == 1) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
            int ys = y0 + v - cy;
            if (cy == 1 && ys < 0) ys = a.ny - 1;
            if (cy == -1 && ys >= a.ny) ys = 0;
            f[v][Q] = line[ys];
        }
    } else {
        Pack<T, V> p = *reinterpret_cast<const Pack<T, V>*>(line + y0);
        const int lane = threadIdx.x & 31;
        if (cy == 1) {  // need rows y0-1 .. y0+V-2
            T prev = __shfl_up_sync(0xffffffffu, p.v[V - 1], 1);
            if (lane == 0 || y0 == 0) prev = line[y0 == 0 ? a.ny - 1 : y0 - 1];
            f[0][Q] = prev;
#pragma unroll
            for (int v = 1; v < V; ++v) f[v][Q] = p.v[v - 1];
        } else {  // need rows y0+1 .. y0+V
            T next = __shfl_down_sync(0xffffffffu, p.v[0], 1);
            if (lane == 31 || y0 + V >= a.ny) next = line[y0 + V >= a.ny ? 0 : y0 + V];
#pragma unroll
            for (int v = 0; v < V - 1; ++v) f[v][Q] = p.v[v + 1];
            f[V - 1][Q] = next;
        }
    }
    (void)active;
}

// STREAM = true : pull from src neighbours (fused stream+collide, or stream only if MODEL = M_NONE)
// STREAM = false: in-place collision on dst (the reference's separate collide_* entry points)
// PRE = true additionally stores the streamed, pre-collision PDFs to a.pre (perform_triple_step); it is a
// separate instantiation because the extra store block costs the hot kernels registers (RR fp64: 84 vs 80
// registers = 2 instead of 3 resident blocks per SM, measured -24 %).
template <typename T, int MODEL, bool STREAM, int V, int LM, bool PRE>
__global__ void __launch_bounds__(256, 3) k_lbm(const LbmArgs<T> a)
{
    const int ldv = a.ld / V;
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int xrel = (int)(g / (size_t)ldv);
    int x = a.x_begin + xrel;
    if (x >= a.x_split) x += a.x_skip;  // second range of a split launch (both boundaries of a slab)
    int y0 = ((int)(g - (size_t)xrel * ldv)) * V;
    const bool active = (x < a.x_end) && (y0 < a.ny);
    if (!active) {  // keep the lane alive for the shuffles, on a harmless address
        if (LM == 0 || LM == 2 || V == 1) return;
        x = a.x_begin;
        y0 = 0;
    }

    T f[V][9];
    if (STREAM) {
        load_pop<T, V, LM, 0>(a, x, y0, active, f);
        load_pop<T, V, LM, 1>(a, x, y0, active, f);
        load_pop<T, V, LM, 2>(a, x, y0, active, f);
        load_pop<T, V, LM, 3>(a, x, y0, active, f);
        load_pop<T, V, LM, 4>(a, x, y0, active, f);
        load_pop<T, V, LM, 5>(a, x, y0, active, f);
        load_pop<T, V, LM, 6>(a, x, y0, active, f);
        load_pop<T, V, LM, 7>(a, x, y0, active, f);
        load_pop<T, V, LM, 8>(a, x, y0, active, f);
    } else {
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            Pack<T, V> p = *reinterpret_cast<const Pack<T, V>*>(a.dst + ((size_t)q * a.nx + x) * (size_t)a.ld + y0);
#pragma unroll
            for (int v = 0; v < V; ++v) f[v][q] = p.v[v];
        }
    }
    if (!active) return;

    if (STREAM && PRE) {  // perform_triple_step keeps the pre-collision lattice
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            Pack<T, V> p;
#pragma unroll
            for (int v = 0; v < V; ++v) p.v[v] = f[v][q];
            *reinterpret_cast<Pack<T, V>*>(a.pre + ((size_t)q * a.nx + x) * (size_t)a.ld + y0) = p;
        }
    }

    if (MODEL != M_NONE) {
#pragma unroll
        for (int v = 0; v < V; ++v) collide<T, MODEL>(f[v], a.cp);
    }

#pragma unroll
    for (int q = 0; q < 9; ++q) {
        Pack<T, V> p;
#pragma unroll
        for (int v = 0; v < V; ++v) p.v[v] = f[v][q];
        if (LM == 2)
            st_stream<T, V>(a.dst + ((size_t)q * a.nx + x) * (size_t)a.ld + y0, p);
        else
            *reinterpret_cast<Pack<T, V>*>(a.dst + ((size_t)q * a.nx + x) * (size_t)a.ld + y0) = p;
    }
}

template <typename T, int MODEL, bool STREAM, int V, int LM> static int launch_one(const LbmArgs<T>& a, cudaStream_t s, bool pre = false)
{
    const size_t nthreads = (size_t)(a.x_end - a.x_begin - a.x_skip) * (size_t)(a.ld / V);
    if (nthreads == 0) return PLBM_OK;
    const int block = 256;
    const size_t nblocks = (nthreads + block - 1) / block;
    if (pre)
        k_lbm<T, MODEL, STREAM, V, 0, STREAM><<<(unsigned)nblocks, block, 0, s>>>(a);  // element-load variant only
    else
        k_lbm<T, MODEL, STREAM, V, LM, false><<<(unsigned)nblocks, block, 0, s>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PLBM_CUDA(cudaGetLastError());
    return PLBM_OK;
}

template <typename T, int MODEL, bool STREAM> static int launch_v(const LbmArgs<T>& a, int variant, cudaStream_t s)
{
    constexpr int VMAX = 16 / (int)sizeof(T);  // 128-bit vectors: 2 x fp64, 4 x fp32
    const bool vec_ok = (a.ny % VMAX) == 0;
    // variant 0 (default) = 128-bit vectors, element loads for the six y-shifted populations
    // (measured fastest on B200: ~1.05x the measured copy bandwidth, profiles/);
    // variant 1 = 128-bit vectors + warp shuffle of the aligned vector; 2 = scalar (1 node/thread)
    if (a.pre != nullptr) return vec_ok ? launch_one<T, MODEL, STREAM, VMAX, 0>(a, s, true) : launch_one<T, MODEL, STREAM, 1, 0>(a, s, true);
    if (!vec_ok || variant == 2) return launch_one<T, MODEL, STREAM, 1, 0>(a, s);
    if (variant == 1) return launch_one<T, MODEL, STREAM, VMAX, 1>(a, s);
    if (variant == 4) return launch_one<T, MODEL, STREAM, VMAX, 2>(a, s);  // streaming cache hints (measurement)
    return launch_one<T, MODEL, STREAM, VMAX, 0>(a, s);
}

template <typename T> int launch_lbm(const LbmArgs<T>& a, int model, bool stream_pdfs, int variant, cudaStream_t s)
{
    if (stream_pdfs) {
        switch (model) {
        case M_NONE: return launch_v<T, M_NONE, true>(a, variant, s);
        case M_BGK: return launch_v<T, M_BGK, true>(a, variant, s);
        case M_TRT: return launch_v<T, M_TRT, true>(a, variant, s);
        case M_RR: return launch_v<T, M_RR, true>(a, variant, s);
        case M_BGK_SPLIT: return launch_v<T, M_BGK_SPLIT, true>(a, variant, s);
        case M_TRT_SPLIT: return launch_v<T, M_TRT_SPLIT, true>(a, variant, s);
        case M_BGK_IMPROVED: return launch_v<T, M_BGK_IMPROVED, true>(a, variant, s);
        }
    } else {
        switch (model) {
        case M_BGK: return launch_v<T, M_BGK, false>(a, variant, s);
        case M_TRT: return launch_v<T, M_TRT, false>(a, variant, s);
        case M_RR: return launch_v<T, M_RR, false>(a, variant, s);
        case M_BGK_SPLIT: return launch_v<T, M_BGK_SPLIT, false>(a, variant, s);
        case M_TRT_SPLIT: return launch_v<T, M_TRT_SPLIT, false>(a, variant, s);
        case M_BGK_IMPROVED: return launch_v<T, M_BGK_IMPROVED, false>(a, variant, s);
        }
    }
    set_error("launch_lbm: unknown collision model");
    return PLBM_ERR_ARG;
}

template int launch_lbm<double>(const LbmArgs<double>&, int, bool, int, cudaStream_t);
template int launch_lbm<float>(const LbmArgs<float>&, int, bool, int, cudaStream_t);

// Pack the PLBM_HALO_LINES boundary lines of each side into contiguous send buffers (multi-GPU ring, NCCL transport):
// send_lo <- lines 0, 1, 2 (become the low neighbour's halo_hi), send_hi <- lines nx-2, nx-1, nx-3 (the high neighbour's
// halo_lo, in the order of plbm_internal.h), all nine populations.  Slabs thinner than three lines repeat a line (unused).
template <typename T> __global__ void k_halo_pack(const T* f, T* send_lo, T* send_hi, int nx, int ld)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= PLBM_HALO_LINES * 9 * ld) return;
    const int lq = i / ld, y = i - lq * ld;
    const int l = lq / 9, q = lq - 9 * l;
    const int line_lo = min(l, nx - 1), line_hi = max(halo_lo_source_line(l, nx), 0);
    send_lo[i] = f[((size_t)q * nx + line_lo) * (size_t)ld + y];
    send_hi[i] = f[((size_t)q * nx + line_hi) * (size_t)ld + y];
}

template <typename T> int launch_halo_pack(const Grid& g, const T* f, T* send_lo, T* send_hi, cudaStream_t s)
{
    const int n = PLBM_HALO_LINES * 9 * g.ld;
    k_halo_pack<T><<<(n + 255) / 256, 256, 0, s>>>(f, send_lo, send_hi, g.nx, g.ld);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PLBM_CUDA(cudaGetLastError());
    return PLBM_OK;
}
// all nine populations of line 0 (-> send_lo) and line nx-1 (-> send_hi): FVM / DUGKS halo
template <typename T> __global__ void k_halo_pack9(const T* f, T* send_lo, T* send_hi, int nx, int ld)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 9 * ld) return;
    const int q = i / ld, y = i - q * ld;
    send_lo[i] = f[((size_t)q * nx + 0) * (size_t)ld + y];
    send_hi[i] = f[((size_t)q * nx + (nx - 1)) * (size_t)ld + y];
}
template <typename T> int launch_halo_pack9(const Grid& g, const T* f, T* send_lo, T* send_hi, cudaStream_t s)
{
    const int n = 9 * g.ld;
    k_halo_pack9<T><<<(n + 255) / 256, 256, 0, s>>>(f, send_lo, send_hi, g.nx, g.ld);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PLBM_CUDA(cudaGetLastError());
    return PLBM_OK;
}
template int launch_halo_pack9<double>(const Grid&, const double*, double*, double*, cudaStream_t);
template int launch_halo_pack9<float>(const Grid&, const float*, float*, float*, cudaStream_t);
template int launch_halo_pack<double>(const Grid&, const double*, double*, double*, cudaStream_t);
template int launch_halo_pack<float>(const Grid&, const float*, float*, float*, cudaStream_t);

}  // namespace plbm
