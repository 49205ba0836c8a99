// plbm_lbm2.cu -- TWO fused stream+collide steps per pass over HBM (temporal blocking, sm_100a).
//
// k_lbm (plbm_lbm.cu) already moves the algorithmic minimum of one step, 9 reads + 9 writes per
// node, at the measured HBM copy rate, so the only way to go faster is to touch HBM less often.
// Here a thread block owns a strip of rows and marches along x: per column it
//   A. pulls the nine populations of column x+1 from lattice `src` (global, periodic wrap),
//      collides them (state after step 1) and parks them in a shared-memory ring,
//   B. pulls the populations of column x from the ring (columns x-1, x, x+1 of step 1),
//      collides again (state after step 2) and stores them to lattice `dst`.
// The intermediate lattice never exists in HBM: 9 reads + 9 writes per node per TWO steps.
// Redundant work: V halo rows above and below the strip in phase A, two warm-up columns per
// x segment.  The ring keeps a population only as long as phase B needs it (cx=-1: consumed in
// the iteration it is produced, depth 2; cx=0: depth 3; cx=+1: depth 4; one __syncthreads per
// column), 27 column slots in all.  Arithmetic per node is collide<T,MODEL> on the same operands
// as two k_lbm launches, so the result is bit-identical.
//   lbm_stream_kernel  src/periodic_lbm.f90:45-127 ;  collisions src/collision_*.F90
#include "plbm_internal.h"

namespace plbm {

namespace {

template <typename T, int V> struct alignas(sizeof(T) * V) Vec {
    T v[V];
};

__host__ __device__ constexpr int ring_depth(int q) { return cxi(q) == -1 ? 2 : (cxi(q) == 0 ? 3 : 4); }
// first ring slot of population q: (3,6,7) depth 2 | (0,2,4) depth 3 | (1,5,8) depth 4
__host__ __device__ constexpr int ring_base(int q)
{
    return q == 3 ? 0 : q == 6 ? 2 : q == 7 ? 4 : q == 0 ? 6 : q == 2 ? 9 : q == 4 ? 12 : q == 1 ? 15 : q == 5 ? 19 : 23;
}
constexpr int RING_SLOTS = 27;

template <typename T> struct Lbm2Args {
    const T* src;
    T* dst;
    int nx, ny, ld;
    int ty;       // interior rows per strip (multiple of V)
    int nstrips;  // strips along y
    int seglen;   // columns per x segment
    CollideParams<T> cp;
};

// streamed (pre-collision) populations of logical column xl, rows yp..yp+V-1
template <typename T, int V, int Q>
__device__ __forceinline__ void pull_global(const Lbm2Args<T>& a, int xm, int xc, int xp, int yp, T (&f)[V][9])
{
    constexpr int cx = cxi(Q), cy = cyi(Q);
    const int xs = cx == 1 ? xm : (cx == 0 ? xc : xp);
    const T* line = a.src + ((size_t)Q * a.nx + xs) * (size_t)a.ld;
    if (cy == 0) {
        const Vec<T, V> p = *reinterpret_cast<const Vec<T, V>*>(line + yp);
#pragma unroll
        for (int v = 0; v < V; ++v) f[v][Q] = p.v[v];
    } else {
#pragma unroll
        for (int v = 0; v < V; ++v) {
            int ys = yp + v - cy;
            if (cy == 1 && ys < 0) ys = a.ny - 1;
            if (cy == -1 && ys >= a.ny) ys = 0;
            f[v][Q] = line[ys];
        }
    }
}

template <typename T, int V> __device__ __forceinline__ void load_column(const Lbm2Args<T>& a, int xl, int yp, T (&f)[V][9])
{
    // xl in [-1, nx]: wrap the three source columns once
    int xc = xl < 0 ? xl + a.nx : (xl >= a.nx ? xl - a.nx : xl);
    int xm = xc == 0 ? a.nx - 1 : xc - 1;
    int xp = xc + 1 == a.nx ? 0 : xc + 1;
    pull_global<T, V, 0>(a, xm, xc, xp, yp, f);
    pull_global<T, V, 1>(a, xm, xc, xp, yp, f);
    pull_global<T, V, 2>(a, xm, xc, xp, yp, f);
    pull_global<T, V, 3>(a, xm, xc, xp, yp, f);
    pull_global<T, V, 4>(a, xm, xc, xp, yp, f);
    pull_global<T, V, 5>(a, xm, xc, xp, yp, f);
    pull_global<T, V, 6>(a, xm, xc, xp, yp, f);
    pull_global<T, V, 7>(a, xm, xc, xp, yp, f);
    pull_global<T, V, 8>(a, xm, xc, xp, yp, f);
}

// ring slot of the column written in this iteration, one counter per depth
struct RingPos {
    int w2, w3, w4;
    __device__ __forceinline__ void advance()
    {
        w2 ^= 1;
        w3 = w3 == 2 ? 0 : w3 + 1;
        w4 = (w4 + 1) & 3;
    }
};

template <typename T, int V, int W> __device__ __forceinline__ void park_column(T* ring, const RingPos& rp, int t, const T (&f)[V][9])
{
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const int slot = ring_depth(q) == 2 ? rp.w2 : (ring_depth(q) == 3 ? rp.w3 : rp.w4);
        Vec<T, V> p;
#pragma unroll
        for (int v = 0; v < V; ++v) p.v[v] = f[v][q];
        *reinterpret_cast<Vec<T, V>*>(ring + (size_t)(ring_base(q) + slot) * W + t * V) = p;
    }
}

// phase B pull: column x - cx of the ring, row - cy; rp is the position of column x + 1
template <typename T, int V, int W> __device__ __forceinline__ void pull_ring(const T* ring, const RingPos& rp, int t, T (&f)[V][9])
{
    const int r3 = rp.w3 == 0 ? 2 : rp.w3 - 1;  // column x
    const int r4 = (rp.w4 + 2) & 3;             // column x - 1
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const int slot = ring_depth(q) == 2 ? rp.w2 : (ring_depth(q) == 3 ? r3 : r4);
        const T* col = ring + (size_t)(ring_base(q) + slot) * W + t * V;
        const int cy = cyi(q);
        if (cy == 0) {
            const Vec<T, V> p = *reinterpret_cast<const Vec<T, V>*>(col);
#pragma unroll
            for (int v = 0; v < V; ++v) f[v][q] = p.v[v];
        } else {
#pragma unroll
            for (int v = 0; v < V; ++v) f[v][q] = col[v - cy];
        }
    }
}

template <typename T, int MODEL, int V, int NT, int MINB> __global__ void __launch_bounds__(NT, MINB) k_lbm2(const Lbm2Args<T> a)
{
    constexpr int W = NT * V;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* ring = reinterpret_cast<T*>(smem_raw);

    const int strip = blockIdx.x % a.nstrips, seg = blockIdx.x / a.nstrips;
    const int y_lo = strip * a.ty;
    const int y_hi = min(y_lo + a.ty, a.ny);
    const int xs = seg * a.seglen;
    const int xe = min(xs + a.seglen, a.nx);
    const int t = threadIdx.x;
    const int yl = y_lo - V + t * V;                                       // logical first row of this thread
    const bool act_a = yl < y_hi + V;                                      // strip + V halo rows on both sides
    const bool act_b = yl >= y_lo && yl < y_hi;                            // strip interior
    const int yp = yl < 0 ? yl + a.ny : (yl >= a.ny ? yl - a.ny : yl);     // ny % V == 0: a vector never straddles the wrap

    RingPos rp = {0, 0, 0};
    T n[V][9];
    // warm-up: step-1 state of columns xs-1 and xs
    if (act_a) {
        load_column<T, V>(a, xs - 1, yp, n);
#pragma unroll
        for (int v = 0; v < V; ++v) collide<T, MODEL>(n[v], a.cp);
        park_column<T, V, W>(ring, rp, t, n);
    }
    rp.advance();
    if (act_a) {
        load_column<T, V>(a, xs, yp, n);
#pragma unroll
        for (int v = 0; v < V; ++v) collide<T, MODEL>(n[v], a.cp);
        park_column<T, V, W>(ring, rp, t, n);
        load_column<T, V>(a, xs + 1, yp, n);
    }
    rp.advance();

    for (int x = xs; x < xe; ++x) {
        // A: step-1 state of column x+1 (operands were loaded one iteration ago)
        if (act_a) {
#pragma unroll
            for (int v = 0; v < V; ++v) collide<T, MODEL>(n[v], a.cp);
            park_column<T, V, W>(ring, rp, t, n);
        }
        __syncthreads();
        if (act_a && x + 1 < xe) load_column<T, V>(a, x + 2, yp, n);  // in flight during phase B
        // B: step-2 state of column x
        if (act_b) {
            T f[V][9];
            pull_ring<T, V, W>(ring, rp, t, f);
#pragma unroll
            for (int v = 0; v < V; ++v) collide<T, MODEL>(f[v], a.cp);
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                Vec<T, V> p;
#pragma unroll
                for (int v = 0; v < V; ++v) p.v[v] = f[v][q];
                *reinterpret_cast<Vec<T, V>*>(a.dst + ((size_t)q * a.nx + x) * (size_t)a.ld + yp) = p;
            }
        }
        rp.advance();
    }
}

template <typename T, int MODEL> int launch_pair(const Grid& g, const T* src, T* dst, const CollideParams<T>& cp, cudaStream_t s)
{
    constexpr int V = 16 / (int)sizeof(T);
    constexpr int NT = 128, MINB = 4;
    constexpr int W = NT * V;
    constexpr size_t smem = (size_t)RING_SLOTS * W * sizeof(T);
    auto kern = k_lbm2<T, MODEL, V, NT, MINB>;
    static bool configured[64] = {false};
    if (g.device < 64 && !configured[g.device]) {
        PLBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PLBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        configured[g.device] = true;
    }
    Lbm2Args<T> a;
    a.src = src;
    a.dst = dst;
    a.nx = g.nx;
    a.ny = g.ny;
    a.ld = g.ld;
    a.cp = cp;
    const int ty_max = (NT - 2) * V;
    a.nstrips = (g.ny + ty_max - 1) / ty_max;
    a.ty = ((g.ny + a.nstrips - 1) / a.nstrips + V - 1) / V * V;
    a.nstrips = (g.ny + a.ty - 1) / a.ty;
    // one wave of equally long x segments, at least 8 columns each
    int nseg = (g.sm_count * MINB) / a.nstrips;
    if (nseg < 1) nseg = 1;
    a.seglen = (g.nx + nseg - 1) / nseg;
    if (a.seglen < 8) a.seglen = g.nx < 8 ? g.nx : 8;
    nseg = (g.nx + a.seglen - 1) / a.seglen;
    kern<<<(unsigned)(a.nstrips * nseg), NT, smem, s>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PLBM_CUDA(cudaGetLastError());
    return PLBM_OK;
}

}  // namespace

bool lbm_pair_applicable(const Grid& g)
{
    const int v = 16 / (int)g.esize();
    return g.nx >= 4 && g.ny >= 2 * v && (g.ny % v) == 0;
}

// Two fused steps src -> dst.  The caller accounts for the lattice roles (see step_lbm_t).
template <typename T> int launch_lbm_pair(const Grid& g, const T* src, T* dst, int model, const CollideParams<T>& cp, cudaStream_t s)
{
    switch (model) {
    case M_BGK: return launch_pair<T, M_BGK>(g, src, dst, cp, s);
    case M_TRT: return launch_pair<T, M_TRT>(g, src, dst, cp, s);
    case M_RR: return launch_pair<T, M_RR>(g, src, dst, cp, s);
    case M_BGK_SPLIT: return launch_pair<T, M_BGK_SPLIT>(g, src, dst, cp, s);
    case M_TRT_SPLIT: return launch_pair<T, M_TRT_SPLIT>(g, src, dst, cp, s);
    case M_BGK_IMPROVED: return launch_pair<T, M_BGK_IMPROVED>(g, src, dst, cp, s);
    }
    set_error("launch_lbm_pair: unknown collision model");
    return PLBM_ERR_ARG;
}

template int launch_lbm_pair<double>(const Grid&, const double*, double*, int, const CollideParams<double>&, cudaStream_t);
template int launch_lbm_pair<float>(const Grid&, const float*, float*, int, const CollideParams<float>&, cudaStream_t);

}  // namespace plbm
