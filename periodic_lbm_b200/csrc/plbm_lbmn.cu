// plbm_lbmn.cu -- NSTEP fused stream+collide steps per pass over HBM (temporal blocking of depth NSTEP, sm_100a).
//
// EXPERIMENTAL (variants 9 and 10 of perform_lbm_step; never selected by default): the generalisation of
// k_lbm2_bulk (plbm_lbm2.cu) from two to NSTEP levels.  k_lbm2_bulk moves 73.6 B of DRAM traffic per update at
// 0.97 of the measured HBM copy rate, i.e. it sits on the HBM roof of a two-step scheme; the only way further up
// is fewer bytes per update: 144 / NSTEP B (fp64).  Not yet run on a GPU when this was written -- the round's
// GPU budget was spent; tests/test_gpu_parity.py::test_multi_step_kernel_experimental is the parity gate
// (set PLBM_TEST_EXPERIMENTAL=1) and tools/pair_ab.py --variants 7,9,10 the A/B.
//
// A block owns a strip of rows and marches along x.  In iteration x, after the raw column x + NSTEP - 1 has
// landed in shared memory by bulk async copies (issued by one thread two columns ahead, one mbarrier per stage),
//   level 1      collides the streamed raw column x + NSTEP - 1            -> ring 0   (state after step 1)
//   level l      pulls column x + NSTEP - l from ring l-2 (columns c-1, c, c+1), collides -> ring l-1
//   level NSTEP  pulls column x from ring NSTEP-2, collides, stores the state after step NSTEP to `dst`.
// Rows: thread t owns V rows from y_lo - (NSTEP-1) V + t V; level l is active on the strip +- (NSTEP - l) V rows.
// Every ring keeps a population 1 / 2 / 3 columns (cx = -1 / 0 / +1): 18 column slots, one barrier after each
// level.  Warm-up: the loop starts 2 (NSTEP - 1) columns before the segment and level l joins 2 (l - 1) iterations
// later, so NSTEP = 2 is exactly k_lbm2_bulk's schedule.  Same collide<T,MODEL> on the same operands as NSTEP
// k_lbm launches -> bit-identical results.
//   lbm_stream_kernel  src/periodic_lbm.f90:45-127 ;  collisions src/collision_*.F90
#include <cstdint>
#include <cstdlib>

#include "plbm_internal.h"

namespace plbm {

namespace {

template <typename T, int V> struct alignas(sizeof(T) * V) VecN {
    T v[V];
};

// ring slots of population q: (3,6,7) live one column | (0,2,4) two | (1,5,8) three
__host__ __device__ constexpr int rn_depth(int q) { return cxi(q) == -1 ? 1 : (cxi(q) == 0 ? 2 : 3); }
__host__ __device__ constexpr int rn_base(int q)
{
    return q == 3 ? 0 : q == 6 ? 1 : q == 7 ? 2 : q == 0 ? 3 : q == 2 ? 5 : q == 4 ? 7 : q == 1 ? 9 : q == 5 ? 12 : 15;
}
constexpr int RN_SLOTS = 18;

template <typename T> struct LbmNArgs {
    const T* src;
    T* dst;
    int nx, ny, ld;
    int x_begin, x_end;  // columns whose final state this launch writes
    int ty;              // interior rows per strip (multiple of V)
    int nstrips;         // strips along y
    int seglen;          // columns per x segment
    CollideParams<T> cp;
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarrier_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarrier_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarrier_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok = 0;
    long long spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!ok && ++spins > (1ll << 26)) __trap();  // a lost copy must fail loudly, never hang the GPU
    } while (!ok);
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
                 : "memory");
}

// pull the V rows of this thread from nine staged / ring columns: `col(q)` = row 0 of this thread in population q
template <typename T, int V, typename ColOf> __device__ __forceinline__ void pull_rows(ColOf col, T (&f)[V][9])
{
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const T* c = col(q);
        const int cy = cyi(q);
        if (cy == 0) {
            const VecN<T, V> p = *reinterpret_cast<const VecN<T, V>*>(c);
#pragma unroll
            for (int v = 0; v < V; ++v) f[v][q] = p.v[v];
        } else {
#pragma unroll
            for (int v = 0; v < V; ++v) f[v][q] = c[v - cy];
        }
    }
}

template <typename T, int MODEL, int V, int NT, int MINB, int NSTEP>
__global__ void __launch_bounds__(NT, MINB) k_lbmn_bulk(const LbmNArgs<T> a)
{
    static_assert(NSTEP >= 2, "depth of the temporal blocking");
    constexpr int NR = NSTEP - 1;  // rings
    constexpr int W = NT * V;      // rows of one ring column: logical rows y_lo - NR V .. y_lo - NR V + W - 1
    constexpr int WS = W + 2 * V;  // rows of one staged raw column: logical rows y_lo - NSTEP V .. (one more vector each side)
    extern __shared__ __align__(128) unsigned char smem_n[];
    T* ring = reinterpret_cast<T*>(smem_n);  // [NR][RN_SLOTS][W]
    T* stage = ring + NR * RN_SLOTS * W;     // [2][9][WS]
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage + 2 * 9 * WS);
    const uint32_t bar0 = smem_addr(bars);
    auto bar = [&](int s) { return bar0 + 8u * (uint32_t)s; };  // one mbarrier per stage

    const int strip = blockIdx.x % a.nstrips, seg = blockIdx.x / a.nstrips;
    const int y_lo = strip * a.ty;
    const int y_hi = min(y_lo + a.ty, a.ny);
    const int xs = a.x_begin + seg * a.seglen;
    const int xe = min(xs + a.seglen, a.x_end);
    const int t = threadIdx.x;
    const int yl = y_lo - NR * V + t * V;  // logical first row of this thread
    const int yp = yl < 0 ? yl + a.ny : (yl >= a.ny ? yl - a.ny : yl);
    // level l works on the strip widened by (NSTEP - l) V rows on both sides
    auto active = [&](int l) {
        const int h = (NSTEP - l) * V;
        return yl >= y_lo - h && yl < y_hi + h;
    };
    // staged logical rows [r0, r1): what level 1 reads, whole vectors
    const int r0 = y_lo - NSTEP * V, r1 = y_hi + NSTEP * V;

    if (t == 0) {
        mbarrier_init(bar(0), 1);
        mbarrier_init(bar(1), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // one raw column (all nine populations, pulled: population q from column xl - cx_q) into stage s
    auto issue = [&](int xl, int s) {
        mbarrier_expect_tx(bar(s), (uint32_t)(9 * (r1 - r0) * sizeof(T)));
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            int col = xl - cxi(q);
            col = col < 0 ? col + a.nx : (col >= a.nx ? col - a.nx : col);
            const T* line = a.src + ((size_t)q * a.nx + col) * (size_t)a.ld;
            T* dst = stage + (s * 9 + q) * WS;
            // periodic pieces of [r0, r1): below 0, inside, beyond ny (all multiples of V rows = 16 bytes)
            if (r0 < 0) bulk_copy_g2s(smem_addr(dst), line + (a.ny + r0), (uint32_t)(-r0 * sizeof(T)), bar(s));
            const int m0 = max(r0, 0), m1 = min(r1, a.ny);
            bulk_copy_g2s(smem_addr(dst + (m0 - r0)), line + m0, (uint32_t)((m1 - m0) * sizeof(T)), bar(s));
            if (r1 > a.ny) bulk_copy_g2s(smem_addr(dst + (a.ny - r0)), line, (uint32_t)((r1 - a.ny) * sizeof(T)), bar(s));
        }
    };

    const int x_first = xs - 2 * NR;  // first iteration (warm-up of the rings)
    const int c_last = xe - 1 + NR;   // raw column of the last iteration; iteration x consumes raw column x + NR
    if (t == 0) {
        issue(x_first + NR, 0);
        if (x_first + NR + 1 <= c_last) issue(x_first + NR + 1, 1);
    }
    int w2 = 0, w3 = 0;  // slot of the column written in this iteration (two- and three-column populations), all rings
    for (int x = x_first; x < xe; ++x) {
        const int k = x - x_first;  // raw column number: stage k & 1, phase (k >> 1) & 1 of its barrier
        const int r2 = w2 ^ 1;                // written one iteration ago
        const int r3 = w3 == 2 ? 0 : w3 + 1;  // written two iterations ago
        // ---- level 1: streamed raw column x + NR -> ring 0
        mbarrier_wait(bar(k & 1), (uint32_t)((k >> 1) & 1));
        if (active(1)) {
            T n[V][9];
            const T* st = stage + ((k & 1) * 9) * WS + V + t * V;  // stage row of this thread's first row
            pull_rows<T, V>([&](int q) { return st + q * WS; }, n);
#pragma unroll
            for (int v = 0; v < V; ++v) collide<T, MODEL>(n[v], a.cp);
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                const int slot = rn_depth(q) == 1 ? 0 : (rn_depth(q) == 2 ? w2 : w3);
                VecN<T, V> p;
#pragma unroll
                for (int v = 0; v < V; ++v) p.v[v] = n[v][q];
                *reinterpret_cast<VecN<T, V>*>(ring + (rn_base(q) + slot) * W + t * V) = p;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // stage reads before the async refill
        __syncthreads();
        if (t == 0 && x + NR + 2 <= c_last) issue(x + NR + 2, k & 1);  // lands during the next two iterations
        // ---- levels 2 .. NSTEP: column x + NSTEP - l from ring l-2 -> ring l-1 (the last one -> dst)
#pragma unroll
        for (int l = 2; l <= NSTEP; ++l) {
            if (x >= xs - 2 * (NSTEP - l) && active(l)) {
                T f[V][9];
                const T* rin = ring + (l - 2) * RN_SLOTS * W + t * V;
                pull_rows<T, V>(
                    [&](int q) {
                        const int slot = rn_depth(q) == 1 ? 0 : (rn_depth(q) == 2 ? r2 : r3);
                        return rin + (rn_base(q) + slot) * W;
                    },
                    f);
#pragma unroll
                for (int v = 0; v < V; ++v) collide<T, MODEL>(f[v], a.cp);
                if (l == NSTEP) {
#pragma unroll
                    for (int q = 0; q < 9; ++q) {
                        VecN<T, V> p;
#pragma unroll
                        for (int v = 0; v < V; ++v) p.v[v] = f[v][q];
                        *reinterpret_cast<VecN<T, V>*>(a.dst + ((size_t)q * a.nx + x) * (size_t)a.ld + yp) = p;
                    }
                } else {
                    T* rout = ring + (l - 1) * RN_SLOTS * W + t * V;
#pragma unroll
                    for (int q = 0; q < 9; ++q) {
                        const int slot = rn_depth(q) == 1 ? 0 : (rn_depth(q) == 2 ? w2 : w3);
                        VecN<T, V> p;
#pragma unroll
                        for (int v = 0; v < V; ++v) p.v[v] = f[v][q];
                        *reinterpret_cast<VecN<T, V>*>(rout + (rn_base(q) + slot) * W) = p;
                    }
                }
            }
            __syncthreads();  // ring l-1 complete before level l+1 reads it; after the last level: slots may be rewritten
        }
        w2 ^= 1;
        w3 = w3 == 2 ? 0 : w3 + 1;
    }
}

int env_knob(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

// NT = 128: 74 KB (NSTEP 2, three blocks per SM) / 111 KB (NSTEP 3, two blocks per SM) of shared memory per block;
// NT = 256 (NSTEP 3 only, PLBM_MULTI_NT=256): one 222 KB block of eight warps per SM, strips of up to 504 rows (fp64)
template <typename T, int MODEL, int NSTEP, int NT>
int launch_n(const Grid& g, const T* src, T* dst, int x_begin, int x_end, const CollideParams<T>& cp, cudaStream_t s)
{
    constexpr int V = 16 / (int)sizeof(T);
    constexpr int MINB = NT == 256 ? 1 : (NSTEP == 2 ? 3 : 2);
    constexpr int W = NT * V, WS = W + 2 * V;
    constexpr size_t smem = ((size_t)(NSTEP - 1) * RN_SLOTS * W + 2 * 9 * WS) * sizeof(T) + 16;
    static_assert(smem <= 227 * 1024, "shared memory of one block");
    if (x_end <= x_begin) return PLBM_OK;
    auto kern = k_lbmn_bulk<T, MODEL, V, NT, MINB, NSTEP>;
    static bool configured[64] = {false};
    if (g.device < 64 && !configured[g.device]) {
        PLBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PLBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        configured[g.device] = true;
    }
    LbmNArgs<T> a;
    a.src = src;
    a.dst = dst;
    a.nx = g.nx;
    a.ny = g.ny;
    a.ld = g.ld;
    a.x_begin = x_begin;
    a.x_end = x_end;
    a.cp = cp;
    // the fewest strips (2 (NSTEP-1) V redundant rows each), 64-column segments (2 (NSTEP-1) warm-up columns each)
    const int ncols = x_end - x_begin;
    const int ty_max = (NT - 2 * (NSTEP - 1)) * V;
    a.nstrips = (g.ny + ty_max - 1) / ty_max;
    a.ty = ((g.ny + a.nstrips - 1) / a.nstrips + V - 1) / V * V;
    a.nstrips = (g.ny + a.ty - 1) / a.ty;
    static const int seg_cols = env_knob("PLBM_MULTI_SEGLEN", 64) < 1 ? 64 : env_knob("PLBM_MULTI_SEGLEN", 64);
    int nseg = (ncols + seg_cols - 1) / seg_cols;
    a.seglen = (ncols + nseg - 1) / nseg;
    nseg = (ncols + a.seglen - 1) / a.seglen;
    kern<<<(unsigned)(a.nstrips * nseg), NT, smem, s>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PLBM_CUDA(cudaGetLastError());
    return PLBM_OK;
}

template <typename T, int NSTEP, int NT>
int dispatch_n(const Grid& g, const T* src, T* dst, int x_begin, int x_end, int model, const CollideParams<T>& cp, cudaStream_t s)
{
    switch (model) {
    case M_BGK: return launch_n<T, M_BGK, NSTEP, NT>(g, src, dst, x_begin, x_end, cp, s);
    case M_TRT: return launch_n<T, M_TRT, NSTEP, NT>(g, src, dst, x_begin, x_end, cp, s);
    case M_RR: return launch_n<T, M_RR, NSTEP, NT>(g, src, dst, x_begin, x_end, cp, s);
    }
    set_error("launch_lbm_multi: collision model not instantiated for the experimental multi-step kernel");
    return PLBM_ERR_ARG;
}

}  // namespace

// the experimental kernel is instantiated for the three reference collision operators of the LBM path, on grids the
// bulk copies can address (every staged piece a multiple of 16 bytes, the wrap pieces inside one line)
bool lbm_multi_applicable(const Grid& g, int model, int nstep)
{
    const int v = 16 / (int)g.esize();
    return (nstep == 2 || nstep == 3) && (model == M_BGK || model == M_TRT || model == M_RR) && g.nx >= 4 && (g.ny % v) == 0 &&
           g.ny >= 8 * v;
}

// `nstep` (2 or 3) fused steps src -> dst for columns [x_begin, x_end), periodic self-wrap (single GPU).
template <typename T>
int launch_lbm_multi(const Grid& g, const T* src, T* dst, int x_begin, int x_end, int model, const CollideParams<T>& cp, int nstep,
                     cudaStream_t s)
{
    if (!lbm_multi_applicable(g, model, nstep)) {
        set_error("launch_lbm_multi: not applicable to this grid / collision / depth");
        return PLBM_ERR_ARG;
    }
    if (nstep == 2) return dispatch_n<T, 2, 128>(g, src, dst, x_begin, x_end, model, cp, s);
    static const int nt = env_knob("PLBM_MULTI_NT", 128);
    if (nt == 256) return dispatch_n<T, 3, 256>(g, src, dst, x_begin, x_end, model, cp, s);
    return dispatch_n<T, 3, 128>(g, src, dst, x_begin, x_end, model, cp, s);
}

template int launch_lbm_multi<double>(const Grid&, const double*, double*, int, int, int, const CollideParams<double>&, int, cudaStream_t);
template int launch_lbm_multi<float>(const Grid&, const float*, float*, int, int, int, const CollideParams<float>&, int, cudaStream_t);

}  // namespace plbm
