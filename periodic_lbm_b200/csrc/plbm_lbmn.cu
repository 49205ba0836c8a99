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
    // slab decomposition: the ring neighbours' three nearest lines, all nine populations, [PLBM_HALO_LINES][9][ld] in the order
    // of plbm_internal.h (lo: lines -2, -1, -3; hi: lines nx, nx+1, nx+2).  nullptr = periodic self-wrap.
    const T* halo_lo;
    const T* halo_hi;
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
    // staged raw column: level 1 works on the strip +- NR V rows and pulls one more row on each side; the staged range is rounded
    // out to 16 bytes (bulk copies): HS rows beyond the strip on each side, thread 0's first row sits at stage row OFF
    constexpr int VA = 16 / (int)sizeof(T);
    constexpr int HS = (NR * V + 1 + VA - 1) / VA * VA;
    constexpr int OFF = HS - NR * V;
    constexpr int WS = W + 2 * OFF;
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
    const int r0 = y_lo - HS, r1 = y_hi + HS;

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
            const T* line;
            if (a.halo_lo && col < 0) {
                line = a.halo_lo + ((size_t)halo_lo_index(col) * 9 + q) * (size_t)a.ld;
            } else if (a.halo_hi && col >= a.nx) {
                line = a.halo_hi + ((size_t)(col - a.nx) * 9 + q) * (size_t)a.ld;
            } else {
                col = col < 0 ? col + a.nx : (col >= a.nx ? col - a.nx : col);
                line = a.src + ((size_t)q * a.nx + col) * (size_t)a.ld;
            }
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
            const T* st = stage + ((k & 1) * 9) * WS + OFF + t * V;  // stage row of this thread's first row
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

#ifndef PLBM_MULTI_WIDE_DEFAULT
#define PLBM_MULTI_WIDE_DEFAULT 1
#endif
int env_knob(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

// NT = 128: 74 KB (NSTEP 2, three blocks per SM) / 111 KB (NSTEP 3, two blocks per SM) of shared memory per block;
// NT = 256 (NSTEP 3 only, PLBM_MULTI_NT=256): one 222 KB block of eight warps per SM, strips of up to 504 rows (fp64)
// WIDE: 8 bytes per thread and 256 threads -- the shared memory of the 128-thread shape with twice the warps (NSTEP 3: two blocks
// = sixteen warps per SM instead of eight)
template <typename T, int MODEL, int NSTEP, int NT, bool WIDE = false>
int launch_n(const Grid& g, const T* src, T* dst, int x_begin, int x_end, const T* halo_lo, const T* halo_hi, const CollideParams<T>& cp,
             cudaStream_t s)
{
    constexpr int VA = 16 / (int)sizeof(T);
    constexpr int V = WIDE ? VA / 2 : VA;
    constexpr int HS = ((NSTEP - 1) * V + 1 + VA - 1) / VA * VA;
    constexpr int W = NT * V, WS = W + 2 * (HS - (NSTEP - 1) * V);
    constexpr size_t smem = ((size_t)(NSTEP - 1) * RN_SLOTS * W + 2 * 9 * WS) * sizeof(T) + 16;
    static_assert(smem <= 227 * 1024, "shared memory of one block");
    // blocks per SM: what the shared memory holds (1 KB per block is the system's), at most 2048 threads and 24 blocks
    constexpr int by_smem = (int)((size_t)(228 * 1024) / (smem + 1024));
    constexpr int by_threads = 2048 / NT;
    constexpr int MINB = by_smem < by_threads ? (by_smem < 24 ? by_smem : 24) : (by_threads < 24 ? by_threads : 24);
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
    a.halo_lo = halo_lo;
    a.halo_hi = halo_hi;
    a.cp = cp;
    // the fewest strips (2 (NSTEP-1) V redundant rows each), 64-column segments (2 (NSTEP-1) warm-up columns each)
    const int ncols = x_end - x_begin;
    const int ty_max = (NT - 2 * (NSTEP - 1)) * V;
    a.nstrips = (g.ny + ty_max - 1) / ty_max;
    a.ty = ((g.ny + a.nstrips - 1) / a.nstrips + VA - 1) / VA * VA;
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

template <typename T, int NSTEP, int NT, bool WIDE = false>
int dispatch_n(const Grid& g, const T* src, T* dst, int x_begin, int x_end, const T* halo_lo, const T* halo_hi, int model,
               const CollideParams<T>& cp, cudaStream_t s)
{
    switch (model) {
    case M_BGK: return launch_n<T, M_BGK, NSTEP, NT, WIDE>(g, src, dst, x_begin, x_end, halo_lo, halo_hi, cp, s);
    case M_TRT: return launch_n<T, M_TRT, NSTEP, NT, WIDE>(g, src, dst, x_begin, x_end, halo_lo, halo_hi, cp, s);
    case M_RR: return launch_n<T, M_RR, NSTEP, NT, WIDE>(g, src, dst, x_begin, x_end, halo_lo, halo_hi, cp, s);
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

// THREE steps per pass over HBM (k_lbmn_bulk; bit-identical like the pairs) where they were measured to win on B200.  Shape: one
// row (fp64) per thread, 128-thread blocks, four blocks = sixteen warps per SM ("wide"): the 128-thread, 16-byte shape has the shared
// memory for two blocks = eight warps per SM only and is bound by latency (ncu r02a: fp64 pipe 47 %, DRAM 62 %, nothing saturated),
// 256-thread blocks of one row per thread wait at the three barriers per column (ncu r02f: barrier stall 2.1 per issue).
// GLUPS, pairs (k_lbm2_bulk) -> triples, 16-byte/128 | one-row/256 | one-row/128 (r02a, r02f, r02g):
//   BGK fp64  8192^2  83.3 -> 80.5 |  91.5 |  97.9     bench slab 32768 x 4096  85.3 -> 83.8 | 93.6 | 99.2     2048^2  72.4 -> 68.1 | 73.0 | 75.6
//   TRT fp64  8192^2  83.4 -> 93.6 | 102.5 | 112.7     RR fp64  8192^2  68.5 -> 57.8 | 66.4 | 69.1
//   BGK fp32  8192^2 149.3 -> 119.5 | 136.4 | 148.2    RR fp32  8192^2 102.5 -> 83.8 | 100.6 | 107.6
// TRT has no division and BGK one, so a third collision fits under the fp64 pipe once enough warps hide its latency; RR fp64 is at
// the fp64 pipe either way and fp32 gains little: both stay on pairs, and so do small grids (1024^2 TRT: 63.0 -> 25.3, too few
// blocks for 64-column segments).  PLBM_TRIPLES=0: never; 2: BGK / TRT fp64 at every size (tests).
int lbm_triples_level(const Grid& g)
{
    if (g.nx < 2 * PLBM_HALO_LINES || !lbm_multi_applicable(g, M_BGK, 3)) return -1;
    const long long nodes = (long long)g.nx * g.ny;
    return nodes >= 4096LL * 4096LL ? 2 : (nodes >= 2048LL * 2048LL ? 1 : 0);
}

bool lbm_triples_wanted(const Grid& g, int level, int model)
{
    static const int mode = env_knob("PLBM_TRIPLES", 1);
    if (mode == 0 || level < 0 || g.variant != 0 || g.prec != PLBM_F64) return false;
    if (mode >= 2) return model == M_TRT || model == M_BGK;
    return (model == M_TRT || model == M_BGK) && level >= 1;
}

// `nstep` (2 or 3) fused steps src -> dst for columns [x_begin, x_end).  halo_lo / halo_hi: the ring neighbours' nearest lines
// under a slab decomposition ([PLBM_HALO_LINES][9][ld]), nullptr = periodic self-wrap.
template <typename T>
int launch_lbm_multi(const Grid& g, const T* src, T* dst, int x_begin, int x_end, int model, const CollideParams<T>& cp, int nstep,
                     cudaStream_t s, const T* halo_lo, const T* halo_hi)
{
    if (!lbm_multi_applicable(g, model, nstep)) {
        set_error("launch_lbm_multi: not applicable to this grid / collision / depth");
        return PLBM_ERR_ARG;
    }
#define PLBM_N(NS, NT, WD) return dispatch_n<T, NS, NT, WD>(g, src, dst, x_begin, x_end, halo_lo, halo_hi, model, cp, s)
    if (nstep == 2) PLBM_N(2, 128, false);
    static const int nt = env_knob("PLBM_MULTI_NT", 0);  // 0: the default of the shape (128 threads)
    static const int wide = env_knob("PLBM_MULTI_WIDE", PLBM_MULTI_WIDE_DEFAULT);
    // small blocks (measurement knobs): the barriers of a level couple two warps / one warp instead of eight
    if (wide && nt == 64) PLBM_N(3, 64, true);
    if (wide && nt == 32) PLBM_N(3, 32, true);
    if (wide && nt == 256) PLBM_N(3, 256, true);
    if (wide) PLBM_N(3, 128, true);  // the default: four blocks of four warps per SM
    if (nt == 64) PLBM_N(3, 64, false);
    if (nt == 32) PLBM_N(3, 32, false);
    if (nt == 256) PLBM_N(3, 256, false);
    PLBM_N(3, 128, false);
#undef PLBM_N
}

template int launch_lbm_multi<double>(const Grid&, const double*, double*, int, int, int, const CollideParams<double>&, int, cudaStream_t,
                                      const double*, const double*);
template int launch_lbm_multi<float>(const Grid&, const float*, float*, int, int, int, const CollideParams<float>&, int, cudaStream_t,
                                     const float*, const float*);

}  // namespace plbm
