// plbm_lbmn.cu -- NSTEP fused stream+collide steps per pass over HBM (temporal blocking of depth NSTEP, sm_100a).
//
// With NSTEP = 3 this was the headline kernel of round 2 until k_lbm3_ws (plbm_lbm3w.cu: the same three steps with the levels skewed
// and a producer warp) took over the launches that read no halo lines; k_lbmn_bulk<3> remains what perform_lbm_step launches for the
// two-relaxation-time collisions, for the boundary launches of a slab (HALO) and under PLBM_TRIPLE_WS=0 (lbm_triple_ws_wanted below;
// variants 9 / 10 force its NSTEP = 2 / 3 instances on every grid they apply to).  DUAL: the launch that closes a call also stores the
// state after step NSTEP - 1 (Grid::spare, plbm_api.cu step_lbm_t).  It is the generalisation of k_lbm2_bulk (plbm_lbm2.cu) from two to
// NSTEP levels: k_lbm2_bulk moves 73.6 B of DRAM traffic per update at 0.97 of the measured HBM copy rate, i.e. it sits on the HBM
// roof of a two-step scheme, and the only way further up is fewer bytes per update: 144 / NSTEP B (fp64; measured 49.4 B with
// NSTEP = 3).
//
// A block owns a strip of rows and marches along x.  In iteration x, after the raw column x + NSTEP - 1 has
// landed in shared memory by bulk async copies (issued two columns ahead by lane 0 of every warp, one mbarrier per stage),
//   level 1      collides the streamed raw column x + NSTEP - 1            -> ring 0   (state after step 1)
//   level l      pulls column x + NSTEP - l from ring l-2 (columns c-1, c, c+1), collides -> ring l-1
//   level NSTEP  pulls column x from ring NSTEP-2, collides, stores the state after step NSTEP to `dst`.
// Rows: thread t owns V rows from y_lo - (NSTEP-1) V + t V; level l is active on the strip +- (NSTEP - l) V rows.
// Every ring keeps a population 1 / 2 / 3 columns (cx = -1 / 0 / +1): 18 column slots, one barrier after each
// level.  Warm-up: the loop starts 2 (NSTEP - 1) columns before the segment and level l joins 2 (l - 1) iterations
// later, so NSTEP = 2 is exactly k_lbm2_bulk's schedule.  Same collide<T,MODEL> on the same operands as NSTEP
// k_lbm launches -> bit-identical results (tests/test_gpu_parity.py; the schedule itself is restated in numpy and
// checked against the CPU restatement of the reference: tests/test_multi_step_schedule.py).
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
    T* dst_mid;  // DUAL: the state after step NSTEP - 1 goes here as well (the launch that closes a call, see step_lbm_t)
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

// HALO: columns below 0 / beyond nx - 1 come from the ring neighbours' halo lines (the boundary launches of a slab).  A template
// parameter, not a run-time test: the copies are issued by ONE thread after the first barrier of every column while its warp
// waits, and with the extra branches in that path the interior launch of the bench slab ran 4.75 ms instead of 4.0 (r02h).
template <typename T, int MODEL, int V, int NT, int MINB, int NSTEP, bool HALO, bool DUAL = false>
__global__ void __launch_bounds__(NT, MINB) k_lbmn_bulk(const LbmNArgs<T> a)
{
    static_assert(NSTEP >= 2, "depth of the temporal blocking");
    // fp32: two nodes per instruction (FFMA2 pairs, plbm_f32x2.cuh; bit-identical to the scalar collisions)
    // -- except for the recursive-regularized collision, which is bound by the fp32 pipe here and loses with the pairs (an FFMA2
    // occupies the pipe for two cycles and the pairing costs register moves): RR fp32 8192^2 112.8 scalar vs 104.3 packed GLUPS,
    // BGK fp32 160.6 vs 164.3 (r02j / r02k)
    constexpr bool PACKED = sizeof(T) == 4 && (V % 2) == 0 && MODEL != M_RR;
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

    // The copies of a column are issued by lane 0 of EVERY warp, population q by warp q mod NW: the address arithmetic of nine
    // population lines (up to three pieces each) is several hundred instructions, and one thread issuing all of it while its
    // warp waits made warp 0 the slowest of every column.  Each issuing lane announces its own bytes: NW arrivals per phase.
    constexpr int NW = NT / 32;
    static_assert(NW >= 1 && NW <= 9, "one to nine issuing warps");
    const int warp = t >> 5;
    const bool issuer = (t & 31) == 0;
    if (t == 0) {
        mbarrier_init(bar(0), NW);
        mbarrier_init(bar(1), NW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // one raw column (all nine populations, pulled: population q from column xl - cx_q) into stage s
    auto issue = [&](int xl, int s) {
        const int np = (9 - warp + NW - 1) / NW;  // populations warp, warp + NW, ... < 9
        mbarrier_expect_tx(bar(s), (uint32_t)(np * (r1 - r0) * sizeof(T)));
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            if (q % NW != warp) continue;
            int col = xl - cxi(q);
            const T* line;
            if (HALO && col < 0) {
                line = a.halo_lo + ((size_t)halo_lo_index(col) * 9 + q) * (size_t)a.ld;
            } else if (HALO && col >= a.nx) {
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

    // Without halo lines (every launch but the boundary launches of a slab) the raw columns are issued strictly one after the
    // other, so an issuing lane keeps the line address of the NEXT column of each population it owns and the three periodic
    // pieces of the staged rows (block constants) instead of redoing the index arithmetic per column: ncu (r02i) counted a fifth of
    // all issued instructions on the uniform datapath, i.e. in that arithmetic.
    const int npop = (9 - warp + NW - 1) / NW;
    const T* nxt[3] = {nullptr, nullptr, nullptr};
    int ncol[3] = {0, 0, 0};
    uint32_t dsto[3] = {0, 0, 0};
    const int m0 = max(r0, 0), m1 = min(r1, a.ny);
    const uint32_t p1_bytes = r0 < 0 ? (uint32_t)(-r0 * sizeof(T)) : 0u, p3_bytes = r1 > a.ny ? (uint32_t)((r1 - a.ny) * sizeof(T)) : 0u;
    const uint32_t p2_bytes = (uint32_t)((m1 - m0) * sizeof(T));
    const uint32_t p2_dst = (uint32_t)((m0 - r0) * sizeof(T)), p3_dst = (uint32_t)((a.ny - r0) * sizeof(T));
    if (!HALO && issuer) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            if (j >= npop) continue;
            const int q = warp + j * NW;
            int col = x_first + NR - cxi(q);
            col = col < 0 ? col + a.nx : (col >= a.nx ? col - a.nx : col);
            ncol[j] = col;
            nxt[j] = a.src + ((size_t)q * a.nx + col) * (size_t)a.ld;
            dsto[j] = smem_addr(stage + q * WS);
        }
    }
    auto issue_next = [&](int s) {
        mbarrier_expect_tx(bar(s), (uint32_t)(npop * (r1 - r0) * sizeof(T)));
        const uint32_t so = (uint32_t)(s * 9 * WS * sizeof(T));
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            if (j >= npop) continue;
            const T* line = nxt[j];
            const uint32_t d = dsto[j] + so;
            if (p1_bytes) bulk_copy_g2s(d, line + (a.ny + r0), p1_bytes, bar(s));
            bulk_copy_g2s(d + p2_dst, line + m0, p2_bytes, bar(s));
            if (p3_bytes) bulk_copy_g2s(d + p3_dst, line, p3_bytes, bar(s));
            if (++ncol[j] == a.nx) {  // the next column wraps around x
                ncol[j] = 0;
                nxt[j] -= (size_t)(a.nx - 1) * (size_t)a.ld;
            } else {
                nxt[j] += a.ld;
            }
        }
    };
    // column xl into stage s; the calls walk through the columns x_first + NR, x_first + NR + 1, ... without gaps
    auto issue_col = [&](int xl, int s) {
        if (HALO) issue(xl, s);
        else issue_next(s);
    };
    if (issuer) {
        issue_col(x_first + NR, 0);
        if (x_first + NR + 1 <= c_last) issue_col(x_first + NR + 1, 1);
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
            collide_nodes<T, MODEL, V, PACKED>(n, a.cp);
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
        if (issuer && x + NR + 2 <= c_last) issue_col(x + NR + 2, k & 1);  // lands during the next two iterations
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
                collide_nodes<T, MODEL, V, PACKED>(f, a.cp);
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
                    if (DUAL && l == NSTEP - 1 && x + 1 >= xs && x + 1 < xe && active(NSTEP)) {  // column x + 1, the strip's own rows
#pragma unroll
                        for (int q = 0; q < 9; ++q) {
                            VecN<T, V> p;
#pragma unroll
                            for (int v = 0; v < V; ++v) p.v[v] = f[v][q];
                            *reinterpret_cast<VecN<T, V>*>(a.dst_mid + ((size_t)q * a.nx + (x + 1)) * (size_t)a.ld + yp) = p;
                        }
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
template <typename T, int MODEL, int NSTEP, int NT, bool WIDE = false, bool DUAL = false>
int launch_n(const Grid& g, const T* src, T* dst, int x_begin, int x_end, const T* halo_lo, const T* halo_hi, const CollideParams<T>& cp,
             cudaStream_t s, T* dst_mid = nullptr)
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
    const bool halo = halo_lo && halo_hi;
    auto kern = halo ? k_lbmn_bulk<T, MODEL, V, NT, MINB, NSTEP, true, DUAL> : k_lbmn_bulk<T, MODEL, V, NT, MINB, NSTEP, false, DUAL>;
    static bool configured[64][2] = {{false}};
    if (g.device < 64 && !configured[g.device][halo]) {
        PLBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PLBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        configured[g.device][halo] = true;
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
    a.dst_mid = dst_mid;
    // the fewest strips (2 (NSTEP-1) V redundant rows each), 64-column segments (2 (NSTEP-1) warm-up columns each)
    const int ncols = x_end - x_begin;
    const int ty_max = (NT - 2 * (NSTEP - 1)) * V;
    a.nstrips = (g.ny + ty_max - 1) / ty_max;
    a.ty = ((g.ny + a.nstrips - 1) / a.nstrips + VA - 1) / VA * VA;
    a.nstrips = (g.ny + a.ty - 1) / a.ty;
    static const int seg_cols = env_knob("PLBM_MULTI_SEGLEN", 64) < 1 ? 64 : env_knob("PLBM_MULTI_SEGLEN", 64);
    int nseg = (ncols + seg_cols - 1) / seg_cols;
    // like k_lbm2_bulk's launcher: the segment count that makes the blocks fill a whole number of rounds of MINB blocks per SM,
    // from below (a last, mostly empty round costs as much as a full one); at least 8 columns per segment
    static const int fill = env_knob("PLBM_MULTI_FILL", 1);
    if (fill) {
        const long long slots = (long long)MINB * g.sm_count;
        const long long blocks64 = (long long)a.nstrips * nseg;
        const long long rounds = blocks64 >= slots ? (blocks64 + slots - 1) / slots : 1;
        nseg = (int)(rounds * slots / a.nstrips);
        if (nseg < 1) nseg = 1;
    }
    a.seglen = (ncols + nseg - 1) / nseg;
    if (a.seglen < 8) a.seglen = ncols < 8 ? ncols : 8;
    nseg = (ncols + a.seglen - 1) / a.seglen;
    kern<<<(unsigned)(a.nstrips * nseg), NT, smem, s>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PLBM_CUDA(cudaGetLastError());
    return PLBM_OK;
}

template <typename T, int NSTEP, int NT, bool WIDE = false, bool DUAL = false>
int dispatch_n(const Grid& g, const T* src, T* dst, int x_begin, int x_end, const T* halo_lo, const T* halo_hi, int model,
               const CollideParams<T>& cp, cudaStream_t s, T* dst_mid = nullptr)
{
#define PLBM_M(M) return launch_n<T, M, NSTEP, NT, WIDE, DUAL>(g, src, dst, x_begin, x_end, halo_lo, halo_hi, cp, s, dst_mid)
    switch (model) {
    case M_BGK: PLBM_M(M_BGK);
    case M_TRT: PLBM_M(M_TRT);
    case M_RR: PLBM_M(M_RR);
    }
    // the -DSPLIT operators and collide_bgk_improved: the default shape only
    if constexpr (NSTEP == 3 && NT == 128 && WIDE) {
        switch (model) {
        case M_BGK_SPLIT: PLBM_M(M_BGK_SPLIT);
        case M_TRT_SPLIT: PLBM_M(M_TRT_SPLIT);
        case M_BGK_IMPROVED: PLBM_M(M_BGK_IMPROVED);
        }
    }
#undef PLBM_M
    set_error("launch_lbm_multi: collision model not instantiated for the multi-step kernel");
    return PLBM_ERR_ARG;
}

}  // namespace

// the experimental kernel is instantiated for the three reference collision operators of the LBM path, on grids the
// bulk copies can address (every staged piece a multiple of 16 bytes, the wrap pieces inside one line)
bool lbm_multi_shape_is_default();
bool lbm_triple_ws_wanted(int model);
bool lbm_multi_applicable(const Grid& g, int model, int nstep)
{
    const int v = 16 / (int)g.esize();
    const bool reference3 = model == M_BGK || model == M_TRT || model == M_RR;
    const bool others = (model == M_BGK_SPLIT || model == M_TRT_SPLIT || model == M_BGK_IMPROVED) && nstep == 3 && lbm_multi_shape_is_default();
    return (nstep == 2 || nstep == 3) && (reference3 || others) && g.nx >= 4 && (g.ny % v) == 0 && g.ny >= 8 * v;
}

// THREE steps per pass over HBM (k_lbmn_bulk; bit-identical like the pairs) is the default of perform_lbm_step wherever the kernel
// applies, from 512^2 nodes (smaller grids: the cluster-resident kernel or k_lbm2).  Shape: one row (fp64) / two rows (fp32) per
// thread, 128-thread blocks, four blocks = sixteen warps per SM; the copies of a column are issued by lane 0 of every warp from
// running line addresses; the segments are cut so that the blocks fill whole rounds.  How it got there (BGK fp64 8192^2, GLUPS;
// two steps per pass, k_lbm2_bulk: 83.3 at 0.97 of the HBM peak in real bytes):
//   16 bytes per thread, 128 threads, 2 blocks per SM (round 1's shape)      80.5   ncu r02a: 8 warps per SM, fp64 pipe 47 %, DRAM 62 %:
//                                                                                   nothing saturated, bound by latency
//   one row per thread, 256 threads, 2 blocks                                91.5   ncu r02f: barrier stall 2.1 per issue
//   one row per thread, 128 threads, 4 blocks                                97.9   (64 / 32 threads: 85.6 / 82.8, r02g)
//   + halo test compiled out of the interior launches, copies issued by
//     every warp (population q by warp q mod 4) instead of by thread 0      100.2   ncu r02i: a fifth of the issued instructions on
//                                                                                   the uniform datapath = address arithmetic of the copies
//   + running line addresses and block-constant pieces                      106.3   ncu r02j: DRAM 5.4 TB/s = 0.84 of the measured
//                                                                                   peak at 49.4 B/update, fp64 pipe 61 %, issue 65 %
// Pairs -> triples at the final shape (r02k, GLUPS; fp64 | fp32):
//   1024^2  BGK 56.7 -> 71.5 | 71.6 -> 89.0    TRT 62.7 -> 81.1 | 83.8 -> 100.6    RR 41.5 -> 47.6 | 48.7 -> 59.0
//   2048^2  BGK 73.1 -> 87.5 | 106.8 -> 126.6  TRT 70.6 -> 101.7 | 121.8 -> 140.2  RR 53.9 -> 57.5 | 73.2 -> 81.5
//   4096^2  BGK 77.7 -> 98.4 | 131.4 -> 150.1  TRT 77.6 -> 114.6 | 150.9 -> 168.1  RR 62.4 -> 65.0 | 90.9 -> 96.3
//   8192^2  BGK 83.2 -> 106.3 | 150.1 -> 164.3 TRT 83.0 -> 123.1 | 162.1 -> 183.7  RR 68.7 -> 69.4 | 102.8 -> 112.8 (scalar collisions)
//   512^2 TRT fp64: k_lbm2 31.4 -> 41.2;  bench slab 32768 x 4096 BGK fp64: 85.3 -> 105.6
// PLBM_TRIPLES=0: never (pairs as in round 1); 2: at every size the kernel applies to (tests).
int lbm_triples_level(const Grid& g)
{
    if (g.nx < 2 * PLBM_HALO_LINES || !lbm_multi_applicable(g, M_BGK, 3)) return -1;
    return (long long)g.nx * g.ny >= 512LL * 512LL ? 1 : 0;
}

// PLBM_TRIPLES=2 (tests): triples wherever the kernel applies, also where the cluster-resident kernel would take the whole call
bool lbm_triples_forced()
{
    static const bool forced = env_knob("PLBM_TRIPLES", 1) >= 2;
    return forced;
}

bool lbm_triples_wanted(const Grid& g, int level, int model)
{
    static const int mode = env_knob("PLBM_TRIPLES", 1);
    if (mode == 0 || level < 0 || g.variant != 0 || !lbm_multi_applicable(g, model, 3)) return false;
    return mode >= 2 || level >= 1;
}

// Which kernel takes the launches of a triple that read no halo lines: k_lbm3_ws (plbm_lbm3w.cu: producer warp, skewed levels, one
// barrier per column, three blocks per SM) or k_lbmn_bulk (this file: a barrier after every level, four blocks per SM).  Measured on
// a B200 (r02q / r02r, 20-step calls closed by a dual triple, GLUPS, k_lbmn_bulk -> k_lbm3_ws with 128-column segments):
//   bench slab 32768 x 4096 BGK fp64 101.7 -> 103.5     8192^2 BGK fp64 99.9 -> 101.6   RR fp64 68.9 -> 70.8   TRT fp64 110.7 -> 107.8
//   8192^2 BGK fp32 161.6 -> 164.5   RR fp32 112.0 -> 114.4     2048^2 BGK fp64 82.3 -> 88.2     1024^2 TRT fp64 74.5 -> 76.4
// k_lbm3_ws wins wherever the collision carries a division (its three independent chains per thread hide the reciprocal); the
// two-relaxation-time collisions have none and are faster with the sixteen warps per SM of k_lbmn_bulk on large grids.
// PLBM_TRIPLE_WS: 0 = always k_lbmn_bulk, 1 = always k_lbm3_ws, unset = that rule.
bool lbm_triple_ws_wanted(int model)
{
    static const int mode = env_knob("PLBM_TRIPLE_WS", -1);
    if (mode >= 0) return mode != 0;
    return model != M_TRT && model != M_TRT_SPLIT;
}

bool lbm_multi_shape_is_default()
{
    static const bool dflt = env_knob("PLBM_MULTI_WIDE", PLBM_MULTI_WIDE_DEFAULT) != 0 &&
                             (env_knob("PLBM_MULTI_NT", 0) == 0 || env_knob("PLBM_MULTI_NT", 0) == 128);
    return dflt;
}

// `nstep` (2 or 3) fused steps src -> dst for columns [x_begin, x_end).  halo_lo / halo_hi: the ring neighbours' nearest lines
// under a slab decomposition ([PLBM_HALO_LINES][9][ld]), nullptr = periodic self-wrap.
template <typename T>
int launch_lbm_multi(const Grid& g, const T* src, T* dst, int x_begin, int x_end, int model, const CollideParams<T>& cp, int nstep,
                     cudaStream_t s, const T* halo_lo, const T* halo_hi, T* dst_mid)
{
    if (!lbm_multi_applicable(g, model, nstep) || (dst_mid && !(nstep == 3 && lbm_multi_shape_is_default()))) {
        set_error("launch_lbm_multi: not applicable to this grid / collision / depth");
        return PLBM_ERR_ARG;
    }
    // three steps, no halo lines to read: the warp-specialised, skewed form of the kernel (plbm_lbm3w.cu)
    if (nstep == 3 && !halo_lo && !halo_hi && lbm_triple_ws_wanted(model)) return launch_lbm_triple_ws<T>(g, src, dst, dst_mid, x_begin, x_end, model, cp, s);
    // the launch that closes a call: the state after step 2 goes to dst_mid as well
    if (dst_mid) return dispatch_n<T, 3, 128, true, true>(g, src, dst, x_begin, x_end, halo_lo, halo_hi, model, cp, s, dst_mid);
#define PLBM_N(NS, NT, WD) return dispatch_n<T, NS, NT, WD>(g, src, dst, x_begin, x_end, halo_lo, halo_hi, model, cp, s)
    static const int wide2 = env_knob("PLBM_MULTI_WIDE2", 0);  // two steps per pass in the one-row shape (measurement, variant 9)
    if (nstep == 2 && wide2) PLBM_N(2, 128, true);
    if (nstep == 2) PLBM_N(2, 128, false);
    static const int nt = env_knob("PLBM_MULTI_NT", 0);  // 0: the default of the shape (128 threads)
    static const int wide = env_knob("PLBM_MULTI_WIDE", PLBM_MULTI_WIDE_DEFAULT);
    // small blocks (measurement knobs): the barriers of a level couple two warps / one warp instead of eight
    // (measured and dropped, r02g: one row per thread with 64- / 32-thread blocks 85.6 / 82.8 GLUPS, 16 bytes per thread with 64- /
    // 32-thread blocks 86.1 / 75.1, against 97.9 for the default below -- BGK fp64 8192^2)
    if (wide && nt == 64) PLBM_N(3, 64, true);
    if (wide && nt == 256) PLBM_N(3, 256, true);
    if (wide) PLBM_N(3, 128, true);  // the default: four blocks of four warps per SM
    if (nt == 256) PLBM_N(3, 256, false);
    PLBM_N(3, 128, false);
#undef PLBM_N
}

template int launch_lbm_multi<double>(const Grid&, const double*, double*, int, int, int, const CollideParams<double>&, int, cudaStream_t,
                                      const double*, const double*, double*);
template int launch_lbm_multi<float>(const Grid&, const float*, float*, int, int, int, const CollideParams<float>&, int, cudaStream_t,
                                     const float*, const float*, float*);

}  // namespace plbm
