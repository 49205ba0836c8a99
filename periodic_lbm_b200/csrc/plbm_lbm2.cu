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
// column), 27 column slots in all = 54 KB per 128-thread block, four blocks per SM.  The
// operands of phase A of the NEXT column are loaded before phase B starts, so the HBM latency
// hides behind phase B.  Arithmetic per node is collide<T,MODEL> on the same operands as two
// k_lbm launches, so the result is bit-identical.
//   lbm_stream_kernel  src/periodic_lbm.f90:45-127 ;  collisions src/collision_*.F90
//
// Measured alternatives (B200, BGK fp64 8192^2, this kernel 64 GLUPS): threads striding over rows
// so that every access is a contiguous run (no bank conflicts, +50 % memory instructions) 55.6;
// two barriers per column with 18 ring slots and six blocks per SM but no prefetch 56.2; 64- or
// 256-thread blocks 58.7 / 58.0; streaming-store hints, shared-memory carve-out and 32-column
// segments: within noise; skewed row ownership (phase A owns rows (odd, even), phase B (even, odd), so
// that the shifted populations become aligned vectors: 42 instead of 48 memory instructions per thread
// and column) 63.7, i.e. no change; issuing the prefetch before the barrier 62.4; fp32 shifted
// populations as aligned vector + one element (15 instead of 27 loads) 111.6 vs 122; one row per thread
// with 256-thread blocks = 32 warps per SM at 64 registers 58.1 (59.6 without the prefetch); adjacent
// strips co-scheduled as clusters of 2/4/8 61.2/59.9/59.6; non-power-of-two grids: no change.  With the
// collision REMOVED the kernel runs 66 (fp32: 126 vs 122): it is bound by the global -> register -> shared
// -> register -> global data path at ~4.7 TB/s of HBM traffic, not by arithmetic, occupancy, instruction
// count or DRAM locality (ncu: stall_mio_throttle dominates, HBM 58 % busy).  Legs of that path at 8192^2,
// per pair launch: ring only (no global loads, no stores) 0.47 ms; loads + ring 1.14 ms (4.2 TB/s of
// reads); ring + stores 0.89 ms (5.5 TB/s of writes); everything 2.08 ms ~ the SUM of the two legs: loads
// and stores do not overlap.  More bytes in flight do not help either: two columns of operands in two
// register sets (168 registers, three blocks per SM, 110 KB in flight) 63.0.  Every structure tried moves
// ~4.7 TB/s of HBM traffic; what caps the mixed load/store stream of this kernel below the 6.5 TB/s of a
// plain copy was the open question -- answered by k_lbm2_bulk below: with the raw columns delivered by bulk
// async copies (cp.async.bulk + mbarrier, issued by one thread two columns ahead) no thread issues an LDG,
// the memory-instruction queue only carries the ring traffic and the result stores (stall_mio_throttle 3.5 ->
// 0.28 per issue), and the kernel moves 6.2-6.4 TB/s of DRAM traffic = 0.95-0.97 of the measured copy peak:
// BGK / TRT / RR fp64 83.0 / 82.7 / 68.5 GLUPS at 8192^2 (k_lbm2: 63.9 / 64.1 / 58.5), fp32 BGK 147 (119).
// k_lbm2_bulk was round 1's default; since round 2 perform_lbm_step advances THREE steps per pass (k_lbmn_bulk, plbm_lbmn.cu) and
// the kernels of this file serve the pair(s) that may close a call, grids below 512^2 nodes, PLBM_TRIPLES=0, the launches of a pair
// that read a neighbour's halo lines (k_lbm2<HALO>) and tiny ny.
#include <climits>
#include <cstdint>
#include <cstdlib>

#include "plbm_internal.h"

// plbm_lbm2_fma.cu compiles this file a second time with -fmad=true: the same kernels with their multiply-adds
// contracted, exported as launch_lbm_pair_fma (opt-in variant 11, see that file).  The kernels live in an anonymous
// namespace and device code is not linked across translation units, so the two builds do not meet.
#ifdef PLBM_FMA_BUILD
#define launch_lbm_pair launch_lbm_pair_fma
#endif

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
    int x_begin, x_end;  // columns whose step-2 state this launch writes
    // two column ranges in one launch (both boundaries of a slab): segments that start at or beyond x_split are shifted by
    // x_skip columns, i.e. the launch covers [x_begin, x_split) and [x_split + x_skip, x_end); no split: x_split = INT_MAX
    int x_split, x_skip;
    int ty;              // interior rows per strip (multiple of V)
    int nstrips;         // strips along y
    int seglen;          // columns per x segment
    // slab decomposition: all nine populations of the neighbours' two nearest lines, [2][9][ld]
    // (lo: lines -2, -1; hi: lines nx, nx+1).  nullptr = periodic self-wrap.
    const T* halo_lo;
    const T* halo_hi;
    CollideParams<T> cp;
};

// row 0 of population Q in column `col` (col in [-2, nx+1])
template <typename T, bool HALO, int Q> __device__ __forceinline__ const T* column_of(const Lbm2Args<T>& a, int col)
{
    if (HALO) {
        if (col < 0) return a.halo_lo + ((size_t)(col + 2) * 9 + Q) * (size_t)a.ld;
        if (col >= a.nx) return a.halo_hi + ((size_t)(col - a.nx) * 9 + Q) * (size_t)a.ld;
    } else {
        col = col < 0 ? col + a.nx : (col >= a.nx ? col - a.nx : col);
    }
    return a.src + ((size_t)Q * a.nx + col) * (size_t)a.ld;
}

// streamed (pre-collision) populations of column xl, rows yp..yp+V-1
template <typename T, int V, bool HALO, int Q>
__device__ __forceinline__ void pull_global(const Lbm2Args<T>& a, int xl, int yp, T (&f)[V][9])
{
    constexpr int cy = cyi(Q);
    const T* line = column_of<T, HALO, Q>(a, xl - cxi(Q));
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

template <typename T, int V, bool HALO> __device__ __forceinline__ void load_column(const Lbm2Args<T>& a, int xl, int yp, T (&f)[V][9])
{
    pull_global<T, V, HALO, 0>(a, xl, yp, f);
    pull_global<T, V, HALO, 1>(a, xl, yp, f);
    pull_global<T, V, HALO, 2>(a, xl, yp, f);
    pull_global<T, V, HALO, 3>(a, xl, yp, f);
    pull_global<T, V, HALO, 4>(a, xl, yp, f);
    pull_global<T, V, HALO, 5>(a, xl, yp, f);
    pull_global<T, V, HALO, 6>(a, xl, yp, f);
    pull_global<T, V, HALO, 7>(a, xl, yp, f);
    pull_global<T, V, HALO, 8>(a, xl, yp, f);
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

template <typename T, int MODEL, int V, int NT, int MINB, bool HALO, bool PK>
__global__ void __launch_bounds__(NT, MINB) k_lbm2(const Lbm2Args<T> a)
{
    constexpr int W = NT * V;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* ring = reinterpret_cast<T*>(smem_raw);

    const int strip = blockIdx.x % a.nstrips, seg = blockIdx.x / a.nstrips;
    const int y_lo = strip * a.ty;
    const int y_hi = min(y_lo + a.ty, a.ny);
    int xs = a.x_begin + seg * a.seglen;
    int xe = min(xs + a.seglen, min(a.x_split, a.x_end));
    if (xs >= a.x_split) {  // second range of a split launch (the launcher makes no segment straddle the split)
        xs += a.x_skip;
        xe = min(xs + a.seglen, a.x_end);
    }
    const int t = threadIdx.x;
    const int yl = y_lo - V + t * V;                                    // logical first row of this thread
    const bool act_a = yl < y_hi + V;                                   // strip + V halo rows on both sides
    const bool act_b = yl >= y_lo && yl < y_hi;                         // strip interior
    const int yp = yl < 0 ? yl + a.ny : (yl >= a.ny ? yl - a.ny : yl);  // ny % V == 0: a vector never straddles the wrap

    RingPos rp = {0, 0, 0};
    T n[V][9];
    // warm-up: step-1 state of columns xs-1 and xs
    if (act_a) {
        load_column<T, V, HALO>(a, xs - 1, yp, n);
        collide_nodes<T, MODEL, V, PK>(n, a.cp);
        park_column<T, V, W>(ring, rp, t, n);
    }
    rp.advance();
    if (act_a) {
        load_column<T, V, HALO>(a, xs, yp, n);
        collide_nodes<T, MODEL, V, PK>(n, a.cp);
        park_column<T, V, W>(ring, rp, t, n);
        load_column<T, V, HALO>(a, xs + 1, yp, n);
    }
    rp.advance();

    for (int x = xs; x < xe; ++x) {
        // A: step-1 state of column x+1 (operands were loaded one iteration ago)
        if (act_a) {
            collide_nodes<T, MODEL, V, PK>(n, a.cp);
            park_column<T, V, W>(ring, rp, t, n);
        }
        __syncthreads();
        if (act_a && x + 1 < xe) load_column<T, V, HALO>(a, x + 2, yp, n);  // in flight during phase B
        // B: step-2 state of column x
        if (act_b) {
            T f[V][9];
            pull_ring<T, V, W>(ring, rp, t, f);
            collide_nodes<T, MODEL, V, PK>(f, a.cp);
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

// ---- k_lbm2_bulk: raw columns land in shared memory by bulk async copies, two columns ahead -----------------
// Ring of 18 slots (a population is kept 1 / 2 / 3 columns for cx = -1 / 0 / +1; two barriers per column) + two
// stages of nine raw population lines, each guarded by one mbarrier: thread 0 arms the barrier with the byte
// count of the column (arrive.expect_tx) and issues one cp.async.bulk per population line -- up to three pieces
// where the staged rows [y_lo - 2V, y_hi + 2V) wrap around y; every piece starts and ends on a 16-byte boundary
// because y_lo, y_hi, ny and ld are multiples of V.  Column k of a segment uses stage k & 1 and completes phase
// (k >> 1) & 1 of its barrier; the refill of a stage is issued right after the barrier that ends its last read
// (fence.proxy.async orders those generic reads before the async writes).  No per-thread LDG.
__host__ __device__ constexpr int ringb_depth(int q) { return cxi(q) == -1 ? 1 : (cxi(q) == 0 ? 2 : 3); }
__host__ __device__ constexpr int ringb_base(int q)
{
    return q == 3 ? 0 : q == 6 ? 1 : q == 7 ? 2 : q == 0 ? 3 : q == 2 ? 5 : q == 4 ? 7 : q == 1 ? 9 : q == 5 ? 12 : 15;
}
constexpr int RINGB_SLOTS = 18;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
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
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
                 : "memory");
}

template <typename T, int MODEL, int V, int NT, int MINB, bool PK>
__global__ void __launch_bounds__(NT, MINB) k_lbm2_bulk(const Lbm2Args<T> a)
{
    constexpr int W = NT * V;       // rows of one ring column: logical rows y_lo - V .. y_lo - V + W - 1
    constexpr int WS = W + 2 * V;   // rows of one staged raw column: logical rows y_lo - 2V .. (one extra vector each side)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* ring = reinterpret_cast<T*>(smem_raw);
    T* stage = ring + RINGB_SLOTS * W;  // [2][9][WS]
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage + 2 * 9 * WS);
    const uint32_t bar0 = smem_u32(bars);
    auto bar = [&](int s) { return bar0 + 8u * (uint32_t)s; };  // one mbarrier per stage

    const int strip = blockIdx.x % a.nstrips, seg = blockIdx.x / a.nstrips;
    const int y_lo = strip * a.ty;
    const int y_hi = min(y_lo + a.ty, a.ny);
    const int xs = a.x_begin + seg * a.seglen;
    const int xe = min(xs + a.seglen, a.x_end);
    const int t = threadIdx.x;
    const int yl = y_lo - V + t * V;
    const bool act_a = yl < y_hi + V;
    const bool act_b = yl >= y_lo && yl < y_hi;
    const int yp = yl < 0 ? yl + a.ny : (yl >= a.ny ? yl - a.ny : yl);
    // staged logical rows [r0, r1): what phase A reads, rounded to whole vectors
    const int r0 = y_lo - 2 * V, r1 = y_hi + 2 * V;

    if (t == 0) {
        mbar_init(bar(0), 1);
        mbar_init(bar(1), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // one raw column (all nine populations, pulled: population q from column xl - cx_q) into stage s
    auto issue = [&](int xl, int s) {
        mbar_expect_tx(bar(s), (uint32_t)(9 * (r1 - r0) * sizeof(T)));
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            int col = xl - cxi(q);
            col = col < 0 ? col + a.nx : (col >= a.nx ? col - a.nx : col);
            const T* line = a.src + ((size_t)q * a.nx + col) * (size_t)a.ld;
            T* dst = stage + (s * 9 + q) * WS;
            // periodic pieces of [r0, r1): below 0, inside, beyond ny (all multiples of V rows = 16 bytes)
            if (r0 < 0) bulk_load(smem_u32(dst), line + (a.ny + r0), (uint32_t)(-r0 * sizeof(T)), bar(s));
            const int m0 = max(r0, 0), m1 = min(r1, a.ny);
            bulk_load(smem_u32(dst + (m0 - r0)), line + m0, (uint32_t)((m1 - m0) * sizeof(T)), bar(s));
            if (r1 > a.ny) bulk_load(smem_u32(dst + (a.ny - r0)), line, (uint32_t)((r1 - a.ny) * sizeof(T)), bar(s));
        }
    };
    int w2 = 0, w3 = 0;  // ring slot of the column being written (depth 2 / depth 3 groups)
    auto phase_a = [&](int s, uint32_t parity) {
        mbar_wait(bar(s), parity);
        if (act_a) {
            T n[V][9];
            const int i = V + t * V;  // stage row of this thread's first row
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                const T* col = stage + (s * 9 + q) * WS + i;
                const int cy = cyi(q);
                if (cy == 0) {
                    const Vec<T, V> p = *reinterpret_cast<const Vec<T, V>*>(col);
#pragma unroll
                    for (int v = 0; v < V; ++v) n[v][q] = p.v[v];
                } else {
#pragma unroll
                    for (int v = 0; v < V; ++v) n[v][q] = col[v - cy];
                }
            }
            collide_nodes<T, MODEL, V, PK>(n, a.cp);
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                const int slot = ringb_depth(q) == 1 ? 0 : (ringb_depth(q) == 2 ? w2 : w3);
                Vec<T, V> p;
#pragma unroll
                for (int v = 0; v < V; ++v) p.v[v] = n[v][q];
                *reinterpret_cast<Vec<T, V>*>(ring + (ringb_base(q) + slot) * W + t * V) = p;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // stage reads before the async refill
    };
    auto advance = [&]() {
        w2 ^= 1;
        w3 = w3 == 2 ? 0 : w3 + 1;
    };

    // columns are numbered k = xl - (xs - 1); column k uses stage k & 1 with parity (k >> 1) & 1
    if (t == 0) {
        issue(xs - 1, 0);
        issue(xs, 1);
    }
    phase_a(0, 0);  // A(xs - 1)
    advance();
    __syncthreads();
    if (t == 0) issue(xs + 1, 0);
    phase_a(1, 0);  // A(xs)
    advance();
    __syncthreads();
    if (t == 0 && xs + 2 <= xe) issue(xs + 2, 1);

    for (int x = xs; x < xe; ++x) {
        const int k = x + 1 - (xs - 1);  // column x + 1
        phase_a(k & 1, (uint32_t)((k >> 1) & 1));
        __syncthreads();
        if (t == 0 && x + 3 <= xe) issue(x + 3, k & 1);  // lands during phase B of this and A/B of the next column
        if (act_b) {
            T f[V][9];
            const int r2 = w2 ^ 1;                // column x
            const int r3 = w3 == 2 ? 0 : w3 + 1;  // column x - 1
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                const int slot = ringb_depth(q) == 1 ? 0 : (ringb_depth(q) == 2 ? r2 : r3);
                const T* col = ring + (ringb_base(q) + slot) * W + t * V;
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
            collide_nodes<T, MODEL, V, PK>(f, a.cp);
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                Vec<T, V> p;
#pragma unroll
                for (int v = 0; v < V; ++v) p.v[v] = f[v][q];
                *reinterpret_cast<Vec<T, V>*>(a.dst + ((size_t)q * a.nx + x) * (size_t)a.ld + yp) = p;
            }
        }
        advance();
        __syncthreads();
    }
}

int env_int(const char* name, int dflt);

// Packed fp32 collisions (two nodes per FFMA2, plbm_f32x2.cuh; bit-identical to the scalar ones): the default for fp32
// (PLBM_F32_PACKED=0/1 overrides); variant 12 selects the other one, for A/B measurements.  Never in the FMA build.
#ifndef PLBM_F32_PACKED_DEFAULT
#define PLBM_F32_PACKED_DEFAULT 1
#endif
#ifndef PLBM_BULK_WIDE_DEFAULT
#define PLBM_BULK_WIDE_DEFAULT 0
#endif
#ifndef PLBM_BULK_NT_DEFAULT
#define PLBM_BULK_NT_DEFAULT 128
#endif
template <typename T> bool packed_collisions(const Grid& g)
{
#ifdef PLBM_FMA_BUILD
    return false;
#else
    if (sizeof(T) != 4) return false;
    static const int dflt = env_int("PLBM_F32_PACKED", PLBM_F32_PACKED_DEFAULT);
    return g.variant == 12 ? !dflt : (dflt != 0);
#endif
}
#ifdef PLBM_FMA_BUILD
template <typename T> struct can_pack {
    static constexpr bool value = false;
};
#else
template <typename T> struct can_pack {
    static constexpr bool value = sizeof(T) == 4;
};
#endif

// V rows per thread, NT threads: (16 bytes, 128) or -- "wide", twice the warps on the same shared memory -- (8 bytes, 256)
template <typename T, int MODEL, bool PK, int V, int NT>
int launch_pair_bulk_t(const Grid& g, const T* src, T* dst, int x_begin, int x_end, const CollideParams<T>& cp, cudaStream_t s)
{
    constexpr int VA = 16 / (int)sizeof(T);  // strips begin and end on 16-byte boundaries (bulk copies)
    constexpr int MINB = NT >= 128 ? 3 : (NT == 64 ? 6 : 12);  // what the shared memory of an SM holds
    constexpr int W = NT * V, WS = W + 2 * V;
    constexpr size_t smem = ((size_t)RINGB_SLOTS * W + 2 * 9 * WS) * sizeof(T) + 16;
    if (x_end <= x_begin) return PLBM_OK;
    auto kern = k_lbm2_bulk<T, MODEL, V, NT, MINB, PK>;
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
    a.x_begin = x_begin;
    a.x_end = x_end;
    a.x_split = INT_MAX;
    a.x_skip = 0;
    a.halo_lo = a.halo_hi = nullptr;
    a.cp = cp;
    // Strips x segments.  All blocks take about the same time and an SM holds MINB of them, so the launch finishes in
    // ceil(blocks / slots) rounds, slots = MINB x SMs: a last, mostly empty round costs as much as a full one.  Measured on
    // B200 with 64-column segments throughout (fp64, GLUPS): 8192^2 = 9.5 rounds 83; 4096^2 = 2.45 rounds 77; 2048^2 = 0.65
    // round 63; 1024^2 = 0.18 round 21 (k_lbm2: 49) -- and 64 at 1024^2 once the segments are cut to fill the round.
    // Always the fewest strips (a block costs the same per column whether its 128 threads own 252 rows or fewer: measured,
    // 1024^2 fp64 with 8 strips of 128 rows 51, with 5 strips of 206 rows 64), and the segment count that makes the blocks
    // fill a whole number of rounds, from below; at least 8 columns per segment (two warm-up columns each).
    const int ncols = x_end - x_begin;
    const int ty_max = (NT - 2) * V;
    const int slots = MINB * g.sm_count;
    const int nstrips = (g.ny + ty_max - 1) / ty_max;
    static const int fill = env_int("PLBM_PAIR_BULK_FILL", 1);  // 0: 64-column segments always (the round-1 launcher, for A/B)
    int nseg = (ncols + 63) / 64;
    if (fill) {
        const long long blocks64 = (long long)nstrips * nseg;
        const long long rounds = blocks64 >= slots ? (blocks64 + slots - 1) / slots : 1;
        nseg = (int)(rounds * slots / nstrips);
    }
    a.ty = ((g.ny + nstrips - 1) / nstrips + VA - 1) / VA * VA;
    a.nstrips = (g.ny + a.ty - 1) / a.ty;
    if (nseg < 1) nseg = 1;
    a.seglen = (ncols + nseg - 1) / nseg;
    if (a.seglen < 8) a.seglen = ncols < 8 ? ncols : 8;
    nseg = (ncols + a.seglen - 1) / a.seglen;
    kern<<<(unsigned)(a.nstrips * nseg), NT, smem, s>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PLBM_CUDA(cudaGetLastError());
    return PLBM_OK;
}

template <typename T, int MODEL>
int launch_pair_bulk(const Grid& g, const T* src, T* dst, int x_begin, int x_end, const CollideParams<T>& cp, cudaStream_t s)
{
    constexpr int V = 16 / (int)sizeof(T);
    // block shape (measurement knobs): PLBM_BULK_WIDE=1: 8 bytes per thread, 256 threads; PLBM_BULK_NT=64 / 32: small blocks, six /
    // twelve per SM, whose barriers couple fewer warps
    static const int wide = env_int("PLBM_BULK_WIDE", PLBM_BULK_WIDE_DEFAULT);
    static const int nt = env_int("PLBM_BULK_NT", PLBM_BULK_NT_DEFAULT);
    const bool pk = packed_collisions<T>(g);
#define PLBM_BULK_SHAPE(VV, NN)                                                                                              \
    return pk ? launch_pair_bulk_t<T, MODEL, can_pack<T>::value, VV, NN>(g, src, dst, x_begin, x_end, cp, s)                  \
              : launch_pair_bulk_t<T, MODEL, false, VV, NN>(g, src, dst, x_begin, x_end, cp, s)
    if (wide) PLBM_BULK_SHAPE(V / 2, 256);
    if (nt == 64) PLBM_BULK_SHAPE(V, 64);
    if (nt == 32) PLBM_BULK_SHAPE(V, 32);
    PLBM_BULK_SHAPE(V, 128);
#undef PLBM_BULK_SHAPE
}

template <typename T>
int dispatch_pair_bulk(const Grid& g, const T* src, T* dst, int x_begin, int x_end, int model, const CollideParams<T>& cp, cudaStream_t s)
{
    switch (model) {
    case M_BGK: return launch_pair_bulk<T, M_BGK>(g, src, dst, x_begin, x_end, cp, s);
    case M_TRT: return launch_pair_bulk<T, M_TRT>(g, src, dst, x_begin, x_end, cp, s);
    case M_RR: return launch_pair_bulk<T, M_RR>(g, src, dst, x_begin, x_end, cp, s);
    case M_BGK_SPLIT: return launch_pair_bulk<T, M_BGK_SPLIT>(g, src, dst, x_begin, x_end, cp, s);
    case M_TRT_SPLIT: return launch_pair_bulk<T, M_TRT_SPLIT>(g, src, dst, x_begin, x_end, cp, s);
    case M_BGK_IMPROVED: return launch_pair_bulk<T, M_BGK_IMPROVED>(g, src, dst, x_begin, x_end, cp, s);
    }
    set_error("launch_lbm_pair: unknown collision model");
    return PLBM_ERR_ARG;
}

// default flavour of the two-step kernel where both apply: 1 = bulk async copies (k_lbm2_bulk), 0 = per-thread loads
// (measured on B200: bulk copies 83.0 / 82.7 / 68.5 GLUPS for BGK / TRT / RR fp64 at 8192^2, per-thread loads 63.9 / 64.1 / 58.5)
#ifndef PLBM_PAIR_BULK_DEFAULT
#define PLBM_PAIR_BULK_DEFAULT 1
#endif

int env_int(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

template <typename T, int MODEL, bool HALO, bool PK>
int launch_pair_t(const Grid& g, const T* src, T* dst, int x_begin, int x_end, const T* halo_lo, const T* halo_hi, const CollideParams<T>& cp,
                  cudaStream_t s, int nb_split)
{
    constexpr int V = 16 / (int)sizeof(T);
    constexpr int NT = 128, MINB = 4;
    constexpr int W = NT * V;
    constexpr size_t smem = (size_t)RING_SLOTS * W * sizeof(T);
    if (x_end <= x_begin) return PLBM_OK;
    auto kern = k_lbm2<T, MODEL, V, NT, MINB, HALO, PK>;
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
    a.x_begin = x_begin;
    a.x_end = x_end;
    a.halo_lo = halo_lo;
    a.halo_hi = halo_hi;
    a.cp = cp;
    a.x_split = INT_MAX;
    a.x_skip = 0;
    if (nb_split > 0) {  // both boundaries of a slab: columns [0, nb) and [nx - nb, nx), one segment each
        a.x_begin = 0;
        a.x_end = g.nx;
        a.x_split = nb_split;
        a.x_skip = g.nx - 2 * nb_split;
    }
    const int ncols = nb_split > 0 ? 2 * nb_split : x_end - x_begin;
    const int ty_max = (NT - 2) * V;
    const int slots = g.sm_count * MINB;  // resident blocks of one wave
    // Strips: among the counts from the minimum upwards take the one whose busiest block has the least work
    // when strips x segments fill one wave (a narrower strip costs 2V/ty redundant rows in phase A).
    const int nstrips_min = (g.ny + ty_max - 1) / ty_max;
    int best = nstrips_min;
    double best_cost = 1e300;
    for (int ns = nstrips_min; ns <= nstrips_min + 8; ++ns) {
        const int ty = ((g.ny + ns - 1) / ns + V - 1) / V * V;
        int nseg = slots / ns;
        if (nseg < 1) nseg = 1;
        const int seglen = (ncols + nseg - 1) / nseg;
        const double cost = (double)(ty + 2 * V) * (seglen + 2);
        if (cost < best_cost) {
            best_cost = cost;
            best = ns;
        }
    }
    a.nstrips = env_int("PLBM_PAIR_NSTRIPS", best);
    if (a.nstrips < nstrips_min) a.nstrips = nstrips_min;
    a.ty = ((g.ny + a.nstrips - 1) / a.nstrips + V - 1) / V * V;
    a.nstrips = (g.ny + a.ty - 1) / a.ty;
    // Segments: at least one wave, and short enough (64 columns, 3 % warm-up) that the block scheduler
    // evens out the SMs over several waves (measured +5 % over one wave of long segments).
    int nseg = slots / a.nstrips;
    if (nseg < (ncols + 63) / 64) nseg = (ncols + 63) / 64;
    nseg = env_int("PLBM_PAIR_NSEG", nseg);
    if (nseg < 1) nseg = 1;
    a.seglen = (ncols + nseg - 1) / nseg;
    if (a.seglen < 8) a.seglen = ncols < 8 ? ncols : 8;
    nseg = (ncols + a.seglen - 1) / a.seglen;
    if (nb_split > 0) {
        a.seglen = nb_split;
        nseg = 2;
    }
    kern<<<(unsigned)(a.nstrips * nseg), NT, smem, s>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PLBM_CUDA(cudaGetLastError());
    return PLBM_OK;
}

template <typename T, int MODEL, bool HALO>
int launch_pair(const Grid& g, const T* src, T* dst, int x_begin, int x_end, const T* halo_lo, const T* halo_hi, const CollideParams<T>& cp,
                cudaStream_t s, int nb_split = 0)
{
    if (packed_collisions<T>(g)) return launch_pair_t<T, MODEL, HALO, can_pack<T>::value>(g, src, dst, x_begin, x_end, halo_lo, halo_hi, cp, s, nb_split);
    return launch_pair_t<T, MODEL, HALO, false>(g, src, dst, x_begin, x_end, halo_lo, halo_hi, cp, s, nb_split);
}

template <typename T, bool HALO>
int dispatch_pair(const Grid& g, const T* src, T* dst, int x_begin, int x_end, const T* halo_lo, const T* halo_hi, int model,
                  const CollideParams<T>& cp, cudaStream_t s, int nb_split = 0)
{
    switch (model) {
    case M_BGK: return launch_pair<T, M_BGK, HALO>(g, src, dst, x_begin, x_end, halo_lo, halo_hi, cp, s, nb_split);
    case M_TRT: return launch_pair<T, M_TRT, HALO>(g, src, dst, x_begin, x_end, halo_lo, halo_hi, cp, s, nb_split);
    case M_RR: return launch_pair<T, M_RR, HALO>(g, src, dst, x_begin, x_end, halo_lo, halo_hi, cp, s, nb_split);
    case M_BGK_SPLIT: return launch_pair<T, M_BGK_SPLIT, HALO>(g, src, dst, x_begin, x_end, halo_lo, halo_hi, cp, s, nb_split);
    case M_TRT_SPLIT: return launch_pair<T, M_TRT_SPLIT, HALO>(g, src, dst, x_begin, x_end, halo_lo, halo_hi, cp, s, nb_split);
    case M_BGK_IMPROVED: return launch_pair<T, M_BGK_IMPROVED, HALO>(g, src, dst, x_begin, x_end, halo_lo, halo_hi, cp, s, nb_split);
    }
    set_error("launch_lbm_pair: unknown collision model");
    return PLBM_ERR_ARG;
}

}  // namespace

#ifndef PLBM_FMA_BUILD
bool lbm_pair_applicable(const Grid& g)
{
    const int v = 16 / (int)g.esize();
    return g.nx >= 4 && g.ny >= 2 * v && (g.ny % v) == 0;
}

// Which kernel advances a pair of steps on this grid: 0 none (one step per launch), 1 k_lbm2 (raw columns by
// per-thread loads), 2 k_lbm2_bulk (raw columns by bulk async copies).
int lbm_pair_flavour(const Grid& g)
{
    if (!lbm_pair_variant(g.variant) || !lbm_pair_applicable(g)) return 0;
    const int v = 16 / (int)g.esize();
    if (g.ny < 8 * v) return 1;
    static const int bulk_default = env_int("PLBM_PAIR_BULK", PLBM_PAIR_BULK_DEFAULT);
    if (g.variant == 6) return 1;
    if (g.variant == 7 || g.variant == 8) return 2;
    if (bulk_default == 0) return 1;
    if (bulk_default >= 2) return 2;  // PLBM_PAIR_BULK=2: on every grid (A/B measurements)
    // k_lbm2_bulk wherever its launcher can fill one round of blocks (three per SM) with full-height strips and segments of
    // at least 8 columns.  Measured (GLUPS, k_lbm2 -> k_lbm2_bulk): fp64 1024^2 TRT 46 -> 64, 2048^2 54 -> 73, 4096^2 58 -> 79,
    // 8192^2 64 -> 83; fp32 4096^2 96 -> 129, 8192^2 119 -> 147.  Smaller grids stay on k_lbm2, whose launcher trades strips for
    // segments down to a few hundred nodes per block (fp64 768^2: 41 vs 26; fp32 1024^2: 68 vs 54).
    const long long strips = (g.ny + 126 * v - 1) / (126 * v), segs = (g.nx + 7) / 8;
    return strips * segs >= 3LL * g.sm_count ? 2 : 1;
}

#endif  // !PLBM_FMA_BUILD

// Two fused steps src -> dst for columns [x_begin, x_end).  halo_lo / halo_hi: the ring neighbours' two
// nearest lines ([2][9][ld]) under a slab decomposition, nullptr = periodic self-wrap.  The caller
// accounts for the lattice roles (see step_lbm_t).
template <typename T>
int launch_lbm_pair(const Grid& g, const T* src, T* dst, int x_begin, int x_end, const T* halo_lo, const T* halo_hi, int model,
                    const CollideParams<T>& cp, cudaStream_t s)
{
    // The launches that read the neighbours' halo lines (two boundary lines per side of a slab) stay on k_lbm2.
    if (!halo_lo && lbm_pair_flavour(g) == 2) return dispatch_pair_bulk<T>(g, src, dst, x_begin, x_end, model, cp, s);
    if (halo_lo && halo_hi) return dispatch_pair<T, true>(g, src, dst, x_begin, x_end, halo_lo, halo_hi, model, cp, s);
    return dispatch_pair<T, false>(g, src, dst, x_begin, x_end, nullptr, nullptr, model, cp, s);
}

#ifndef PLBM_FMA_BUILD
// Both boundaries of a slab in ONE launch: two fused steps for columns [0, nb) and [nx - nb, nx), reading the ring
// neighbours' halo lines (nullptr: periodic self-wrap, the single-GPU emulation of the slab schedule, variant 8).
template <typename T>
int launch_lbm_pair_boundaries(const Grid& g, const T* src, T* dst, int nb, const T* halo_lo, const T* halo_hi, int model,
                               const CollideParams<T>& cp, cudaStream_t s)
{
    if (halo_lo && halo_hi) return dispatch_pair<T, true>(g, src, dst, 0, g.nx, halo_lo, halo_hi, model, cp, s, nb);
    return dispatch_pair<T, false>(g, src, dst, 0, g.nx, nullptr, nullptr, model, cp, s, nb);
}
template int launch_lbm_pair_boundaries<double>(const Grid&, const double*, double*, int, const double*, const double*, int,
                                                const CollideParams<double>&, cudaStream_t);
template int launch_lbm_pair_boundaries<float>(const Grid&, const float*, float*, int, const float*, const float*, int,
                                               const CollideParams<float>&, cudaStream_t);
#endif

template int launch_lbm_pair<double>(const Grid&, const double*, double*, int, int, const double*, const double*, int,
                                     const CollideParams<double>&, cudaStream_t);
template int launch_lbm_pair<float>(const Grid&, const float*, float*, int, int, const float*, const float*, int,
                                    const CollideParams<float>&, cudaStream_t);

}  // namespace plbm
