// plbm_lbm3w.cu -- three fused stream+collide steps per pass over HBM, warp-specialised and skewed (sm_100a).
//
// Same arithmetic, same operands and therefore the same bits as three k_lbm launches (and as k_lbmn_bulk<NSTEP = 3>, plbm_lbmn.cu,
// whose strips, segments, staged bulk copies and mbarriers it keeps); what changes is the schedule inside a block:
//   * k_lbmn_bulk runs its three levels one after the other on column x + 2, x + 1, x with a block barrier after each, because
//     level l reads what level l - 1 wrote in the same iteration.  ncu (profiles/r02_z_ncu_full_k_lbmn_bulk3_bgk_f64_c5.csv): issue
//     slots 65 % busy, fp64 pipe 61 %, DRAM 84 %; top stall `wait` (fixed-latency dependency, 1.6 per issue), then barrier.
//   * here the levels are skewed by two columns each (level 1 on column j, level 2 on j - 2, level 3 on j - 4), so every level reads
//     only what EARLIER iterations wrote: the three collisions of an iteration are independent instruction streams of one thread
//     (the dependency stalls of one fp64 chain hide behind the other two), and the consumers need ONE barrier per column.  The rings
//     keep a population one column longer for that (2 / 3 / 4 columns for cx = -1 / 0 / +1: 27 slots instead of 18), which costs one
//     block per SM (three of 73 KB instead of four of 56 KB).
//   * the copies are issued by a PRODUCER warp (population q by lane q from a running line address), released per stage by the
//     consumers through a named barrier they only arrive at; the four consumer warps never execute copy-issue code.
//   * the IEEE divisions of the three nodes are issued first (node_reciprocals, plbm_math.cuh): the branch to the division's slow
//     path would otherwise end the basic block between the collisions and keep the compiler from interleaving them.
// DUAL: the launch that closes a call also stores the state after step 2 (what lattice `inew` must hold, state n - 1) to a third
// buffer, so the call does not have to end with a single-step launch.
//   lbm_stream_kernel  src/periodic_lbm.f90:45-127 ;  collisions src/collision_*.F90
#include <cstdint>
#include <cstdlib>

#include "plbm_internal.h"

namespace plbm {

namespace {

template <typename T, int V> struct alignas(sizeof(T) * V) VecN {
    T v[V];
};

template <typename T> struct Lbm3Args {
    const T* src;
    T* dst;
    T* dst_mid;  // DUAL: the state after step 2
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
// named barriers: 1 / 2 = "stage 0 / 1 has been read" (consumers arrive, the producer warp waits), 3 = the consumers' column barrier
template <int ID> __device__ __forceinline__ void named_arrive(int count) { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "r"(count) : "memory"); }
template <int ID> __device__ __forceinline__ void named_sync(int count) { asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(count) : "memory"); }

// skewed rings: a population is kept 2 / 3 / 4 columns (cx = -1 / 0 / +1): 27 column slots per ring
__host__ __device__ constexpr int rs_depth(int q) { return cxi(q) == -1 ? 2 : (cxi(q) == 0 ? 3 : 4); }
__host__ __device__ constexpr int rs_base(int q)
{
    return q == 3 ? 0 : q == 6 ? 2 : q == 7 ? 4 : q == 0 ? 6 : q == 2 ? 9 : q == 4 ? 12 : q == 1 ? 15 : q == 5 ? 19 : 23;
}
constexpr int RS_SLOTS = 27;

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

// Three fused steps per pass, warp-specialised and skewed.  NTC consumer threads (one per V rows of the strip + halo) and ONE
// producer warp.  In iteration j (raw column j has landed in its stage):
//   level 1  collides the streamed raw column j                              -> ring 0, column j
//   level 2  pulls column j - 2 from ring 0 (columns j-3, j-2, j-1), collides -> ring 1, column j - 2
//   level 3  pulls column j - 4 from ring 1 (columns j-5, j-4, j-3), collides -> dst, column j - 4
// Every level reads only what EARLIER iterations wrote, so the three collisions of an iteration are independent instruction
// streams of one thread (the fixed-latency dependency stalls of one fp64 chain hide behind the other two) and the consumers
// need ONE barrier per column.  The producer warp owns the copies: population q by lane q from a running line address.
template <typename T, int MODEL, int V, int NTC, int MINB, bool DUAL>
__global__ void __launch_bounds__(NTC + 32, MINB) k_lbm3_ws(const Lbm3Args<T> a)
{
    constexpr bool PACKED = sizeof(T) == 4 && (V % 2) == 0 && MODEL != M_RR;
    constexpr int NR = 2;
    constexpr int NT = NTC + 32;
    constexpr int W = NTC * V;
    constexpr int VA = 16 / (int)sizeof(T);
    constexpr int HS = (NR * V + 1 + VA - 1) / VA * VA;
    constexpr int OFF = HS - NR * V;
    constexpr int WS = W + 2 * OFF;
    // a ring slot = V pad rows, the W rows of the threads, V pad rows: the levels run unpredicated on the halo threads as well, and
    // thread 0 / the last thread pull the row below / above theirs -- that row is a pad row nobody writes, not a row of the
    // neighbouring slot (which another thread may be writing in the same iteration: harmless, the halo thread's result is thrown
    // away, but compute-sanitizer racecheck reports it, r02r)
    constexpr int WP = W + 2 * V;
    extern __shared__ __align__(128) unsigned char smem_n[];
    T* stage = reinterpret_cast<T*>(smem_n);       // [2][9][WS]
    T* ring = stage + 2 * 9 * WS;                  // [2][RS_SLOTS][WP]
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + 2 * RS_SLOTS * WP);
    const uint32_t bar0 = smem_addr(bars);

    const int strip = blockIdx.x % a.nstrips, seg = blockIdx.x / a.nstrips;
    const int y_lo = strip * a.ty;
    const int y_hi = min(y_lo + a.ty, a.ny);
    const int xs = a.x_begin + seg * a.seglen;
    const int xe = min(xs + a.seglen, a.x_end);
    const int t = threadIdx.x;
    const int r0 = y_lo - HS, r1 = y_hi + HS;  // staged logical rows
    const int x_first = xs - 2;                // first raw column; the last one is xe + 1
    const int n_raw = xe + 2 - x_first;
    const int n_iter = n_raw + 2;

    // every shared-memory word a thread may read holds a finite positive number from the start: the levels run unpredicated on
    // the halo threads too (only their stores are predicated), and 1 / rho of an arbitrary bit pattern could take the slow path
    {
        constexpr int NWORDS = (2 * 9 * WS + 2 * RS_SLOTS * WP) / VA;
        static_assert((2 * 9 * WS + 2 * RS_SLOTS * WP) % VA == 0, "whole 16-byte words");
        VecN<T, VA> one;
#pragma unroll
        for (int v = 0; v < VA; ++v) one.v[v] = T(1);
        VecN<T, VA>* p = reinterpret_cast<VecN<T, VA>*>(smem_n);
        for (int i = t; i < NWORDS; i += NT) p[i] = one;
    }
    if (t == 0) {
        mbarrier_init(bar0, 1);
        mbarrier_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    if (t >= NTC) {
        // ---- producer warp: lane q owns population q
        const int q = t - NTC;
        const T* nxt = nullptr;
        int ncol = 0;
        uint32_t dsto = 0;
        const int m0 = max(r0, 0), m1 = min(r1, a.ny);
        const uint32_t p1_bytes = r0 < 0 ? (uint32_t)(-r0 * sizeof(T)) : 0u, p3_bytes = r1 > a.ny ? (uint32_t)((r1 - a.ny) * sizeof(T)) : 0u;
        const uint32_t p2_bytes = (uint32_t)((m1 - m0) * sizeof(T));
        const uint32_t p2_dst = (uint32_t)((m0 - r0) * sizeof(T)), p3_dst = (uint32_t)((a.ny - r0) * sizeof(T));
        if (q < 9) {
            int col = x_first - cxi(q);
            col = col < 0 ? col + a.nx : (col >= a.nx ? col - a.nx : col);
            ncol = col;
            nxt = a.src + ((size_t)q * a.nx + col) * (size_t)a.ld;
            dsto = smem_addr(stage + q * WS);
        }
        auto issue_next = [&](int s) {
            const uint32_t b = bar0 + 8u * (uint32_t)s;
            if (q == 0) mbarrier_expect_tx(b, (uint32_t)(9 * (r1 - r0) * sizeof(T)));
            if (q < 9) {
                const uint32_t d = dsto + (uint32_t)(s * 9 * WS * sizeof(T));
                if (p1_bytes) bulk_copy_g2s(d, nxt + (a.ny + r0), p1_bytes, b);
                bulk_copy_g2s(d + p2_dst, nxt + m0, p2_bytes, b);
                if (p3_bytes) bulk_copy_g2s(d + p3_dst, nxt, p3_bytes, b);
                if (++ncol == a.nx) {
                    ncol = 0;
                    nxt -= (size_t)(a.nx - 1) * (size_t)a.ld;
                } else {
                    nxt += a.ld;
                }
            }
        };
        issue_next(0);
        if (n_raw > 1) issue_next(1);
        for (int k = 0; k < n_iter; ++k) {
            if (k & 1) named_sync<2>(NT);  // the consumers have read stage k & 1
            else named_sync<1>(NT);
            if (k + 2 < n_raw) issue_next(k & 1);
        }
        return;
    }

    // ---- consumers
    const int yl = y_lo - NR * V + t * V;  // logical first row of this thread
    const bool a1 = yl < y_hi + 2 * V, a2 = yl >= y_lo - V && yl < y_hi + V, a3 = yl >= y_lo && yl < y_hi;
    const T* st0 = stage + OFF + t * V;
    T* rg0 = ring + V + t * V;
    T* rg1 = rg0 + RS_SLOTS * WP;
    const size_t qs = (size_t)a.nx * a.ld;
    T* out3 = a.dst + (size_t)(x_first - 4) * a.ld + yl;  // column j - 4 of population 0, this thread's first row
    T* out2 = DUAL ? a.dst_mid + (size_t)(x_first - 2) * a.ld + yl : nullptr;

    auto slot_w = [&](int q, int w2, int w3, int w4) { return rs_base(q) + (rs_depth(q) == 2 ? w2 : (rs_depth(q) == 3 ? w3 : w4)); };
    auto to_ring = [&](T* rg, const T (&n)[V][9], int w2, int w3, int w4) {
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            VecN<T, V> p;
#pragma unroll
            for (int v = 0; v < V; ++v) p.v[v] = n[v][q];
            *reinterpret_cast<VecN<T, V>*>(rg + slot_w(q, w2, w3, w4) * WP) = p;
        }
    };
    auto to_global = [&](T* o, const T (&n)[V][9]) {
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            VecN<T, V> p;
#pragma unroll
            for (int v = 0; v < V; ++v) p.v[v] = n[v][q];
            *reinterpret_cast<VecN<T, V>*>(o) = p;
            o += qs;
        }
    };

    // stage k & 1 has been read by this thread: the producer may refill it
    auto stage_read = [&](int k) {
        if (k & 1) named_arrive<2>(NT);
        else named_arrive<1>(NT);
    };
    int w2 = 0, w3 = 0, w4 = 0;
    for (int k = 0; k < n_iter; ++k) {
        const int j = x_first + k;
        const int r2 = w2 ^ 1, r3 = w3 == 2 ? 0 : w3 + 1, r4 = (w4 + 1) & 3;  // columns j-1 (depth 2), j-2 (depth 3), j-3 (depth 4)
        const bool l1 = k < n_raw, l2 = j >= xs + 1 && j <= xe + 2, l3 = j >= xs + 4;
        if (l1 && l2 && l3) {
            T n1[V][9], n2[V][9], n3[V][9];
            mbarrier_wait(bar0 + 8u * (uint32_t)(k & 1), (uint32_t)((k >> 1) & 1));
            {
                const T* st = st0 + (k & 1) * 9 * WS;
                pull_rows<T, V>([&](int q) { return st + q * WS; }, n1);
            }
            pull_rows<T, V>([&](int q) { return rg0 + slot_w(q, r2, r3, r4) * WP; }, n2);
            pull_rows<T, V>([&](int q) { return rg1 + slot_w(q, r2, r3, r4) * WP; }, n3);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            stage_read(k);
            T i1[V], i2[V], i3[V];
            node_reciprocals<T, MODEL, V>(n1, i1);
            node_reciprocals<T, MODEL, V>(n2, i2);
            node_reciprocals<T, MODEL, V>(n3, i3);
            collide_nodes<T, MODEL, V, PACKED>(n1, a.cp, i1);
            collide_nodes<T, MODEL, V, PACKED>(n2, a.cp, i2);
            collide_nodes<T, MODEL, V, PACKED>(n3, a.cp, i3);
            if (a1) to_ring(rg0, n1, w2, w3, w4);
            if (a2) to_ring(rg1, n2, w2, w3, w4);
            if (a3) {
                to_global(out3, n3);
                if (DUAL) to_global(out2, n2);
            }
        } else {
            if (l1) {
                T n1[V][9];
                mbarrier_wait(bar0 + 8u * (uint32_t)(k & 1), (uint32_t)((k >> 1) & 1));
                const T* st = st0 + (k & 1) * 9 * WS;
                pull_rows<T, V>([&](int q) { return st + q * WS; }, n1);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                stage_read(k);
                collide_nodes<T, MODEL, V, PACKED>(n1, a.cp);
                if (a1) to_ring(rg0, n1, w2, w3, w4);
            } else {
                stage_read(k);
            }
            if (l2) {
                T n2[V][9];
                pull_rows<T, V>([&](int q) { return rg0 + slot_w(q, r2, r3, r4) * WP; }, n2);
                collide_nodes<T, MODEL, V, PACKED>(n2, a.cp);
                if (a2) to_ring(rg1, n2, w2, w3, w4);
                if (DUAL && a3 && j - 2 >= xs && j - 2 < xe) to_global(out2, n2);
            }
            if (l3) {
                T n3[V][9];
                pull_rows<T, V>([&](int q) { return rg1 + slot_w(q, r2, r3, r4) * WP; }, n3);
                collide_nodes<T, MODEL, V, PACKED>(n3, a.cp);
                if (a3) to_global(out3, n3);
            }
        }
        named_sync<3>(NTC);
        out3 += a.ld;
        if (DUAL) out2 += a.ld;
        w2 ^= 1;
        w3 = w3 == 2 ? 0 : w3 + 1;
        w4 = (w4 + 1) & 3;
    }
}

int env_knob3(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

template <typename T, int MODEL, bool DUAL>
int launch_3w(const Grid& g, const T* src, T* dst, T* dst_mid, int x_begin, int x_end, const CollideParams<T>& cp, cudaStream_t s)
{
    constexpr int NTC = 128;  // consumer threads
    constexpr int VA = 16 / (int)sizeof(T);
    constexpr int V = VA / 2;  // one row (fp64) / two rows (fp32) per thread, like the default shape of k_lbmn_bulk
    constexpr int HS = (2 * V + 1 + VA - 1) / VA * VA;
    constexpr int W = NTC * V, WS = W + 2 * (HS - 2 * V), WP = W + 2 * V;
    constexpr size_t smem = ((size_t)2 * 9 * WS + 2 * RS_SLOTS * WP) * sizeof(T) + 16;
    constexpr int MINB = (int)((size_t)(228 * 1024) / (smem + 1024));
    static_assert(MINB == 3, "three blocks of five warps per SM");
    if (x_end <= x_begin) return PLBM_OK;
    auto kern = k_lbm3_ws<T, MODEL, V, NTC, MINB, DUAL>;
    static bool configured[64] = {false};
    if (g.device < 64 && !configured[g.device]) {
        PLBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PLBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        configured[g.device] = true;
    }
    Lbm3Args<T> a;
    a.src = src;
    a.dst = dst;
    a.dst_mid = dst_mid;
    a.nx = g.nx;
    a.ny = g.ny;
    a.ld = g.ld;
    a.x_begin = x_begin;
    a.x_end = x_end;
    a.cp = cp;
    // the fewest strips (four redundant rows of V each), segments cut so that the blocks fill whole rounds of MINB blocks per SM
    // (a last, mostly empty round costs as much as a full one); six ramp iterations per segment, at least 8 columns each
    const int ncols = x_end - x_begin;
    const int ty_max = (NTC - 4) * V;
    a.nstrips = (g.ny + ty_max - 1) / ty_max;
    a.ty = ((g.ny + a.nstrips - 1) / a.nstrips + VA - 1) / VA * VA;
    a.nstrips = (g.ny + a.ty - 1) / a.ty;
    // 128-column segments (six ramp iterations each): r02r, 64 / 96 / 128 / 192 / 256 columns within 1 % of each other, 128 best overall
    static const int seg_cols = env_knob3("PLBM_WS_SEGLEN", 128) < 1 ? 128 : env_knob3("PLBM_WS_SEGLEN", 128);
    int nseg = (ncols + seg_cols - 1) / seg_cols;
    const long long slots = (long long)MINB * g.sm_count;
    const long long blocks64 = (long long)a.nstrips * nseg;
    const long long rounds = blocks64 >= slots ? (blocks64 + slots - 1) / slots : 1;
    nseg = (int)(rounds * slots / a.nstrips);
    if (nseg < 1) nseg = 1;
    a.seglen = (ncols + nseg - 1) / nseg;
    if (a.seglen < 8) a.seglen = ncols < 8 ? ncols : 8;
    nseg = (ncols + a.seglen - 1) / a.seglen;
    kern<<<(unsigned)(a.nstrips * nseg), NTC + 32, smem, s>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PLBM_CUDA(cudaGetLastError());
    return PLBM_OK;
}

template <typename T, bool DUAL>
int dispatch_3w(const Grid& g, const T* src, T* dst, T* dst_mid, int x_begin, int x_end, int model, const CollideParams<T>& cp, cudaStream_t s)
{
    // (block shapes measured and dropped, r02r, bench slab BGK fp64: 96 consumer threads x 4 blocks per SM 102.8 GLUPS, 192 x 2 89.1,
    // against 103.5 for 128 x 3)
    switch (model) {
    case M_BGK: return launch_3w<T, M_BGK, DUAL>(g, src, dst, dst_mid, x_begin, x_end, cp, s);
    case M_TRT: return launch_3w<T, M_TRT, DUAL>(g, src, dst, dst_mid, x_begin, x_end, cp, s);
    case M_RR: return launch_3w<T, M_RR, DUAL>(g, src, dst, dst_mid, x_begin, x_end, cp, s);
    case M_BGK_SPLIT: return launch_3w<T, M_BGK_SPLIT, DUAL>(g, src, dst, dst_mid, x_begin, x_end, cp, s);
    case M_TRT_SPLIT: return launch_3w<T, M_TRT_SPLIT, DUAL>(g, src, dst, dst_mid, x_begin, x_end, cp, s);
    case M_BGK_IMPROVED: return launch_3w<T, M_BGK_IMPROVED, DUAL>(g, src, dst, dst_mid, x_begin, x_end, cp, s);
    }
    set_error("launch_lbm_triple_ws: collision model not instantiated");
    return PLBM_ERR_ARG;
}

}  // namespace

// Three fused steps src -> dst for columns [x_begin, x_end) of a grid lbm_multi_applicable(g, model, 3) accepts, periodic self-wrap
// (the halo-reading boundary launches of a slab stay on k_lbmn_bulk<HALO>).  dst_mid != nullptr: also store the state after step 2.
template <typename T>
int launch_lbm_triple_ws(const Grid& g, const T* src, T* dst, T* dst_mid, int x_begin, int x_end, int model, const CollideParams<T>& cp,
                         cudaStream_t s)
{
    if (dst_mid) return dispatch_3w<T, true>(g, src, dst, dst_mid, x_begin, x_end, model, cp, s);
    return dispatch_3w<T, false>(g, src, dst, nullptr, x_begin, x_end, model, cp, s);
}

template int launch_lbm_triple_ws<double>(const Grid&, const double*, double*, double*, int, int, int, const CollideParams<double>&, cudaStream_t);
template int launch_lbm_triple_ws<float>(const Grid&, const float*, float*, float*, int, int, int, const CollideParams<float>&, cudaStream_t);

}  // namespace plbm
