// plbm_fvm_tma.cu -- the FVM / DUGKS tile kernel with TMA staging and an mbarrier pipeline (sm_100a).
//
// Same arithmetic as k_fv_fused (plbm_fvm.cu), different data movement: blocks are persistent
// (2 per SM) and walk over the (FY x FX) tiles of the grid; one elected thread issues a single
// `cp.async.bulk.tensor.3d` per tile -- a 16-byte-aligned box covering the (FY+2) x (FX+2) x 9
// halo tile of the lattice viewed as a (ny, nx, 9) tensor -- into the *other* half of a two-stage shared-memory ring while all 256
// threads work on the current tile, so the global-memory latency of the 3x3-stencil halo tile
// is off the critical path (the plain-load kernel stalls on it: ncu long_scoreboard + barrier).
// TMA zero-fills out-of-range coordinates; the periodic wrap is restored by the threads that own
// an out-of-range cell (boundary tiles only), which fetch it from the wrapped global address.
#include <cuda.h>

#include <cstdlib>

#include "plbm_internal.h"
#include "plbm_fv.cuh"

// plbm_fvm_tma_fma.cu compiles this file a second time with -fmad=true (see there): the same kernels with their
// multiply-adds contracted, exported as launch_fv_tma_fma.  The kernels live in an anonymous namespace and device
// code is not linked across translation units, so the two builds do not meet.
#ifdef PLBM_FMA_BUILD
#define launch_fv_tma launch_fv_tma_fma
#endif

namespace plbm {

namespace {

#ifndef PLBM_TMA_FX
#define PLBM_TMA_FX 8
#endif
constexpr int FY = 32, FX = PLBM_TMA_FX;
constexpr int GY = FY + 2, GX = FX + 2;
constexpr int NHALO = GX * GY - FX * FY;
constexpr int FPITCH = GY + 1;  // fbar tile pitch (DUGKS)

template <typename T> struct Box {
    // TMA needs the box to START on a 16-byte boundary of the unit-stride dimension (measured:
    // an odd fp64 start coordinate is an illegal instruction, tools/tma_probe.cu) and its inner
    // extent to be a multiple of 16 bytes.  The halo tile starts at y0-1, so the box starts OFF
    // elements before y0 (2 doubles / 4 floats) and is BY wide; tile column sy lives at sy + COL0.
    static constexpr int OFF = 16 / (int)sizeof(T);
    static constexpr int COL0 = OFF - 1;
    static constexpr int BY = sizeof(T) == 8 ? 36 : 40;
    static constexpr int PLANE = GX * BY;
    static constexpr int STAGE_ELEMS = 9 * PLANE;
    static constexpr int STAGE_BYTES = (STAGE_ELEMS * (int)sizeof(T) + 127) / 128 * 128;
    static constexpr int TX_BYTES = STAGE_ELEMS * (int)sizeof(T);
    static constexpr int FBAR_BYTES = (9 * GX * FPITCH * (int)sizeof(T) + 127) / 128 * 128;
};

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
        if (!ok && ++spins > (1ll << 28)) __trap();  // a lost TMA must fail loudly, never hang the GPU
    } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}

__device__ __forceinline__ int pmod(int i, int n)
{
    i %= n;
    return i < 0 ? i + n : i;
}

// Line (row 0) of population q at slow index gx, which may lie outside the slab: periodic wrap on a
// single GPU, the ring neighbours' halo lines ([9][ld]) under a slab decomposition.
template <typename T>
__device__ __forceinline__ const T* wrapped_line(const T* fin, const T* hlo, const T* hhi, int q, int gx, int nx, int ld)
{
    if (gx < 0) {
        if (hlo) return hlo + (size_t)q * ld;
        gx = pmod(gx, nx);
    } else if (gx >= nx) {
        if (hhi) return hhi + (size_t)q * ld;
        gx = pmod(gx, nx);
    }
    return fin + ((size_t)q * nx + gx) * (size_t)ld;
}

// MODE_LBM: the standard pull stream + collide with the halo tile staged by TMA (variant 3 of
// perform_lbm_step).  It exists to MEASURE the north-star's "TMA staging of the row halo" against the
// direct-load kernel k_lbm: staging cannot reduce the 144 B/node and over-fetches the halo ring.
// MODE_FDM_BARDOW / MODE_FDM_SOFONEA: the reference's finite-difference streaming schemes.
enum { MODE_DUGKS = 0, MODE_DUGKS_OFF = 1, MODE_BARDOW = 2, MODE_LBM = 3, MODE_FDM_BARDOW = 4, MODE_FDM_SOFONEA = 5 };

// DUGKS keeps ONE raw stage (+ the fbar tile): the raw tile is dead after stage 1, so the next tile's
// TMA is issued right after the stage-1 barrier and lands during stage 2 (the long phase).  Bardow
// reads the raw tile in stage 2 and therefore double-buffers it.
template <typename T, int MODE, int MODEL, int MINB>
__global__ void __launch_bounds__(FY* FX, MINB)
    k_fv_tma(const __grid_constant__ CUtensorMap tmap, const T* __restrict__ fin, T* __restrict__ fout, int nx, int ny, int ld,
             int nty, int ntiles, T dt, T omega_full, T omega_half, T omega_face, CollideParams<T> cp, const T* __restrict__ hlo,
             const T* __restrict__ hhi)
{
    using B = Box<T>;
    constexpr bool IS_DUGKS = MODE == MODE_DUGKS || MODE == MODE_DUGKS_OFF;
    extern __shared__ __align__(128) unsigned char smem[];
    T* const raw0 = reinterpret_cast<T*>(smem);
    T* const raw1 = reinterpret_cast<T*>(smem + B::STAGE_BYTES);  // Bardow only
    T* fbar = reinterpret_cast<T*>(smem + B::STAGE_BYTES);        // DUGKS only
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (IS_DUGKS ? B::STAGE_BYTES + B::FBAR_BYTES : 2 * B::STAGE_BYTES));
    const uint32_t bar0 = smem_u32(&bars[0]), bar1 = smem_u32(&bars[1]);

    const int tx = threadIdx.y, ty = threadIdx.x;
    const int tid = tx * FY + ty;

    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // NB: the descriptor must be addressed in the kernel-parameter space: take &tmap directly in the
    // kernel body (a lambda capture would hand TMA a local-memory copy -> illegal instruction).
#define PLBM_ISSUE_TILE(t_, s_)                                                                          \
    do {                                                                                                 \
        const int ty0_ = ((t_) % nty) * FY, tx0_ = ((t_) / nty) * FX;                                    \
        const uint32_t b_ = (s_) ? bar1 : bar0;                                                          \
        mbar_expect_tx(b_, B::TX_BYTES);                                                                 \
        tma_load_3d(smem_u32((s_) ? raw1 : raw0), &tmap, ty0_ - B::OFF, tx0_ - 1, 0, b_);                     \
    } while (0)
    if (tid == 0 && (int)blockIdx.x < ntiles) PLBM_ISSUE_TILE((int)blockIdx.x, 0);

    // halo-ring cell owned by this thread (first NHALO threads)
    int hsx = 0, hsy = 0;
    if (tid < NHALO) {
        if (tid < 2 * GY) {
            hsx = tid < GY ? 0 : GX - 1;
            hsy = tid < GY ? tid : tid - GY;
        } else {
            const int k = tid - 2 * GY;
            hsx = 1 + (k >> 1);
            hsy = (k & 1) ? GY - 1 : 0;
        }
    }

    int k = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++k) {
        const int s = IS_DUGKS ? 0 : (k & 1);
        const int tnext = t + gridDim.x;
        if (!IS_DUGKS) {
            if (tid == 0 && tnext < ntiles) PLBM_ISSUE_TILE(tnext, s ^ 1);  // stage s^1 was released by the barrier ending iteration k-1
            mbar_wait(s ? bar1 : bar0, (k >> 1) & 1);
        } else {
            mbar_wait(bar0, k & 1);
        }

        const int y0 = (t % nty) * FY, x0 = (t / nty) * FX;
        const int x = x0 + tx, y = y0 + ty;
        const bool active = x < nx && y < ny;
        T* rw = s ? raw1 : raw0;
        T fp[9];

        // ---- stage 1 -------------------------------------------------------------------
        {
            T b[9];
            const T* c = rw + (tx + 1) * B::BY + (ty + 1) + B::COL0;
            if (active) {
#pragma unroll
                for (int q = 0; q < 9; ++q) b[q] = c[q * B::PLANE];
            } else {  // tile overhangs the grid: TMA zero-filled this cell, fetch the wrapped node
                const int ys = pmod(y, ny);
#pragma unroll
                for (int q = 0; q < 9; ++q) b[q] = wrapped_line(fin, hlo, hhi, q, x, nx, ld)[ys];
            }
#pragma unroll
            for (int q = 0; q < 9; ++q) fp[q] = b[q];
            if (IS_DUGKS) {
                collide_bgk_split(b, omega_half);
                collide_bgk_split(fp, omega_full);
                T* d = fbar + (tx + 1) * FPITCH + (ty + 1);
#pragma unroll
                for (int q = 0; q < 9; ++q) d[q * (GX * FPITCH)] = b[q];
            } else if (!active) {
                T* d = rw + (tx + 1) * B::BY + (ty + 1) + B::COL0;
#pragma unroll
                for (int q = 0; q < 9; ++q) d[q * B::PLANE] = b[q];
            }
        }
        if (tid < NHALO) {
            const int gx = x0 + hsx - 1, gy = y0 + hsy - 1;
            const bool inb = gx >= 0 && gx < nx && gy >= 0 && gy < ny;
            T b[9];
            T* c = rw + hsx * B::BY + hsy + B::COL0;
            if (inb) {
                if (IS_DUGKS) {
#pragma unroll
                    for (int q = 0; q < 9; ++q) b[q] = c[q * B::PLANE];
                }
            } else {
                const int ys = pmod(gy, ny);
#pragma unroll
                for (int q = 0; q < 9; ++q) b[q] = wrapped_line(fin, hlo, hhi, q, gx, nx, ld)[ys];
            }
            if (IS_DUGKS) {
                collide_bgk_split(b, omega_half);
                T* d = fbar + hsx * FPITCH + hsy;
#pragma unroll
                for (int q = 0; q < 9; ++q) d[q * (GX * FPITCH)] = b[q];
            } else if (!inb) {
#pragma unroll
                for (int q = 0; q < 9; ++q) c[q * B::PLANE] = b[q];
            }
        }
        if (IS_DUGKS) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // raw reads before the TMA refill
        __syncthreads();
        if (IS_DUGKS && tid == 0 && tnext < ntiles) PLBM_ISSUE_TILE(tnext, 0);  // lands during stage 2

        // ---- stage 2 -------------------------------------------------------------------
        if (active) {
            if (IS_DUGKS) {
                const T* c0 = fbar + (tx + 1) * FPITCH + (ty + 1);
                flux_update<T, MODE == MODE_DUGKS, FPITCH, GX * FPITCH>(c0, dt, omega_face, fp);
            } else if (MODE == MODE_BARDOW) {
                const T* c0 = rw + (tx + 1) * B::BY + (ty + 1) + B::COL0;
                flux_update<T, false, B::BY, B::PLANE>(c0, dt, omega_face, fp);
                if (MODEL != M_NONE) collide<T, MODEL>(fp, cp);
            } else if (MODE == MODE_FDM_BARDOW || MODE == MODE_FDM_SOFONEA) {
                const T* c0 = rw + (tx + 1) * B::BY + (ty + 1) + B::COL0;
                fdm_update<T, MODE == MODE_FDM_SOFONEA, B::BY, B::PLANE>(c0, dt, fp);
                if (MODEL != M_NONE) collide<T, MODEL>(fp, cp);
            } else {  // MODE_LBM: fdst(y,x,q) = fsrc(y-cy, x-cx, q) from the staged tile, then collide
                const T* c0 = rw + (tx + 1) * B::BY + (ty + 1) + B::COL0;
#pragma unroll
                for (int q = 1; q < 9; ++q) fp[q] = c0[q * B::PLANE - cxi(q) * B::BY - cyi(q)];
                if (MODEL != M_NONE) collide<T, MODEL>(fp, cp);
            }
#pragma unroll
            for (int q = 0; q < 9; ++q) fout[((size_t)q * nx + x) * (size_t)ld + y] = fp[q];
        }
        // order this tile's generic-proxy accesses to raw[s] before the TMA (async proxy) write that
        // refills it two iterations from now
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();  // fbar / raw[s] may be overwritten from here on
    }
#undef PLBM_ISSUE_TILE
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

template <typename T, int MODE, int MODEL, int MINB>
int launch_one(const Grid& g, int which_src, const T* fin, T* fout, T dt, T of, T oh, T oc, const CollideParams<T>& cp, cudaStream_t s)
{
    using B = Box<T>;
    constexpr bool IS_DUGKS = MODE == MODE_DUGKS || MODE == MODE_DUGKS_OFF;
    const int nty = (g.ny + FY - 1) / FY, ntx = (g.nx + FX - 1) / FX;
    const int ntiles = nty * ntx;
    const size_t smem = (IS_DUGKS ? (size_t)B::STAGE_BYTES + B::FBAR_BYTES : 2 * (size_t)B::STAGE_BYTES) + 16;
    static bool configured[64] = {false};
    if (g.device < 64 && !configured[g.device]) {
        PLBM_CUDA(cudaFuncSetAttribute(k_fv_tma<T, MODE, MODEL, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[g.device] = true;
    }
    int nblocks = MINB * g.sm_count;
    if (nblocks > ntiles) nblocks = ntiles;
    CUtensorMap map;
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
    memcpy(&map, g.tmap[which_src - 1], sizeof(map));
    k_fv_tma<T, MODE, MODEL, MINB><<<nblocks, dim3(FY, FX), smem, s>>>(map, fin, fout, g.nx, g.ny, g.ld, nty, ntiles, dt, of, oh, oc, cp,
                                                                       static_cast<const T*>(g.fv_halo_lo), static_cast<const T*>(g.fv_halo_hi));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PLBM_CUDA(cudaGetLastError());
    return PLBM_OK;
}

}  // namespace

#ifndef PLBM_FMA_BUILD
int make_tensor_maps(Grid& g)
{
    g.tmap_ok = false;
    EncodeTiledFn enc = encode_fn();
    if (!enc) return PLBM_OK;  // no driver entry point: the plain-load tile kernel is used instead
    const bool f64 = g.prec == PLBM_F64;
    const cuuint64_t es = f64 ? 8 : 4;
    const cuuint64_t gdim[3] = {(cuuint64_t)g.ny, (cuuint64_t)g.nx, 9};
    const cuuint64_t gstride[2] = {(cuuint64_t)g.ld * es, (cuuint64_t)g.ld * g.nx * es};
    const cuuint32_t box[3] = {(cuuint32_t)(f64 ? Box<double>::BY : Box<float>::BY), (cuuint32_t)GX, 9};
    const cuuint32_t estr[3] = {1, 1, 1};
    for (int i = 0; i < g.nf; ++i) {
        CUtensorMap m;
        CUresult r = enc(&m, f64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, g.f[i], gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return PLBM_OK;  // e.g. stride limits on extreme shapes: fall back to the plain-load kernel
        memcpy(g.tmap[i], &m, sizeof(m));
    }
    g.tmap_ok = true;
    return PLBM_OK;
}
#endif  // !PLBM_FMA_BUILD

template <typename T>
int launch_fv_tma(const Grid& g, int which_src, const T* fin, T* fout, int mode, int model, T dt, T of, T oh, T oc,
                  const CollideParams<T>& cp, cudaStream_t s)
{
    // Blocks per SM for the DUGKS instantiation, measured on B200 at 2048^2 (profiles/): fp64 is fastest
    // with ONE block per SM and an uncapped register budget (230 regs, no spills: 14.9 GLUPS vs 13.8 at two
    // blocks with 128 regs + spills); fp32 is fastest with two (26.7 vs 21.5 GLUPS).  PLBM_FV_MINB overrides.
    static const int minb_env = getenv("PLBM_FV_MINB") ? atoi(getenv("PLBM_FV_MINB")) : 0;
    const int minb = minb_env ? minb_env : (sizeof(T) == 8 ? 1 : 2);
    if (mode == MODE_DUGKS) {
        if (minb == 1) return launch_one<T, MODE_DUGKS, M_NONE, 1>(g, which_src, fin, fout, dt, of, oh, oc, cp, s);
        if (minb == 3) return launch_one<T, MODE_DUGKS, M_NONE, 3>(g, which_src, fin, fout, dt, of, oh, oc, cp, s);
        if (minb == 4) return launch_one<T, MODE_DUGKS, M_NONE, 4>(g, which_src, fin, fout, dt, of, oh, oc, cp, s);
        return launch_one<T, MODE_DUGKS, M_NONE, 2>(g, which_src, fin, fout, dt, of, oh, oc, cp, s);
    }
    if (mode == MODE_DUGKS_OFF) return launch_one<T, MODE_DUGKS_OFF, M_NONE, 2>(g, which_src, fin, fout, dt, of, oh, oc, cp, s);
    if (mode == MODE_LBM) {
        switch (model) {
        case M_BGK: return launch_one<T, MODE_LBM, M_BGK, 2>(g, which_src, fin, fout, dt, of, oh, oc, cp, s);
        case M_TRT: return launch_one<T, MODE_LBM, M_TRT, 2>(g, which_src, fin, fout, dt, of, oh, oc, cp, s);
        case M_RR: return launch_one<T, MODE_LBM, M_RR, 2>(g, which_src, fin, fout, dt, of, oh, oc, cp, s);
        default: set_error("fv_tma: the TMA-staged LBM variant supports bgk, trt, rr"); return PLBM_ERR_ARG;
        }
    }
    if (mode == MODE_FDM_BARDOW || mode == MODE_FDM_SOFONEA) {
        const bool sof = mode == MODE_FDM_SOFONEA;
        switch (model) {
        case M_NONE: return sof ? launch_one<T, MODE_FDM_SOFONEA, M_NONE, 2>(g, which_src, fin, fout, dt, of, oh, oc, cp, s)
                                : launch_one<T, MODE_FDM_BARDOW, M_NONE, 2>(g, which_src, fin, fout, dt, of, oh, oc, cp, s);
        case M_BGK: return sof ? launch_one<T, MODE_FDM_SOFONEA, M_BGK, 2>(g, which_src, fin, fout, dt, of, oh, oc, cp, s)
                               : launch_one<T, MODE_FDM_BARDOW, M_BGK, 2>(g, which_src, fin, fout, dt, of, oh, oc, cp, s);
        case M_TRT: return sof ? launch_one<T, MODE_FDM_SOFONEA, M_TRT, 2>(g, which_src, fin, fout, dt, of, oh, oc, cp, s)
                               : launch_one<T, MODE_FDM_BARDOW, M_TRT, 2>(g, which_src, fin, fout, dt, of, oh, oc, cp, s);
        case M_RR: return sof ? launch_one<T, MODE_FDM_SOFONEA, M_RR, 2>(g, which_src, fin, fout, dt, of, oh, oc, cp, s)
                              : launch_one<T, MODE_FDM_BARDOW, M_RR, 2>(g, which_src, fin, fout, dt, of, oh, oc, cp, s);
        default: set_error("fv_tma: the FDM streaming kernels fuse bgk, trt, rr (others: stream, then collide)"); return PLBM_ERR_ARG;
        }
    }
    switch (model) {
    case M_NONE: return launch_one<T, MODE_BARDOW, M_NONE, 2>(g, which_src, fin, fout, dt, of, oh, oc, cp, s);
    case M_BGK: return launch_one<T, MODE_BARDOW, M_BGK, 2>(g, which_src, fin, fout, dt, of, oh, oc, cp, s);
    case M_TRT: return launch_one<T, MODE_BARDOW, M_TRT, 2>(g, which_src, fin, fout, dt, of, oh, oc, cp, s);
    case M_RR: return launch_one<T, MODE_BARDOW, M_RR, 2>(g, which_src, fin, fout, dt, of, oh, oc, cp, s);
    case M_BGK_SPLIT: return launch_one<T, MODE_BARDOW, M_BGK_SPLIT, 2>(g, which_src, fin, fout, dt, of, oh, oc, cp, s);
    case M_TRT_SPLIT: return launch_one<T, MODE_BARDOW, M_TRT_SPLIT, 2>(g, which_src, fin, fout, dt, of, oh, oc, cp, s);
    case M_BGK_IMPROVED: return launch_one<T, MODE_BARDOW, M_BGK_IMPROVED, 2>(g, which_src, fin, fout, dt, of, oh, oc, cp, s);
    }
    set_error("fv_tma: unknown collision model");
    return PLBM_ERR_ARG;
}

template int launch_fv_tma<double>(const Grid&, int, const double*, double*, int, int, double, double, double, double,
                                   const CollideParams<double>&, cudaStream_t);
template int launch_fv_tma<float>(const Grid&, int, const float*, float*, int, int, float, float, float, float,
                                  const CollideParams<float>&, cudaStream_t);

}  // namespace plbm
