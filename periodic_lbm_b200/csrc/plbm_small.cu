// plbm_small.cu -- cluster-resident multi-step LBM for small grids (sm_100a).
//
// A 64 x 64 fp64 lattice is 295 KB: one fused step is ~2 us of work, so a kernel launch per step
// (~4 us on B200) is pure latency (BASELINE config C1).  Here ONE thread-block cluster (up to 16
// CTAs) keeps both lattices in its distributed shared memory and advances `nsteps` steps in a
// single launch: every CTA owns a slab of lines, pulls the populations it needs from its own
// shared memory or -- for the lines next to a slab boundary -- from the neighbour CTA's shared
// memory over DSMEM (cluster.map_shared_rank), collides, writes the other buffer, and the step
// ends with one cluster barrier.  Same per-node arithmetic as k_lbm, so results are bit-identical.
//   lbm_stream_kernel  src/periodic_lbm.f90:45-127 ;  collisions src/collision_*.F90
#include <cooperative_groups.h>

#include "plbm_internal.h"

namespace cg = cooperative_groups;

namespace plbm {

namespace {

constexpr int MAX_CTAS = 16;  // non-portable cluster size (one GPC of a B200 holds 16+ SMs); falls back to 8

template <typename T, int MODEL>
__global__ void __launch_bounds__(512, 1)
    k_lbm_cluster(T* __restrict__ fa, T* __restrict__ fb, int nx, int ny, int ld, int lpc, int nsteps, CollideParams<T> cp)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int plane = lpc * ny;  // one population of this CTA's slab: [xl][y]
    T* buf0 = reinterpret_cast<T*>(smem_raw);
    T* buf1 = buf0 + 9 * plane;
    const int x_lo = rank * lpc;
    const int nlines = min(lpc, nx - x_lo);
    const int nnodes = nlines * ny;
    const int tid = threadIdx.x, nthr = blockDim.x;

    for (int i = tid; i < 9 * nnodes; i += nthr) {
        const int q = i / nnodes, r = i - q * nnodes, xl = r / ny, y = r - xl * ny;
        buf0[q * plane + xl * ny + y] = fa[((size_t)q * nx + x_lo + xl) * (size_t)ld + y];
    }
    cluster.sync();

    // owners of the lines just outside this slab (periodic ring of CTAs)
    const int lo_x = x_lo == 0 ? nx - 1 : x_lo - 1;
    const int hi_x = x_lo + nlines == nx ? 0 : x_lo + nlines;
    const int lo_rank = lo_x / lpc, hi_rank = hi_x / lpc;
    const int lo_xl = lo_x - lo_rank * lpc, hi_xl = hi_x - hi_rank * lpc;

    for (int s = 0; s < nsteps; ++s) {
        T* src = (s & 1) ? buf1 : buf0;
        T* dst = (s & 1) ? buf0 : buf1;
        const T* src_lo = cluster.map_shared_rank(src, lo_rank);
        const T* src_hi = cluster.map_shared_rank(src, hi_rank);
        for (int n = tid; n < nnodes; n += nthr) {
            const int xl = n / ny, y = n - xl * ny;
            const int ym1 = y == 0 ? ny - 1 : y - 1, yp1 = y + 1 == ny ? 0 : y + 1;
            // line xl-1 / xl+1: own shared memory, or the neighbour CTA's over DSMEM
            const T* wm = xl == 0 ? src_lo + lo_xl * ny : src + (xl - 1) * ny;           // x - 1
            const T* wp = xl + 1 == nlines ? src_hi + hi_xl * ny : src + (xl + 1) * ny;  // x + 1
            const T* wc = src + xl * ny;
            T f[9];
            f[0] = wc[0 * plane + y];
            f[1] = wm[1 * plane + y];
            f[2] = wc[2 * plane + ym1];
            f[3] = wp[3 * plane + y];
            f[4] = wc[4 * plane + yp1];
            f[5] = wm[5 * plane + ym1];
            f[6] = wp[6 * plane + ym1];
            f[7] = wp[7 * plane + yp1];
            f[8] = wm[8 * plane + yp1];
            collide<T, MODEL>(f, cp);
            T* d = dst + xl * ny + y;
#pragma unroll
            for (int q = 0; q < 9; ++q) d[q * plane] = f[q];
        }
        cluster.sync();
    }

    // newest state -> the lattice the reference's index swaps make `iold`; previous state -> `inew`
    const T* newest = (nsteps & 1) ? buf1 : buf0;
    const T* previous = (nsteps & 1) ? buf0 : buf1;
    T* g_new = (nsteps & 1) ? fb : fa;
    T* g_prev = (nsteps & 1) ? fa : fb;
    for (int i = tid; i < 9 * nnodes; i += nthr) {
        const int q = i / nnodes, r = i - q * nnodes, xl = r / ny, y = r - xl * ny;
        const size_t gi = ((size_t)q * nx + x_lo + xl) * (size_t)ld + y;
        g_new[gi] = newest[q * plane + xl * ny + y];
        if (nsteps > 0) g_prev[gi] = previous[q * plane + xl * ny + y];
    }
}

template <typename T, int MODEL>
int launch_cluster(const Grid& g, T* fa, T* fb, int ncta, int lpc, size_t smem, int nsteps, const CollideParams<T>& cp, cudaStream_t s)
{
    static bool configured[64] = {false};
    if (g.device < 64 && !configured[g.device]) {
        PLBM_CUDA(cudaFuncSetAttribute(k_lbm_cluster<T, MODEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        PLBM_CUDA(cudaFuncSetAttribute(k_lbm_cluster<T, MODEL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        configured[g.device] = true;
    }
    int threads = lpc * g.ny;
    threads = threads > 512 ? 512 : (threads + 31) / 32 * 32;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ncta);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ncta;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // Ask first, for every cluster size: a 16-CTA cluster needs a GPC with 16 free SMs, and on a partitioned device (MIG,
    // green contexts, SM-limited GPCs) even 8 or fewer may not be schedulable.  "Not applicable" (-1) lets the caller try a
    // smaller cluster and finally fall back to the per-step kernels instead of failing perform_lbm_step.
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, k_lbm_cluster<T, MODEL>, &cfg) != cudaSuccess || nclusters < 1) {
        cudaGetLastError();
        return -1;
    }
    const cudaError_t e = cudaLaunchKernelEx(&cfg, k_lbm_cluster<T, MODEL>, fa, fb, g.nx, g.ny, g.ld, lpc, nsteps, cp);
    if (e == cudaErrorInvalidConfiguration || e == cudaErrorLaunchOutOfResources || e == cudaErrorInvalidValue ||
        e == cudaErrorNotSupported) {
        cudaGetLastError();  // a launch-configuration failure launched nothing and is not sticky: clear it
        return -1;
    }
    PLBM_CUDA(e);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return PLBM_OK;
}

}  // namespace

// Returns PLBM_OK and sets *done = true when the steps were executed by the cluster kernel;
// *done = false means "does not apply" (grid too large for shared memory): use the per-step kernels.
template <typename T>
int try_lbm_cluster_steps(const Grid& g, T* f_iold, T* f_inew, int model, const CollideParams<T>& cp, int nsteps, bool* done,
                          cudaStream_t s)
{
    *done = false;
    if (g.nx < 2 || g.ny < 2) return PLBM_OK;
    int rc = -1;
    for (int max_ctas = MAX_CTAS; max_ctas >= 2 && rc == -1; max_ctas /= 2) {
        int ncta = g.nx < max_ctas ? g.nx : max_ctas;
        const int lpc = (g.nx + ncta - 1) / ncta;
        ncta = (g.nx + lpc - 1) / lpc;  // every CTA owns at least one line
        const size_t smem = 2 * 9 * (size_t)lpc * g.ny * sizeof(T);
        if (smem > 200 * 1024) return PLBM_OK;
        switch (model) {
        case M_BGK: rc = launch_cluster<T, M_BGK>(g, f_iold, f_inew, ncta, lpc, smem, nsteps, cp, s); break;
        case M_TRT: rc = launch_cluster<T, M_TRT>(g, f_iold, f_inew, ncta, lpc, smem, nsteps, cp, s); break;
        case M_RR: rc = launch_cluster<T, M_RR>(g, f_iold, f_inew, ncta, lpc, smem, nsteps, cp, s); break;
        case M_BGK_SPLIT: rc = launch_cluster<T, M_BGK_SPLIT>(g, f_iold, f_inew, ncta, lpc, smem, nsteps, cp, s); break;
        case M_TRT_SPLIT: rc = launch_cluster<T, M_TRT_SPLIT>(g, f_iold, f_inew, ncta, lpc, smem, nsteps, cp, s); break;
        case M_BGK_IMPROVED: rc = launch_cluster<T, M_BGK_IMPROVED>(g, f_iold, f_inew, ncta, lpc, smem, nsteps, cp, s); break;
        default: return PLBM_OK;
        }
    }
    if (rc == -1) return PLBM_OK;  // no cluster shape available: per-step kernels
    if (rc) return rc;
    *done = true;
    return PLBM_OK;
}

template int try_lbm_cluster_steps<double>(const Grid&, double*, double*, int, const CollideParams<double>&, int, bool*, cudaStream_t);
template int try_lbm_cluster_steps<float>(const Grid&, float*, float*, int, const CollideParams<float>&, int, bool*, cudaStream_t);

}  // namespace plbm
