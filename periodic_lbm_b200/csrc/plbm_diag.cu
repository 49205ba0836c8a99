// plbm_diag.cu -- initial condition, macroscopic moments, vorticity and the warp-shuffle
// reductions for the moment / energy diagnostics (sm_100a).
//
//   set_pdf_to_equilibrium  src/fvm_bardow.F90:272-305
//   update_macros_kernel    src/fvm_bardow.F90:356-388      (9 reads + 3 writes per node)
//   vorticity_2nd / _4th    src/vorticity.f90:13-87         (4th-order weights as shipped)
//   maxval/minval(hypot(ux,uy)), norm2(hypot(..))  app/main_taylor_green.f90:106,155,195-198
#include "plbm_internal.h"

namespace plbm {

// ---- init / macros ---------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_init_eq(const T* __restrict__ rho, const T* __restrict__ ux,
                                                 const T* __restrict__ uy, T* __restrict__ f, int nx, int ny, int ld)
{
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int x = (int)(g / (size_t)ld);
    const int y = (int)(g - (size_t)x * ld);
    if (x >= nx || y >= ny) return;
    const size_t m = (size_t)x * ny + y;
    T feq[9];
    equilibrium(rho[m], ux[m], uy[m], feq);
#pragma unroll
    for (int q = 0; q < 9; ++q) f[((size_t)q * nx + x) * (size_t)ld + y] = feq[q];
}

template <typename T>
__global__ void __launch_bounds__(256) k_macros(const T* __restrict__ f, T* __restrict__ rho, T* __restrict__ ux,
                                                T* __restrict__ uy, int nx, int ny, int ld)
{
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int x = (int)(g / (size_t)ld);
    const int y = (int)(g - (size_t)x * ld);
    if (x >= nx || y >= ny) return;
    T fs[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) fs[q] = f[((size_t)q * nx + x) * (size_t)ld + y];
    T r, u, v;
    macros(fs, r, u, v);
    const size_t m = (size_t)x * ny + y;
    rho[m] = r;
    ux[m] = u;
    uy[m] = v;
}

template <typename T> int launch_init_eq(const Grid& g, T* f, cudaStream_t s)
{
    const size_t n = (size_t)g.nx * g.ld;
    k_init_eq<T><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(g.rho<T>(), g.ux<T>(), g.uy<T>(), f, g.nx, g.ny, g.ld);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PLBM_CUDA(cudaGetLastError());
    return PLBM_OK;
}

template <typename T> int launch_macros(const Grid& g, const T* f, cudaStream_t s)
{
    const size_t n = (size_t)g.nx * g.ld;
    k_macros<T><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(f, g.rho<T>(), g.ux<T>(), g.uy<T>(), g.nx, g.ny, g.ld);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PLBM_CUDA(cudaGetLastError());
    return PLBM_OK;
}

// ---- vorticity ---------------------------------------------------------------------------
// Periodic wrap in both directions on one GPU.  Under a slab decomposition the x-derivative of uy reaches into the ring
// neighbours' slabs: their two nearest lines of uy arrive in uy_lo (lines -2, -1) and uy_hi (lines nx, nx + 1), [2][ny] each.
template <typename T, int ORDER>
__global__ void __launch_bounds__(256) k_vorticity(const T* __restrict__ ux, const T* __restrict__ uy,
                                                   T* __restrict__ om, int nx, int ny, const T* __restrict__ uy_lo,
                                                   const T* __restrict__ uy_hi)
{
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int x = (int)(g / (size_t)ny);
    const int y = (int)(g - (size_t)x * ny);
    if (x >= nx) return;
    auto wp = [](int i, int n) { return i >= n ? i - n : i; };
    auto wm = [](int i, int n) { return i < 0 ? i + n : i; };
    const int yp1 = wp(y + 1, ny), ym1 = wm(y - 1, ny);
#define M(yy, xx) ((size_t)(xx) * ny + (yy))
    // uy on line xx in [-2, nx + 1]: the neighbours' lines under a ring, the periodic image otherwise
    // (mod(x+1,nx)+1 etc. on 1-based indices in the reference; the +-2 neighbours may wrap twice when nx < 2)
    auto uyx = [&](int xx) -> T {
        if (uy_lo != nullptr) {
            if (xx < 0) return uy_lo[M(y, xx + 2)];
            if (xx >= nx) return uy_hi[M(y, xx - nx)];
            return uy[M(y, xx)];
        }
        return uy[M(y, ((xx % nx) + nx) % nx)];
    };
    T duydx, duxdy;
    if (ORDER == 2) {
        duydx = T(0.5) * (uyx(x + 1) - uyx(x - 1));
        duxdy = T(0.5) * (ux[M(yp1, x)] - ux[M(ym1, x)]);
    } else {
        const T t1 = T(1) / T(12), t2 = T(2) / T(3);
        const int yp2 = (y + 2) % ny, ym2 = (ny + y - 2 + ny) % ny;
        duydx = t1 * (uyx(x + 1) - uyx(x - 1)) + t2 * (uyx(x - 2) - uyx(x + 2));
        duxdy = t1 * (ux[M(yp1, x)] - ux[M(ym1, x)]) + t2 * (ux[M(ym2, x)] - ux[M(yp2, x)]);
    }
#undef M
    om[(size_t)x * ny + y] = duydx - duxdy;
}

template <typename T> int launch_vorticity(const Grid& g, int order, const T* ux, const T* uy, T* out, cudaStream_t s, const T* uy_lo, const T* uy_hi)
{
    const size_t n = (size_t)g.nx * g.ny;
    const unsigned nb = (unsigned)((n + 255) / 256);
    if (order == 2)
        k_vorticity<T, 2><<<nb, 256, 0, s>>>(ux, uy, out, g.nx, g.ny, uy_lo, uy_hi);
    else if (order == 4)
        k_vorticity<T, 4><<<nb, 256, 0, s>>>(ux, uy, out, g.nx, g.ny, uy_lo, uy_hi);
    else {
        set_error("vorticity: order must be 2 or 4");
        return PLBM_ERR_ARG;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PLBM_CUDA(cudaGetLastError());
    return PLBM_OK;
}

// ---- reductions ---------------------------------------------------------------------------
// Two-stage and deterministic: each block reduces its grid-stride share with warp shuffles and
// writes one partial per quantity; the (few hundred) partials are combined on the host in a
// fixed order.  Accumulation is in double for both precisions.
struct Red4 {
    double a, b, c, d;
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// max / min that PROPAGATE NaN (fmax / fmin drop it): a diverged field must not report max |u| = 0.  Any NaN speed makes
// the result NaN (gfortran's maxval / minval, app/main_taylor_green.f90:106, return NaN only for an all-NaN array and skip
// NaNs otherwise; for a blow-up detector "any" is the useful reading, and sum(rho) / kinetic energy behave that way too).
__host__ __device__ __forceinline__ double nan_max(double a, double b) { return a != a ? a : ((b > a || b != b) ? b : a); }
__host__ __device__ __forceinline__ double nan_min(double a, double b) { return a != a ? a : ((b < a || b != b) ? b : a); }
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = nan_max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = nan_min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// MODE 0: a = max |u|, b = min |u|, c = sum rho, d = 1/2 sum rho |u|^2
// MODE 1: a = sum |u-ua|^2, b = sum |ua|^2   (p0 = uxa, p1 = uya)
template <typename T, int MODE>
__global__ void __launch_bounds__(256) k_reduce(const T* __restrict__ rho, const T* __restrict__ ux,
                                                const T* __restrict__ uy, const T* __restrict__ p0,
                                                const T* __restrict__ p1, size_t n, Red4* __restrict__ partial)
{
    double a = MODE == 0 ? 0.0 : 0.0, b = MODE == 0 ? 1e300 : 0.0, c = 0.0, d = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double u = (double)ux[i], v = (double)uy[i];
        if (MODE == 0) {
            const double sp = hypot(u, v);
            a = nan_max(a, sp);
            b = nan_min(b, sp);
            const double r = (double)rho[i];
            c += r;
            d += 0.5 * r * (u * u + v * v);
        } else {
            const double ua = (double)p0[i], va = (double)p1[i];
            a += (u - ua) * (u - ua) + (v - va) * (v - va);
            b += ua * ua + va * va;
        }
    }
    if (MODE == 0) {
        a = warp_max(a);
        b = warp_min(b);
        c = warp_sum(c);
        d = warp_sum(d);
    } else {
        a = warp_sum(a);
        b = warp_sum(b);
    }
    __shared__ Red4 sm[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sm[w] = Red4{a, b, c, d};
    __syncthreads();
    if (threadIdx.x == 0) {
        Red4 r = sm[0];
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) {
            if (MODE == 0) {
                r.a = nan_max(r.a, sm[i].a);
                r.b = nan_min(r.b, sm[i].b);
                r.c += sm[i].c;
                r.d += sm[i].d;
            } else {
                r.a += sm[i].a;
                r.b += sm[i].b;
            }
        }
        partial[blockIdx.x] = r;
    }
}

template <typename T, int MODE>
static int reduce_impl(Grid& g, const T* p0, const T* p1, Red4& out, cudaStream_t s)
{
    const size_t n = (size_t)g.nx * g.ny;
    int nb = (int)((n + 255) / 256);
    if (nb > g.npartial) nb = g.npartial;
    k_reduce<T, MODE><<<nb, 256, 0, s>>>(g.rho<T>(), g.ux<T>(), g.uy<T>(), p0, p1, n, static_cast<Red4*>(g.partial));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PLBM_CUDA(cudaGetLastError());
    PLBM_CUDA(cudaMemcpyAsync(g.partial_host, g.partial, sizeof(Red4) * nb, cudaMemcpyDeviceToHost, s));
    PLBM_CUDA(cudaStreamSynchronize(s));
    const Red4* h = static_cast<const Red4*>(g.partial_host);
    out = h[0];
    for (int i = 1; i < nb; ++i) {
        if (MODE == 0) {
            out.a = nan_max(out.a, h[i].a);
            out.b = nan_min(out.b, h[i].b);
            out.c += h[i].c;
            out.d += h[i].d;
        } else {
            out.a += h[i].a;
            out.b += h[i].b;
        }
    }
    // slab decomposition: the diagnostics are those of the GLOBAL grid (app/main_taylor_green.f90:106,155,195-198 reduce over
    // the whole field) -- every rank of the ring calls this, and every rank gets the same numbers
    if (g.comm) {
        int rc;
        if (MODE == 0) {
            if ((rc = comm_allreduce(g, &out.a, 1, PLBM_REDUCE_MAX))) return rc;
            if ((rc = comm_allreduce(g, &out.b, 1, PLBM_REDUCE_MIN))) return rc;
            if ((rc = comm_allreduce(g, &out.c, 2, PLBM_REDUCE_SUM))) return rc;
        } else {
            if ((rc = comm_allreduce(g, &out.a, 2, PLBM_REDUCE_SUM))) return rc;
        }
    }
    return PLBM_OK;
}

template <typename T> int launch_diagnostics(Grid& g, double out[PLBM_DIAG_COUNT], cudaStream_t s)
{
    Red4 r{};
    int rc = reduce_impl<T, 0>(g, nullptr, nullptr, r, s);
    if (rc) return rc;
    out[PLBM_DIAG_MAX_SPEED] = r.a;
    out[PLBM_DIAG_MIN_SPEED] = r.b;
    out[PLBM_DIAG_SUM_RHO] = r.c;
    out[PLBM_DIAG_KINETIC] = r.d;
    return PLBM_OK;
}

template <typename T> int launch_l2_sums(Grid& g, const T* uxa, const T* uya, double out[2], cudaStream_t s)
{
    Red4 r{};
    int rc = reduce_impl<T, 1>(g, uxa, uya, r, s);
    if (rc) return rc;
    out[0] = r.a;
    out[1] = r.b;
    return PLBM_OK;
}

// ---- lattice checksum ------------------------------------------------------------------------
// 64-bit checksum of the rows 0..ny-1 of one lattice (padding rows excluded): sum over all values of
// splitmix64(bits + golden * local_index), modulo 2^64.  Integer addition commutes, so the result does not depend on the
// reduction order: two lattices have the same checksum iff (up to 2^-64) they hold the same bits at the same places.
// Used by bench.py's selfcheck (multi-GPU slabs == the single-GPU result without moving 38 GB to the host) and by tests.
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
template <typename T>
__global__ void __launch_bounds__(256) k_lattice_hash(const T* __restrict__ f, int nx, int ny, int ld, unsigned long long* __restrict__ out)
{
    const size_t n = (size_t)9 * nx * ny;
    unsigned long long h = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t line = i / ny;            // q * nx + x
        const int y = (int)(i - line * ny);
        const T v = f[line * (size_t)ld + y];
        unsigned long long bits;
        if (sizeof(T) == 8) bits = (unsigned long long)__double_as_longlong((double)v);
        else bits = (unsigned long long)__float_as_uint((float)v);
        h += splitmix64(bits + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, h);
}

template <typename T> int launch_lattice_hash(Grid& g, const T* f, unsigned long long* out, cudaStream_t s)
{
    unsigned long long* dev = static_cast<unsigned long long*>(g.partial);
    PLBM_CUDA(cudaMemsetAsync(dev, 0, sizeof(unsigned long long), s));
    const size_t n = (size_t)9 * g.nx * g.ny;
    int nb = (int)((n + 255) / 256);
    if (nb > 8 * g.sm_count) nb = 8 * g.sm_count;
    k_lattice_hash<T><<<nb, 256, 0, s>>>(f, g.nx, g.ny, g.ld, dev);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PLBM_CUDA(cudaGetLastError());
    PLBM_CUDA(cudaMemcpyAsync(g.partial_host, dev, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    PLBM_CUDA(cudaStreamSynchronize(s));
    *out = *static_cast<const unsigned long long*>(g.partial_host);
    return PLBM_OK;
}
template int launch_lattice_hash<double>(Grid&, const double*, unsigned long long*, cudaStream_t);
template int launch_lattice_hash<float>(Grid&, const float*, unsigned long long*, cudaStream_t);

#define INST(T)                                                                                   \
    template int launch_init_eq<T>(const Grid&, T*, cudaStream_t);                                \
    template int launch_macros<T>(const Grid&, const T*, cudaStream_t);                           \
    template int launch_vorticity<T>(const Grid&, int, const T*, const T*, T*, cudaStream_t, const T*, const T*); \
    template int launch_diagnostics<T>(Grid&, double[PLBM_DIAG_COUNT], cudaStream_t);             \
    template int launch_l2_sums<T>(Grid&, const T*, const T*, double[2], cudaStream_t);
INST(double)
INST(float)
#undef INST

}  // namespace plbm
