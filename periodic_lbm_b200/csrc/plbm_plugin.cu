// plbm_plugin.cu -- the reference's C plugin seam (sim/lbm.h:13-19) implemented on the GPU.
//
// Exports c_plbm_{init,step,vars,free,norm} with the exact signatures of the reference's
// `slbm` plugin (sim/sim_slbm.F90:135-205), so the reference's loader
// `SimPlugin("./libplbm_b200.so")` (sim/cases.py:13-66, name = "plbm") works unmodified, plus
// c_plbm_step_n to batch steps without one ctypes round trip per step.
//
// Semantics follow sim/sim.F90: populations are stored DDF-shifted (f - w), arrays are
// x-fastest (nx,ny), the `rho` argument of init carries PRESSURE (rho = rho0 + p/cs^2,
// sim/sim.F90:181-199), one step is collide -> push-stream -> periodic fold
// (sim/sim_slbm.F90:80-110; sim/sim.F90:404-505, 568-624).  On the device the halo ring of the
// reference is replaced by periodic index arithmetic in the scatter; results are identical.
#include <cmath>
#include <cstdio>

#include "plbm_internal.h"

namespace plbm {
namespace {

struct SimState {
    int nx, ny;
    double *f1, *f2;   // [9][ny][nx], DDF-shifted
    double* f3 = nullptr;  // third buffer of the Heun finite-volume plugin (fc)
    double *rho, *u;   // device staging for vars(): rho[ny][nx], u[2][ny][nx]
    cudaStream_t stream;
    double dt;         // time step (the Lax-Wendroff plugins accept any dt, slbm requires 1)
    int order = 2;     // Lax-Wendroff plugins: 2 (lw), 4 (lw4), 6 (lw6)
};

__device__ __forceinline__ void sim_equilibrium(double rho, double ux, double uy, double (&feq)[9])
{
    const double w0 = 4.0 / 9.0, ws = 1.0 / 9.0, wd = 1.0 / 36.0, rho0 = 1.0;
    const double w[9] = {w0, ws, ws, ws, ws, wd, wd, wd, wd};
    double uxx = ux * ux, uyy = uy * uy;
    double uxpy = ux + uy, uxmy = ux - uy;
    double indp = -1.5 * (uxx + uyy);
    feq[0] = w0 * rho * (indp);
    feq[1] = ws * rho * (indp + 3.0 * ux + 4.5 * uxx);
    feq[2] = ws * rho * (indp + 3.0 * uy + 4.5 * uyy);
    feq[3] = ws * rho * (indp - 3.0 * ux + 4.5 * uxx);
    feq[4] = ws * rho * (indp - 3.0 * uy + 4.5 * uyy);
    feq[5] = wd * rho * (indp + 3.0 * uxpy + 4.5 * uxpy * uxpy);
    feq[7] = wd * rho * (indp - 3.0 * uxpy + 4.5 * uxpy * uxpy);
    feq[6] = wd * rho * (indp - 3.0 * uxmy + 4.5 * uxmy * uxmy);
    feq[8] = wd * rho * (indp + 3.0 * uxmy + 4.5 * uxmy * uxmy);
#pragma unroll
    for (int k = 0; k < 9; ++k) feq[k] = feq[k] + w[k] * (rho - rho0);
}

// lbm_eqinit, sim/sim.F90:181-199
__global__ void k_sim_init(const double* __restrict__ p, const double* __restrict__ u, double* __restrict__ f, int nx, int ny)
{
    const size_t n = (size_t)nx * ny;
    const size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    const double csqr = 1.0 / 3.0, rho0 = 1.0;
    double rho = rho0 + p[m] / csqr;
    double feq[9];
    sim_equilibrium(rho, u[m], u[n + m], feq);
#pragma unroll
    for (int k = 0; k < 9; ++k) f[k * n + m] = feq[k];
}

// lbm_collide_and_stream_fused (push) + lbm_periodic_bc_push, sim/sim.F90:404-505, 568-624
__global__ void __launch_bounds__(256) k_sim_step(const double* __restrict__ fsrc, double* __restrict__ fdst, int nx, int ny,
                                                  double omega)
{
    const size_t n = (size_t)nx * ny;
    const size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    const int j = (int)(m / nx), i = (int)(m - (size_t)j * nx);
    const double rho0 = 1.0;
    double f[9], feq[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) f[k] = fsrc[k * n + m];
    double rho = (((f[5] + f[7]) + (f[6] + f[8])) + ((f[1] + f[3]) + (f[2] + f[4]))) + f[0];
    rho = rho + rho0;
    double ux = (((f[5] - f[7]) + (f[8] - f[6])) + (f[1] - f[3])) / rho;
    double uy = (((f[5] - f[7]) + (f[6] - f[8])) + (f[2] - f[4])) / rho;
    sim_equilibrium(rho, ux, uy, feq);
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        int id = i + cxi(k), jd = j + cyi(k);
        id = id < 0 ? nx - 1 : (id >= nx ? 0 : id);
        jd = jd < 0 ? ny - 1 : (jd >= ny ? 0 : jd);
        fdst[k * n + (size_t)jd * nx + id] = omega * (feq[k] - f[k]) + f[k];
    }
}

// lbm_macros, sim/sim.F90:148-179
__global__ void k_sim_macros(const double* __restrict__ fsrc, double* __restrict__ rho, double* __restrict__ u, int nx, int ny)
{
    const size_t n = (size_t)nx * ny;
    const size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    double f[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) f[k] = fsrc[k * n + m];
    double r = (((f[5] + f[7]) + (f[6] + f[8])) + ((f[1] + f[3]) + (f[2] + f[4]))) + f[0];
    r = r + 1.0;
    rho[m] = r;
    u[m] = (((f[5] - f[7]) + (f[8] - f[6])) + (f[1] - f[3])) / r;
    u[n + m] = (((f[5] - f[7]) + (f[6] - f[8])) + (f[2] - f[4])) / r;
}

// The reference's Lax-Wendroff plugins `lw`, `lw4`, `lw6` (sim/sim_lw.F90, sim_lw4.F90, sim_lw6.F90):
// Lax-Wendroff streaming of every population with 2nd / 4th / 6th-order central differences (lw_stream
// sim_lw.F90:24-85, lw4_stream sim_lw4.F90:26-115, lw6_stream sim_lw6.F90:26-127) followed by the DDF-shifted
// BGK collision (lw_collision sim_lw.F90:87-166, identical in all three); the periodic halo copies lw*_bc
// are index arithmetic here.  One fused kernel per step.
template <int ORDER>
__global__ void __launch_bounds__(256) k_lw_step(const double* __restrict__ fsrc, double* __restrict__ fdst, int nx, int ny,
                                                 double dt, double omega)
{
    constexpr int H = ORDER / 2;
    const size_t n = (size_t)nx * ny;
    const size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    const int j = (int)(m / nx), i = (int)(m - (size_t)j * nx);
    int ix[2 * H + 1], jy[2 * H + 1];  // periodic neighbours i-H..i+H, j-H..j+H (nx, ny >= H)
#pragma unroll
    for (int d = -H; d <= H; ++d) {
        int ii = i + d, jj = j + d;
        ix[d + H] = ii < 0 ? ii + nx : (ii >= nx ? ii - nx : ii);
        jy[d + H] = jj < 0 ? jj + ny : (jj >= ny ? jj - ny : jj);
    }
    const double w0 = 4.0 / 9.0, ws = 1.0 / 9.0, wd = 1.0 / 36.0, rho0 = 1.0;
    const double w[9] = {w0, ws, ws, ws, ws, wd, wd, wd, wd};
    double f[9], feq[9];
    f[0] = fsrc[m];
#pragma unroll
    for (int k = 1; k < 9; ++k) {
        const double* fk = fsrc + k * n;
        const double vx = dt * cxi(k), vy = dt * cyi(k);
        const double vxx = 0.5 * vx * vx, vyy = 0.5 * vy * vy, vxy = vx * vy;
#define F(di, dj) fk[(size_t)jy[(dj) + H] * nx + ix[(di) + H]]
        double dfx, dfy, dfxx, dfyy, dfxy;
        if (ORDER == 2) {
            dfx = 0.5 * (F(1, 0) - F(-1, 0));
            dfy = 0.5 * (F(0, 1) - F(0, -1));
            dfxx = F(1, 0) - 2.0 * F(0, 0) + F(-1, 0);
            dfyy = F(0, 1) - 2.0 * F(0, 0) + F(0, -1);
            dfxy = 0.25 * (F(1, 1) - F(-1, 1) + F(-1, -1) - F(1, -1));
        } else if (ORDER == 4) {
            const double fc = F(0, 0), fe = F(1, 0), fw = F(-1, 0), fee = F(2, 0), fww = F(-2, 0);
            const double fn = F(0, 1), fs = F(0, -1), fnn = F(0, 2), fss = F(0, -2);
            const double fne = F(1, 1), fnw = F(-1, 1), fsw = F(-1, -1), fse = F(1, -1);
            const double fne2 = F(2, 2), fnw2 = F(-2, 2), fsw2 = F(-2, -2), fse2 = F(2, -2);
            dfx = (1.0 / 12.0) * (fww - fee) + (2.0 / 3.0) * (fe - fw);
            dfy = (1.0 / 12.0) * (fss - fnn) + (2.0 / 3.0) * (fn - fs);
            dfxx = -(1.0 / 12.0) * fww + (4.0 / 3.0) * fw - (5.0 / 2.0) * fc + (4.0 / 3.0) * fe - (1.0 / 12.0) * fee;
            dfyy = -(1.0 / 12.0) * fss + (4.0 / 3.0) * fs - (5.0 / 2.0) * fc + (4.0 / 3.0) * fn - (1.0 / 12.0) * fnn;
            dfxy = (1.0 / 3.0) * (fne - fnw + fsw - fse) - (1.0 / 48.0) * (fne2 - fnw2 + fsw2 - fse2);
        } else {
            dfx = (1.0 / 60.0) * (F(3, 0) - F(-3, 0)) - (3.0 / 20.0) * (F(2, 0) - F(-2, 0)) + (3.0 / 4.0) * (F(1, 0) - F(-1, 0));
            dfy = (1.0 / 60.0) * (F(0, 3) - F(0, -3)) - (3.0 / 20.0) * (F(0, 2) - F(0, -2)) + (3.0 / 4.0) * (F(0, 1) - F(0, -1));
            dfxx = (1.0 / 90.0) * (F(-3, 0) + F(3, 0)) - (3.0 / 20.0) * (F(-2, 0) + F(2, 0)) + (3.0 / 2.0) * (F(-1, 0) + F(1, 0)) -
                   (49.0 / 18.0) * (F(0, 0));
            dfyy = (1.0 / 90.0) * (F(0, -3) + F(0, 3)) - (3.0 / 20.0) * (F(0, -2) + F(0, 2)) + (3.0 / 2.0) * (F(0, -1) + F(0, 1)) -
                   (49.0 / 18.0) * (F(0, 0));
            dfxy = (3.0 / 8.0) * (F(1, 1) - F(-1, 1) + F(-1, -1) - F(1, -1)) -
                   (3.0 / 80.0) * (F(2, 2) - F(-2, 2) + F(-2, -2) - F(2, -2)) +
                   (1.0 / 360.0) * (F(3, 3) - F(-3, 3) + F(-3, -3) - F(3, -3));
        }
        f[k] = F(0, 0) - vx * dfx - vy * dfy + (vxx * dfxx + vxy * dfxy + vyy * dfyy);
#undef F
    }
    double rho = f[0] + (((f[5] + f[7]) + (f[6] + f[8])) + ((f[1] + f[3]) + (f[2] + f[4]))) + rho0;
    double irho = 1.0 / rho;
    double ux = (((f[5] - f[7]) + (f[8] - f[6])) + (f[1] - f[3])) * irho;
    double uy = (((f[5] - f[7]) + (f[6] - f[8])) + (f[2] - f[4])) * irho;
    double uxx = ux * ux, uyy = uy * uy;
    double indp = -1.5 * (uxx + uyy);
    feq[0] = w0 * rho * (indp);
    feq[1] = ws * rho * (indp + 3.0 * ux + 4.5 * uxx);
    feq[2] = ws * rho * (indp + 3.0 * uy + 4.5 * uyy);
    feq[3] = ws * rho * (indp - 3.0 * ux + 4.5 * uxx);
    feq[4] = ws * rho * (indp - 3.0 * uy + 4.5 * uyy);
    double uxpy = ux + uy;
    feq[5] = wd * rho * (indp + 3.0 * uxpy + 4.5 * uxpy * uxpy);
    feq[7] = wd * rho * (indp - 3.0 * uxpy + 4.5 * uxpy * uxpy);
    double uxmy = ux - uy;
    feq[6] = wd * rho * (indp - 3.0 * uxmy + 4.5 * uxmy * uxmy);
    feq[8] = wd * rho * (indp + 3.0 * uxmy + 4.5 * uxmy * uxmy);
#pragma unroll
    for (int k = 0; k < 9; ++k) feq[k] = feq[k] + w[k] * (rho - rho0);
#pragma unroll
    for (int k = 0; k < 9; ++k) fdst[k * n + m] = f[k] + omega * (feq[k] - f[k]);
}

// ---- Heun finite-volume plugin `fvm` (sim/sim_fvm.F90) ------------------------------------------------------
// One step (sim_fvm%step, :282-322) on the periodic grid = three kernels; the reference's halo copies
// (fvm_bc :191-218) and its collisions of the halo layer are periodic index arithmetic here.
//   k_fvm_collide2: fc = collide(f1, omega), f1 = collide(f1, omega/2)                       (:297-302)
//   k_fvm_predict : f2 = collide(fc - dt flux(f1) [q > 0], f1 [q = 0]; omega/2)             (:304-310, 67-100)
//   k_fvm_correct : fc = fc - dt/2 (flux(f1) + flux(f2))  [q > 0]                             (:311-312, 103-136)
__device__ __forceinline__ void fvm_collide(double (&f)[9], double omega)
{
    double feq[9];
    double rho = f[0] + (((f[5] + f[7]) + (f[6] + f[8])) + ((f[1] + f[3]) + (f[2] + f[4]))) + 1.0;
    double ux = (((f[5] - f[7]) + (f[8] - f[6])) + (f[1] - f[3])) / rho;
    double uy = (((f[5] - f[7]) + (f[6] - f[8])) + (f[2] - f[4])) / rho;
    sim_equilibrium(rho, ux, uy, feq);
#pragma unroll
    for (int k = 0; k < 9; ++k) f[k] = f[k] + omega * (feq[k] - f[k]);
}

__global__ void __launch_bounds__(256) k_fvm_collide2(double* __restrict__ f1, double* __restrict__ fc, int nx, int ny, double omega)
{
    const size_t n = (size_t)nx * ny;
    const size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    double a[9], b[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) a[k] = b[k] = f1[k * n + m];
    fvm_collide(a, omega);
    fvm_collide(b, 0.5 * omega);
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        fc[k * n + m] = a[k];
        f1[k * n + m] = b[k];
    }
}

// central flux of population k at (i, j), sim/sim_fvm.F90:91-92
template <int K> __device__ __forceinline__ double fvm_flux(const double* __restrict__ fk, int nx, int i, int j, int ip, int im, int jp, int jm)
{
    const double c = fk[(size_t)j * nx + i];
    const double cx = (double)cxi(K), cy = (double)cyi(K);
    return cx * (0.5 * (fk[(size_t)j * nx + ip] + c) - 0.5 * (c + fk[(size_t)j * nx + im])) +
           cy * (0.5 * (fk[(size_t)jp * nx + i] + c) - 0.5 * (c + fk[(size_t)jm * nx + i]));
}

__global__ void __launch_bounds__(256) k_fvm_predict(const double* __restrict__ f1, const double* __restrict__ fc, double* __restrict__ f2,
                                                     int nx, int ny, double dt, double omega)
{
    const size_t n = (size_t)nx * ny;
    const size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    const int j = (int)(m / nx), i = (int)(m - (size_t)j * nx);
    const int ip = i + 1 == nx ? 0 : i + 1, im = i == 0 ? nx - 1 : i - 1;
    const int jp = j + 1 == ny ? 0 : j + 1, jm = j == 0 ? ny - 1 : j - 1;
    double g[9];
    g[0] = f1[m];
#define PLBM_P(K) g[K] = fc[K * n + m] - dt * fvm_flux<K>(f1 + K * n, nx, i, j, ip, im, jp, jm)
    PLBM_P(1);
    PLBM_P(2);
    PLBM_P(3);
    PLBM_P(4);
    PLBM_P(5);
    PLBM_P(6);
    PLBM_P(7);
    PLBM_P(8);
#undef PLBM_P
    fvm_collide(g, 0.5 * omega);
#pragma unroll
    for (int k = 0; k < 9; ++k) f2[k * n + m] = g[k];
}

__global__ void __launch_bounds__(256) k_fvm_correct(const double* __restrict__ f1, const double* __restrict__ f2, double* __restrict__ fc,
                                                     int nx, int ny, double dt)
{
    const size_t n = (size_t)nx * ny;
    const size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    const int j = (int)(m / nx), i = (int)(m - (size_t)j * nx);
    const int ip = i + 1 == nx ? 0 : i + 1, im = i == 0 ? nx - 1 : i - 1;
    const int jp = j + 1 == ny ? 0 : j + 1, jm = j == 0 ? ny - 1 : j - 1;
#define PLBM_C(K)                                                                         \
    {                                                                                     \
        const double flux = fvm_flux<K>(f1 + K * n, nx, i, j, ip, im, jp, jm);            \
        const double fluxp = fvm_flux<K>(f2 + K * n, nx, i, j, ip, im, jp, jm);           \
        fc[K * n + m] = fc[K * n + m] - dt * 0.5 * (flux + fluxp);                        \
    }
    PLBM_C(1) PLBM_C(2) PLBM_C(3) PLBM_C(4) PLBM_C(5) PLBM_C(6) PLBM_C(7) PLBM_C(8)
#undef PLBM_C
}

void sim_destroy(SimState* s)
{
    if (!s) return;
    if (s->f1) cudaFree(s->f1);
    if (s->f2) cudaFree(s->f2);
    if (s->f3) cudaFree(s->f3);
    if (s->rho) cudaFree(s->rho);
    if (s->u) cudaFree(s->u);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

}  // namespace
}  // namespace plbm

using namespace plbm;

extern "C" {

// void *siminit(int nx, int ny, double dt, double *rho, double *u, double *sigma, void *params)
// Returns NULL on failure (the reference `error stop`s); see plbm_last_error().
static void* sim_init_common(int nx, int ny, double dt, const double* rho, const double* u, const double* sigma, bool need_unit_dt)
{
    if (nx < 1 || ny < 1 || !rho || !u) {
        set_error("c_plbm_init: bad argument");
        return nullptr;
    }
    if (need_unit_dt && dt != 1.0) {  // sim/sim_slbm.F90:48-51
        set_error("Standard LBM only supports dt = 1.0!");
        return nullptr;
    }
    if (sigma) {  // sim/sim_slbm.F90:65-70
        set_error("c_plbm_init: non-equilibrium initialisation is not implemented (NotImplementedError in the reference)");
        return nullptr;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device available: libplbm_b200 has no CPU fallback");
        return nullptr;
    }
    SimState* s = new SimState();
    s->nx = nx;
    s->ny = ny;
    s->f1 = s->f2 = s->rho = s->u = nullptr;
    s->stream = nullptr;
    s->dt = dt;
    const size_t n = (size_t)nx * ny;
    bool ok = cudaMalloc(&s->f1, 9 * n * sizeof(double)) == cudaSuccess && cudaMalloc(&s->f2, 9 * n * sizeof(double)) == cudaSuccess &&
              cudaMalloc(&s->rho, n * sizeof(double)) == cudaSuccess && cudaMalloc(&s->u, 2 * n * sizeof(double)) == cudaSuccess &&
              cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(s->rho, rho, n * sizeof(double), cudaMemcpyHostToDevice, s->stream) == cudaSuccess &&
         cudaMemcpyAsync(s->u, u, 2 * n * sizeof(double), cudaMemcpyHostToDevice, s->stream) == cudaSuccess;
    if (ok) {
        k_sim_init<<<(unsigned)((n + 255) / 256), 256, 0, s->stream>>>(s->rho, s->u, s->f1, nx, ny);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        ok = cudaStreamSynchronize(s->stream) == cudaSuccess;
    }
    if (!ok) {
        cuda_fail(cudaGetLastError(), "c_plbm_init");
        sim_destroy(s);
        return nullptr;
    }
    return s;
}

void* c_plbm_init(int nx, int ny, double dt, const double* rho, const double* u, const double* sigma, void* params)
{
    (void)params;
    return sim_init_common(nx, ny, dt, rho, u, sigma, true);
}

void c_plbm_step_n(void* sim, double omega, int n)
{
    SimState* s = static_cast<SimState*>(sim);
    if (!s) return;
    const size_t nn = (size_t)s->nx * s->ny;
    for (int it = 0; it < n; ++it) {
        k_sim_step<<<(unsigned)((nn + 255) / 256), 256, 0, s->stream>>>(s->f1, s->f2, s->nx, s->ny, omega);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        double* t = s->f1;
        s->f1 = s->f2;
        s->f2 = t;
    }
}

// void simstep(void *sim, double omega)
void c_plbm_step(void* sim, double omega) { c_plbm_step_n(sim, omega, 1); }

// void simvars(void *sim, double *rho, double *u)
void c_plbm_vars(void* sim, double* rho, double* u)
{
    SimState* s = static_cast<SimState*>(sim);
    if (!s || !rho || !u) return;
    const size_t n = (size_t)s->nx * s->ny;
    k_sim_macros<<<(unsigned)((n + 255) / 256), 256, 0, s->stream>>>(s->f1, s->rho, s->u, s->nx, s->ny);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaMemcpyAsync(rho, s->rho, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream);
    cudaMemcpyAsync(u, s->u, 2 * n * sizeof(double), cudaMemcpyDeviceToHost, s->stream);
    if (cudaStreamSynchronize(s->stream) != cudaSuccess) cuda_fail(cudaGetLastError(), "c_plbm_vars");
}

// void simfree(void *sim)
void c_plbm_free(void* sim) { sim_destroy(static_cast<SimState*>(sim)); }

// c_slbm_norm (sim/sim_slbm.F90:198-205): norm2(u - ua) / norm2(ua) over nx*ny host values.
// Host arrays in, scalar out: evaluated on the host like the reference does.
double c_plbm_norm(int nx, int ny, const double* u, const double* ua)
{
    const size_t n = (size_t)nx * ny;
    double a = 0.0, b = 0.0;
    for (size_t i = 0; i < n; ++i) {
        a += (u[i] - ua[i]) * (u[i] - ua[i]);
        b += ua[i] * ua[i];
    }
    return std::sqrt(a) / std::sqrt(b);
}


// The same plugin under the reference's own name: the loader derives the symbol prefix from the file
// name (sim/cases.py:27-39, lib<name>.so -> c_<name>_*), so `ln -s libplbm_b200.so libslbm.so` makes this
// library a drop-in for the reference's libslbm.so in sim/cases.py and sim/standard_lbm.F90.
void* c_slbm_init(int nx, int ny, double dt, const double* rho, const double* u, const double* sigma, void* params)
{
    return c_plbm_init(nx, ny, dt, rho, u, sigma, params);
}
void c_slbm_step(void* sim, double omega) { c_plbm_step_n(sim, omega, 1); }
void c_slbm_vars(void* sim, double* rho, double* u) { c_plbm_vars(sim, rho, u); }
void c_slbm_free(void* sim) { c_plbm_free(sim); }
double c_slbm_norm(int nx, int ny, const double* u, const double* ua) { return c_plbm_norm(nx, ny, u, ua); }


// The reference's Lax-Wendroff plugin `lw` (sim/sim_lw.F90:295-425) under its own symbol names:
// `ln -s libplbm_b200.so liblw.so` is a drop-in for the reference's liblw.so.  Any dt is accepted.
static void* lw_init_order(int order, int nx, int ny, double dt, const double* rho, const double* u, const double* sigma)
{
    if (nx < order / 2 || ny < order / 2) {  // the reference's halo copies need nx, ny >= nhalo
        set_error("c_lw*_init: grid smaller than the stencil halo");
        return nullptr;
    }
    SimState* s = static_cast<SimState*>(sim_init_common(nx, ny, dt, rho, u, sigma, false));
    if (s) s->order = order;
    return s;
}
void* c_lw_init(int nx, int ny, double dt, const double* rho, const double* u, const double* sigma, void* params)
{
    (void)params;
    return lw_init_order(2, nx, ny, dt, rho, u, sigma);
}
void c_lw_step_n(void* sim, double omega, int n)
{
    SimState* s = static_cast<SimState*>(sim);
    if (!s) return;
    const size_t nn = (size_t)s->nx * s->ny;
    const unsigned nb = (unsigned)((nn + 255) / 256);
    for (int it = 0; it < n; ++it) {
        if (s->order == 6)
            k_lw_step<6><<<nb, 256, 0, s->stream>>>(s->f1, s->f2, s->nx, s->ny, s->dt, omega);
        else if (s->order == 4)
            k_lw_step<4><<<nb, 256, 0, s->stream>>>(s->f1, s->f2, s->nx, s->ny, s->dt, omega);
        else
            k_lw_step<2><<<nb, 256, 0, s->stream>>>(s->f1, s->f2, s->nx, s->ny, s->dt, omega);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        double* t = s->f1;
        s->f1 = s->f2;
        s->f2 = t;
    }
}
void c_lw_step(void* sim, double omega) { c_lw_step_n(sim, omega, 1); }
void c_lw_vars(void* sim, double* rho, double* u) { c_plbm_vars(sim, rho, u); }
void c_lw_free(void* sim) { c_plbm_free(sim); }
double c_lw_norm(int nx, int ny, const double* u, const double* ua) { return c_plbm_norm(nx, ny, u, ua); }

// `lw4` (sim/sim_lw4.F90:370-477) and `lw6` (sim/sim_lw6.F90) under their own symbol names: the step, vars,
// free and norm entries are shared with `lw` (the order lives in the handle).
void* c_lw4_init(int nx, int ny, double dt, const double* rho, const double* u, const double* sigma, void* params)
{
    (void)params;
    return lw_init_order(4, nx, ny, dt, rho, u, sigma);
}
void c_lw4_step_n(void* sim, double omega, int n) { c_lw_step_n(sim, omega, n); }
void c_lw4_step(void* sim, double omega) { c_lw_step_n(sim, omega, 1); }
void c_lw4_vars(void* sim, double* rho, double* u) { c_plbm_vars(sim, rho, u); }
void c_lw4_free(void* sim) { c_plbm_free(sim); }
double c_lw4_norm(int nx, int ny, const double* u, const double* ua) { return c_plbm_norm(nx, ny, u, ua); }

void* c_lw6_init(int nx, int ny, double dt, const double* rho, const double* u, const double* sigma, void* params)
{
    (void)params;
    return lw_init_order(6, nx, ny, dt, rho, u, sigma);
}
void c_lw6_step_n(void* sim, double omega, int n) { c_lw_step_n(sim, omega, n); }
void c_lw6_step(void* sim, double omega) { c_lw_step_n(sim, omega, 1); }
void c_lw6_vars(void* sim, double* rho, double* u) { c_plbm_vars(sim, rho, u); }
void c_lw6_free(void* sim) { c_plbm_free(sim); }
double c_lw6_norm(int nx, int ny, const double* u, const double* ua) { return c_plbm_norm(nx, ny, u, ua); }

// The Heun finite-volume plugin `fvm` (sim/sim_fvm.F90:340-432) under its own symbol names (libfvm.so drop-in).
void* c_fvm_init(int nx, int ny, double dt, const double* rho, const double* u, const double* sigma, void* params)
{
    (void)params;
    if (nx < 2 || ny < 2) {
        set_error("c_fvm_init: grid smaller than the stencil halo");
        return nullptr;
    }
    SimState* s = static_cast<SimState*>(sim_init_common(nx, ny, dt, rho, u, sigma, false));
    if (s && cudaMalloc(&s->f3, 9 * (size_t)nx * ny * sizeof(double)) != cudaSuccess) {
        cuda_fail(cudaGetLastError(), "c_fvm_init");
        sim_destroy(s);
        return nullptr;
    }
    return s;
}
void c_fvm_step_n(void* sim, double omega, int n)
{
    SimState* s = static_cast<SimState*>(sim);
    if (!s || !s->f3) return;
    const size_t nn = (size_t)s->nx * s->ny;
    const unsigned nb = (unsigned)((nn + 255) / 256);
    for (int it = 0; it < n; ++it) {
        k_fvm_collide2<<<nb, 256, 0, s->stream>>>(s->f1, s->f3, s->nx, s->ny, omega);
        k_fvm_predict<<<nb, 256, 0, s->stream>>>(s->f1, s->f3, s->f2, s->nx, s->ny, s->dt, omega);
        k_fvm_correct<<<nb, 256, 0, s->stream>>>(s->f1, s->f2, s->f3, s->nx, s->ny, s->dt);
        g_launches.fetch_add(3, std::memory_order_relaxed);
        double* t = s->f1;  // move_alloc swap of f1 and fc
        s->f1 = s->f3;
        s->f3 = t;
    }
}
void c_fvm_step(void* sim, double omega) { c_fvm_step_n(sim, omega, 1); }
void c_fvm_vars(void* sim, double* rho, double* u) { c_plbm_vars(sim, rho, u); }
void c_fvm_free(void* sim) { c_plbm_free(sim); }
double c_fvm_norm(int nx, int ny, const double* u, const double* ua) { return c_plbm_norm(nx, ny, u, ua); }

}  // extern "C"
