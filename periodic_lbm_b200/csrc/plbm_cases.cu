// plbm_cases.cu -- host-side flow cases: the analytic input generators the reference drivers
// call once per run / per error evaluation (O(N), not on the hot path):
//   taylor_green_t%eval, %decay_time   src/benchmarks/taylor_green.f90:31-84
//   vortex_case_t%eval                 src/benchmarks/barotropic_vortex_case.F90:34-83
// They fill caller-owned host arrays (ny,nx) in the working precision with the same expression
// order and the same libm the Fortran intrinsics resolve to, so a driver switching to this
// library gets bit-identical initial conditions.
#include <cmath>
#include <thread>
#include <vector>

#include "plbm_internal.h"

namespace plbm {
namespace {

template <typename T> T tg_decay_time(T kx, T ky, T nu) { return T(1) / (nu * (kx * kx + ky * ky)); }

// lines [xa, xb) of a slab whose first line is global line x0 (x0 = 0 for the whole grid)
template <typename T> void tg_eval_range(int xa, int xb, int x0, int ny, T kx, T ky, T umax, T td, T t, T* p, T* ux, T* uy)
{
    for (int x = xa; x < xb; ++x) {
        const T xx = T(x0 + x) + T(0.5);
        for (int y = 0; y < ny; ++y) {
            const T yy = T(y) + T(0.5);
            const size_t m = (size_t)x * ny + y;
            ux[m] = -umax * std::sqrt(ky / kx) * std::cos(kx * xx) * std::sin(ky * yy) * std::exp(-t / td);
            uy[m] = umax * std::sqrt(kx / ky) * std::sin(kx * xx) * std::cos(ky * yy) * std::exp(-t / td);
            p[m] = -T(0.25) * (umax * umax) * ((ky / kx) * std::cos(T(2) * kx * xx) + (kx / ky) * std::cos(T(2) * ky * yy)) *
                   std::exp(-T(2) * t / td);
        }
    }
}

// the evaluation is embarrassingly parallel over lines: split it over the host cores
template <typename T> void tg_eval(int nx, int x0, int ny, T kx, T ky, T umax, T td, T t, T* p, T* ux, T* uy)
{
    unsigned nt = std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if ((size_t)nx * ny < (1u << 16) || nt == 1) return tg_eval_range<T>(0, nx, x0, ny, kx, ky, umax, td, t, p, ux, uy);
    if (nt > (unsigned)nx) nt = nx;
    std::vector<std::thread> pool;
    for (unsigned i = 0; i < nt; ++i) {
        const int xa = (int)((size_t)nx * i / nt), xb = (int)((size_t)nx * (i + 1) / nt);
        pool.emplace_back([=] { tg_eval_range<T>(xa, xb, x0, ny, kx, ky, umax, td, t, p, ux, uy); });
    }
    for (auto& th : pool) th.join();
}

template <typename T> void vortex_eval(int nx, int ny, T U0, T xc, T yc, T Rc, T eps, T rho0, T csqr, T* rho, T* ux, T* uy)
{
    const T half = T(1) / T(2);
    const T Rcsqr = Rc * Rc;
    const T vMa_sq = eps * eps / csqr;
    for (int x = 0; x < nx; ++x) {
        const T xl = T(x) + T(0.5);
        for (int y = 0; y < ny; ++y) {
            const T yl = T(y) + T(0.5);
            const T xr = xl - xc, yr = yl - yc;
            const T rsqr = xr * xr + yr * yr;
            const size_t m = (size_t)x * ny + y;
            ux[m] = U0 - eps * (yr / Rc) * std::exp(-half * rsqr / Rcsqr);
            uy[m] = eps * (xr / Rc) * std::exp(-half * rsqr / Rcsqr);
            rho[m] = rho0 * std::exp(-half * vMa_sq * std::exp(-rsqr / Rcsqr));
        }
    }
}

}  // namespace
}  // namespace plbm

using namespace plbm;

extern "C" {

double plbm_case_tg_decay_time(int precision, double kx, double ky, double nu)
{
    return precision == PLBM_F64 ? tg_decay_time<double>(kx, ky, nu) : (double)tg_decay_time<float>((float)kx, (float)ky, (float)nu);
}

int plbm_case_taylor_green_slab(int precision, int nx, int x0, int ny, double kx, double ky, double umax, double td, double t, void* p,
                                void* ux, void* uy)
{
    if (nx < 1 || ny < 1 || x0 < 0 || !p || !ux || !uy) {
        set_error("case_taylor_green: bad argument");
        return PLBM_ERR_ARG;
    }
    if (precision == PLBM_F64)
        tg_eval<double>(nx, x0, ny, kx, ky, umax, td, t, (double*)p, (double*)ux, (double*)uy);
    else
        tg_eval<float>(nx, x0, ny, (float)kx, (float)ky, (float)umax, (float)td, (float)t, (float*)p, (float*)ux, (float*)uy);
    return PLBM_OK;
}

int plbm_case_taylor_green(int precision, int nx, int ny, double kx, double ky, double umax, double td, double t, void* p, void* ux,
                           void* uy)
{
    return plbm_case_taylor_green_slab(precision, nx, 0, ny, kx, ky, umax, td, t, p, ux, uy);
}

int plbm_case_vortex(int precision, int nx, int ny, double U0, double xc, double yc, double Rc, double eps, double rho0, double csqr,
                     void* rho, void* ux, void* uy)
{
    if (nx < 1 || ny < 1 || !rho || !ux || !uy) {
        set_error("case_vortex: bad argument");
        return PLBM_ERR_ARG;
    }
    if (precision == PLBM_F64)
        vortex_eval<double>(nx, ny, U0, xc, yc, Rc, eps, rho0, csqr, (double*)rho, (double*)ux, (double*)uy);
    else
        vortex_eval<float>(nx, ny, (float)U0, (float)xc, (float)yc, (float)Rc, (float)eps, (float)rho0, (float)csqr, (float*)rho,
                           (float*)ux, (float*)uy);
    return PLBM_OK;
}

}  // extern "C"
