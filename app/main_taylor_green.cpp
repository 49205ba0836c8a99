// main_taylor_green.cpp -- C++ driver over the C ABI with the structure of the reference's
// app/main_taylor_green.f90 (parameter derivation :44-89, initial condition :133-149, time loop
// :98-119, MLUPS :121-122, L2 norm :174-212), taking what the Fortran driver hard-codes
//  at run time:
//
//   main_taylor_green <n> <scheme: lbm|fvm|dugks> <collision: bgk|trt|rr> <dt | r=dt/tau with 'r' prefix> [f32]
//
// e.g.  main_taylor_green 64 dugks bgk r50      -> graphs/fvm_dugks_64.txt row 50: 2.2824885E-02
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <string>
#include <vector>

#include "../include/plbm.h"

#define CHECK(call)                                                              \
    do {                                                                         \
        if ((call) != 0) {                                                       \
            std::fprintf(stderr, "%s failed: %s\n", #call, plbm_last_error());   \
            return 1;                                                            \
        }                                                                        \
    } while (0)

template <typename T> static int run(int n, const std::string& scheme, int collision, const char* dtarg, int prec)
{
    const int nx = n, ny = n;
    plbm_handle grid;
    CHECK(plbm_alloc_grid(&grid, nx, ny, 2, prec));

    const T umax = T(0.01) / std::sqrt(T(3));
    const T nu = (umax * T(nx)) / T(100);
    const T tau = T(3) * nu;
    const T dt = dtarg[0] == 'r' ? T(std::atof(dtarg + 1)) * tau : T(std::atof(dtarg));
    CHECK(plbm_set_properties(grid, nu, dt, 0.25, 1));
    double props[6];
    CHECK(plbm_get_properties(grid, props));
    std::printf(" tau = %.17g\n dt/tau = %.17g\n omega = %.17g\n", (double)tau, (double)(dt / tau), props[3]);

    const T pi = T(4) * std::atan(T(1));
    const T kx = T(2) * pi / T(nx), ky = T(2) * pi / T(ny);
    const T td = (T)plbm_case_tg_decay_time(prec, kx, ky, nu);
    const T tmax = std::log(T(2)) * td;
    const long nsteps = (long)(T(1.1) * tmax / dt);
    std::printf(" umax = %.17g\n tc   = %.17g\n nsteps = %ld\n", (double)umax, (double)td, nsteps);

    const size_t N = (size_t)nx * ny;
    std::vector<T> rho(N), ux(N), uy(N), pa(N), uxa(N), uya(N);
    CHECK(plbm_case_taylor_green(prec, nx, ny, kx, ky, umax, td, 0.0, rho.data(), ux.data(), uy.data()));
    const T csqr = (T)props[5];
    for (size_t i = 0; i < N; ++i) rho[i] = rho[i] / csqr + T(1);  // pressure -> lattice density
    CHECK(plbm_set_pdf_to_equilibrium(grid, rho.data(), ux.data(), uy.data()));

    // the drivers accumulate t by repeated addition and stop at the first t >= tmax
    T t = 0;
    long step = 0, last = 0;
    for (step = 1; step <= nsteps; ++step) {
        t = t + dt;
        if (t >= tmax) break;
    }
    last = step > nsteps ? nsteps : step;
    const auto t0 = std::chrono::steady_clock::now();
    const int chunk = 1 << 20;
    for (long done = 0; done < last;) {
        const int k = (int)std::min<long>(chunk, last - done);
        if (scheme == "lbm")
            CHECK(plbm_perform_lbm_step(grid, collision, k));
        else if (scheme == "fvm")
            CHECK(plbm_perform_step(grid, PLBM_STREAM_FVM_BARDOW, collision, k));
        else
            CHECK(plbm_perform_dugks_step(grid, 1, k));
        done += k;
    }
    CHECK(plbm_update_macros(grid, rho.data(), ux.data(), uy.data(), 1));  // lagged like the reference (F3)
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf(" MLUPS %.3f\n", (double)nx * ny * last * 1e-6 / secs);

    CHECK(plbm_case_taylor_green(prec, nx, ny, kx, ky, umax, td, t, pa.data(), uxa.data(), uya.data()));
    double sums[2], diag[PLBM_DIAG_COUNT];
    CHECK(plbm_l2_sums(grid, uxa.data(), uya.data(), sums));
    CHECK(plbm_diagnostics(grid, diag));
    std::printf(" max|u| = %.10e  min|u| = %.10e\n", diag[PLBM_DIAG_MAX_SPEED], diag[PLBM_DIAG_MIN_SPEED]);
    std::printf(" L2-norm = %.8E\n Final time = %.17g\n", std::sqrt(sums[0] / sums[1]), (double)t);
    CHECK(plbm_dealloc_grid(grid));
    return 0;
}

int main(int argc, char** argv)
{
    if (argc < 5) {
        std::fprintf(stderr, "usage: %s <n> <lbm|fvm|dugks> <bgk|trt|rr> <dt | r<dt/tau>> [f32]\n", argv[0]);
        return 2;
    }
    const int n = std::atoi(argv[1]);
    const std::string scheme = argv[2], coll = argv[3];
    const int collision = coll == "bgk" ? PLBM_BGK : coll == "trt" ? PLBM_TRT : PLBM_RR;
    const bool f32 = argc > 5 && std::strcmp(argv[5], "f32") == 0;
    return f32 ? run<float>(n, scheme, collision, argv[4], PLBM_F32) : run<double>(n, scheme, collision, argv[4], PLBM_F64);
}
