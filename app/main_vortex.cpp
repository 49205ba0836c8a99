// main_vortex.cpp -- the reference driver app/main_vortex.f90 (convected barotropic vortex, Wissocq et
// al. 2020) on top of the C++ mirror of its modules (include/plbm_grid.hpp).
//
//   main_vortex <dt> [n=128] [scheme=fvm|lbm|fdm|sofonea] [collision=bgk|trt|rr] [max_steps]
//
// Reference defaults: 128 x 128, collide_bgk + stream_fvm_bardow stepped by perform_step (:19,43-44,114).
#include <chrono>
#include <cstdlib>
#include <string>

#include "../include/plbm_grid.hpp"

using namespace plbm;

static wp dt;

static void my_logger(const lattice_grid& grid, int step)  // :156-163
{
    double d[PLBM_DIAG_COUNT];
    check(plbm_diagnostics(grid.dev, d), "diagnostics");
    if (grid.logunit) {
        std::fprintf(grid.logunit, " %d %.17g %.17g\n", step, (double)(step * dt), d[PLBM_DIAG_MAX_SPEED]);
        std::fflush(grid.logunit);
    }
}

int main(int argc, char** argv)
{
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s <dt> [n] [fvm|lbm|fdm|sofonea] [bgk|trt|rr] [max_steps]\n", argv[0]);
        return 2;
    }
    const int nx = argc > 2 ? std::atoi(argv[2]) : 128, ny = nx;
    const std::string scheme = argc > 3 ? argv[3] : "fvm", coll = argc > 4 ? argv[4] : "bgk";
    const long max_steps = argc > 5 ? std::atol(argv[5]) : -1;
    const int nprint = 10000;
    try {
        lattice_grid grid;
        alloc_grid(grid, nx, ny, 2);
        grid.filename = "results";
        set_output_folder(grid, "vortex");  // :41
        grid.collision = coll == "trt" ? collide_trt : coll == "rr" ? collide_rr : collide_bgk;
        grid.streaming = scheme == "lbm" ? lbm_stream : scheme == "fdm" ? stream_fdm_bardow
                       : scheme == "sofonea" ? stream_fdm_sofonea : stream_fvm_bardow;
        grid.logger = my_logger;

        const wp U0 = wp(0.1) / std::sqrt(wp(3));      // advective Mach number * cs
        const wp kappa = wp(0.2) / std::sqrt(wp(3));   // vortex Mach number * cs
        const wp nu = wp(0.00001);
        std::printf(" Reynolds =  %.17g\n", (double)(kappa * wp(nx) / nu));
        const wp tau = wp(0.0001);                     // as printed by the reference (:61-62)
        const wp tmax = (wp(nx) / U0) * wp(4);
        dt = wp(std::atof(argv[1]));
        std::printf(" tau =  %.17g\n dt/tau =  %.17g\n cfl =  %.17g\n", (double)tau, (double)(dt / tau), (double)dt);
        set_properties(grid, nu, dt, wp(1) / wp(4));
        std::printf(" omega =  %.17g\n", (double)grid.omega);

        vortex_case_t vcase;
        vcase.U0 = U0;
        vcase.xc = wp(nx) / wp(2);
        vcase.yc = wp(ny) / wp(2);
        vcase.Rc = wp(nx) / wp(10);
        vcase.eps = kappa;
        long nsteps = (long)(tmax / dt) + 1;
        if (max_steps >= 0 && max_steps < nsteps) nsteps = max_steps;
        std::printf(" nsteps =  %ld\n", nsteps);

        wp t = 0;
        vcase.eval(grid.nx, grid.ny, grid.rho, grid.ux, grid.uy);  // apply_initial_condition (:144-154)
        set_pdf_to_equilibrium(grid);
        output_gnuplot(grid, 0);  // :104-106
        output_npy(grid, 0);
        output_vtk(grid, 0);
        grid.logger(grid, 0);

        const auto sbegin = std::chrono::steady_clock::now();
        long step = 0;
        while (step < nsteps && t < tmax) {
            long batch = 0;
            while (step + batch < nsteps && t < tmax) {
                t = t + dt;
                ++batch;
                if ((step + batch) % nprint == 0) break;
            }
            perform_step(grid, (int)batch);
            step += batch;
            if (step % nprint == 0) {
                std::printf(" step =  %ld\n", step);
                update_macros(grid);
                output_gnuplot(grid, (int)step);  // :122-124
                output_npy(grid, (int)step);
                output_vtk(grid, (int)step);
                grid.logger(grid, (int)step);
            }
        }
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - sbegin).count();
        check(plbm_synchronize(grid.dev), "synchronize");
        std::printf(" MLUPS  %.3f\n", (double)nx * ny * (double)step * 1e-6 / secs);
        std::printf(" Final time =  %.17g\n", (double)t);
        update_macros(grid);
        double d[PLBM_DIAG_COUNT];
        check(plbm_diagnostics(grid.dev, d), "diagnostics");
        std::printf(" max|u| =  %.10e  sum(rho) =  %.12e\n", d[PLBM_DIAG_MAX_SPEED], d[PLBM_DIAG_SUM_RHO]);
        dealloc_grid(grid);
    } catch (const plbm::error& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
