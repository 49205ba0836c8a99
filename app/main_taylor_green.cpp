// main_taylor_green.cpp -- the reference driver app/main_taylor_green.f90 on top of the C++ mirror of
// its modules (include/plbm_grid.hpp).  Same structure, same names, same prints.  What the Fortran
// program hard-codes can be overridden on the command line:
//
//   main_taylor_green <dt | r<dt/tau>> [n=320] [scheme=lbm|fvm|fdm|sofonea|dugks] [collision=rr|bgk|trt]
//
// The first argument is the reference's own CLI (the time step, :54-57); `r50` selects dt = 50*tau, the
// variant the author used for graphs/fvm_*_64.txt (:60-62).  Example:
//   main_taylor_green r50 64 dugks      ->  L2-norm =  2.28248848E-02  (graphs/fvm_dugks_64.txt)
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../include/plbm_grid.hpp"

using namespace plbm;

static taylor_green_t tg;
static wp dt;

static void my_logger(const lattice_grid& grid, int step)  // :151-158
{
    double d[PLBM_DIAG_COUNT];
    check(plbm_diagnostics(grid.dev, d), "diagnostics");  // maxval(hypot(grid%ux,grid%uy)) on the device
    if (grid.logunit) {
        std::fprintf(grid.logunit, " %d %.17g %.17g\n", step, (double)(step * dt), d[PLBM_DIAG_MAX_SPEED]);
        std::fflush(grid.logunit);
    }
}

static void apply_initial_condition(const taylor_green_t& c, lattice_grid& grid)  // :133-149
{
    const wp rho0 = 1;
    c.eval(wp(0), grid.rho, grid.ux, grid.uy);
    for (size_t i = 0; i < grid.size(); ++i) grid.rho[i] = grid.rho[i] / grid.csqr + rho0;  // pressure -> lattice density
    set_pdf_to_equilibrium(grid);
}

static wp calc_L2_norm(const taylor_green_t& c, lattice_grid& grid, wp t)  // :174-212
{
    std::vector<wp> pa(grid.size()), uxa(grid.size()), uya(grid.size());
    c.eval(t, pa.data(), uxa.data(), uya.data());
    double s[2];
    check(plbm_l2_sums(grid.dev, uxa.data(), uya.data(), s), "l2_sums");  // norm2(hypot(..))^2, reduced on the device
    return (wp)std::sqrt(s[0] / s[1]);
}

int main(int argc, char** argv)
{
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s <dt | r<dt/tau>> [n] [lbm|fvm|fdm|sofonea|dugks] [rr|bgk|trt]\n", argv[0]);
        return 2;
    }
    const int nx = argc > 2 ? std::atoi(argv[2]) : 320, ny = nx;  // :16
    const std::string scheme = argc > 3 ? argv[3] : "lbm", coll = argc > 4 ? argv[4] : "rr";
    const int nprint = 20000;
    try {
        lattice_grid grid;
        alloc_grid(grid, nx, ny, 2);
        grid.filename = "results";
        set_output_folder(grid, "taylor_green");                                                                 // :37
        grid.collision = coll == "bgk" ? collide_bgk : coll == "trt" ? collide_trt : collide_rr;               // :39
        grid.streaming = scheme == "fvm" ? stream_fvm_bardow : scheme == "fdm" ? stream_fdm_bardow
                       : scheme == "sofonea" ? stream_fdm_sofonea : lbm_stream;                                  // :40
        if (scheme == "dugks") grid.collision = dugks_collide, grid.streaming = dugks_stream;
        grid.logger = my_logger;

        const wp umax = wp(0.01) / std::sqrt(wp(3));          // umax = Mach * cs
        const wp nu = (umax * wp(nx)) / wp(100);              // nu = (umax * L) / Re
        const wp tau = wp(3) * nu;
        dt = argv[1][0] == 'r' ? wp(std::atof(argv[1] + 1)) * tau : wp(std::atof(argv[1]));
        std::printf(" tau =  %.17g\n dt/tau =  %.17g\n cfl =  %.17g\n", (double)tau, (double)(dt / tau), (double)dt);
        set_properties(grid, nu, dt, wp(1) / wp(4));
        std::printf(" omega =  %.17g\n", (double)grid.omega);

        const wp kx = 2 * pi() / wp(nx), ky = 2 * pi() / wp(ny);
        tg = taylor_green_t(nx, ny, kx, ky, umax, nu);
        std::printf(" umax =  %.17g\n tc   =  %.17g\n", (double)umax, (double)tg.td);
        const wp tmax = std::log(wp(2)) * tg.decay_time();
        const long nsteps = (long)(wp(1.1) * tmax / dt);
        std::printf(" nsteps =  %ld\n", nsteps);

        wp t = 0;
        apply_initial_condition(tg, grid);
        output_gnuplot(grid, 0);  // :91-92
        output_vtk(grid, 0);
        grid.logger(grid, 0);

        const auto sbegin = std::chrono::steady_clock::now();
        long step;
        for (step = 1; step <= nsteps;) {
            // the reference steps one at a time; batching up to the next print / the final time is
            // equivalent because nothing reads the lattice in between (t is still accumulated per step)
            long batch = 0;
            bool last = false;
            while (step + batch <= nsteps && !last) {
                t = t + dt;
                ++batch;
                if ((step + batch - 1) % nprint == 0) break;
                if (t >= tmax) last = true;
            }
            if (scheme == "dugks")
                perform_dugks_step(grid, (int)batch);
            else
                perform_lbm_step(grid, (int)batch);
            step += batch;
            const long done = step - 1;
            if (done % nprint == 0 || t >= tmax) {
                std::printf(" step =  %ld\n", done);
                update_macros(grid);
                output_gnuplot(grid, (int)done);  // :107-108, :114-115
                output_vtk(grid, (int)done);
                grid.logger(grid, (int)done);
            }
            if (t >= tmax) break;
        }
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - sbegin).count();
        std::printf(" MLUPS  %.3f\n", (double)nx * ny * (double)(step - 1) * 1e-6 / secs);

        const wp nrm = calc_L2_norm(tg, grid, t);
        std::printf(" L2-norm =  %.8E\n Final time =  %.17g\n", (double)nrm, (double)t);
        dealloc_grid(grid);
    } catch (const plbm::error& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
