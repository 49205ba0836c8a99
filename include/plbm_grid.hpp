// plbm_grid.hpp -- C++ host-side mirror of the reference's Fortran module interfaces over the C ABI
// (include/plbm.h).  The reference's toolchain (Fortran) is absent from the build image, so this is the
// compiled-language host side: same type, procedure and component names, same argument meaning, same call
// sequences, so a driver written against the Fortran modules reads the same here (app/*.cpp).
//
//   module fvm_bardow            lattice_grid, alloc_grid, dealloc_grid, set_properties, perform_step,
//                                perform_triple_step, update_macros, set_pdf_to_equilibrium, equilibrium,
//                                stream_fvm_bardow, stream_fdm_bardow, stream_fdm_sofonea, cx, cy, csqr
//                                                                      (src/fvm_bardow.F90:13-33)
//   module periodic_lbm          perform_lbm_step, lbm_stream          (src/periodic_lbm.f90:9-11)
//   module collision_bgk/trt/regularized/bgk_improved   collide_bgk, collide_trt, magic_number, lambda_d,
//                                collide_rr, collide_bgk_improved
//   module periodic_dugks        perform_dugks_step, dugks_collide, dugks_stream
//   module vorticity             vorticity_2nd, vorticity_4th          (src/vorticity.f90)
//   module taylor_green          taylor_green_t, pi                    (src/benchmarks/taylor_green.f90)
//   module barotropic_vortex_case  vortex_case_t                       (src/benchmarks/barotropic_vortex_case.F90)
//
// Differences, all forced by the device-resident lattices: `grid.f` does not exist on the host (no driver
// reads it); `grid.dev` is the opaque device handle; failures throw plbm::error (the reference `error stop`s).
// Arrays keep the Fortran layout: a field (ny,nx) is stored y-fastest, element (y,x) at [y + ny*x], 0-based.
// Define PRECISION_SINGLE for the reference's -DPRECISION_SINGLE build (src/precision.F90:11-15).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "plbm.h"

namespace plbm {

#ifdef PRECISION_SINGLE
using wp = float;
constexpr int plbm_precision_id = PLBM_F32;
#else
using wp = double;
constexpr int plbm_precision_id = PLBM_F64;
#endif

struct error : std::runtime_error {
    using std::runtime_error::runtime_error;
};
inline void check(int stat, const char* what)
{
    if (stat != 0) throw error(std::string(what) + " failed: " + plbm_last_error());
}

// ---- module fvm_bardow -----------------------------------------------------------------------------
constexpr wp cx[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
constexpr wp cy[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
constexpr wp csqr = wp(1) / wp(3);

struct lattice_grid;
using collision_interface = void (*)(lattice_grid&);
using streaming_interface = void (*)(lattice_grid&);
using gridlog_interface = void (*)(const lattice_grid&, int step);

struct lattice_grid {
    int nx = 0, ny = 0;
    std::vector<wp> mf;                     // macroscopic fields (ny,nx,3), contiguous
    wp *rho = nullptr, *ux = nullptr, *uy = nullptr;  // views of mf(:,:,1..3)
    wp nu = 0, dt = 0, tau = 0;
    wp omega = 0, trt_magic = 0;
    wp csqr = 0;
    int iold = 2, inew = 1, imid = -1;
    collision_interface collision = nullptr;
    streaming_interface streaming = nullptr;
    std::string filename, foldername, logfile;
    gridlog_interface logger = nullptr;
    std::FILE* logunit = nullptr;
    plbm_handle dev = nullptr;              // device-resident lattices
    bool dugks = true;                      // periodic_dugks as built with -DDUGKS
    size_t size() const { return (size_t)nx * ny; }
};

inline void sync_indices(lattice_grid& grid) { check(plbm_get_indices(grid.dev, &grid.iold, &grid.inew, &grid.imid), "get_indices"); }

inline void alloc_grid(lattice_grid& grid, int nx, int ny, int nf = 2, bool log = true)
{
    grid.nx = nx;
    grid.ny = ny;
    grid.mf.assign((size_t)3 * nx * ny, wp(0));
    grid.rho = grid.mf.data();
    grid.ux = grid.rho + grid.size();
    grid.uy = grid.ux + grid.size();
    check(plbm_alloc_grid(&grid.dev, nx, ny, nf, plbm_precision_id), "alloc_grid");
    sync_indices(grid);
    if (log) {
        if (grid.logfile.empty()) grid.logfile = "lattice_grid_log.txt";
        grid.logunit = std::fopen(grid.logfile.c_str(), "w");
    }
}

inline void dealloc_grid(lattice_grid& grid)
{
    if (grid.logunit) std::fclose(grid.logunit);
    grid.logunit = nullptr;
    grid.rho = grid.ux = grid.uy = nullptr;
    grid.mf.clear();
    if (grid.dev) check(plbm_dealloc_grid(grid.dev), "dealloc_grid");
    grid.dev = nullptr;
}

inline void set_properties(lattice_grid& grid, wp nu, wp dt, const wp* magic = nullptr)
{
    check(plbm_set_properties(grid.dev, nu, dt, magic ? *magic : 0.0, magic != nullptr), "set_properties");
    double p[6];
    check(plbm_get_properties(grid.dev, p), "get_properties");  // derived in working precision by the library
    grid.nu = (wp)p[0];
    grid.dt = (wp)p[1];
    grid.tau = (wp)p[2];
    grid.omega = (wp)p[3];
    grid.trt_magic = (wp)p[4];
    grid.csqr = (wp)p[5];
    std::printf(" trt magic =  %.17g\n", (double)grid.trt_magic);
}
inline void set_properties(lattice_grid& grid, wp nu, wp dt, wp magic) { set_properties(grid, nu, dt, &magic); }

// grid.omega is a public component a driver may overwrite: pushed before every launch
inline void push_omega(lattice_grid& grid) { check(plbm_set_omega(grid.dev, grid.omega), "set_omega"); }

// host-side convenience (the device kernels carry the bit-exact evaluation order)
inline void equilibrium(wp rho, wp ux, wp uy, wp feq[9])
{
    const wp w[9] = {wp(4) / 9, wp(1) / 9, wp(1) / 9, wp(1) / 9, wp(1) / 9, wp(1) / 36, wp(1) / 36, wp(1) / 36, wp(1) / 36};
    const wp indp = wp(1) - wp(1.5) * (ux * ux + uy * uy);
    for (int q = 0; q < 9; ++q) {
        const wp cu = cx[q] * ux + cy[q] * uy;
        feq[q] = w[q] * rho * (indp + wp(3) * cu + wp(4.5) * cu * cu);
    }
}

inline void set_pdf_to_equilibrium(lattice_grid& grid)
{
    check(plbm_set_pdf_to_equilibrium(grid.dev, grid.rho, grid.ux, grid.uy), "set_pdf_to_equilibrium");
}

// update_macros: rho, ux, uy of lattice `inew` -- the reference's one-step lag
inline void update_macros(lattice_grid& grid) { check(plbm_update_macros(grid.dev, grid.rho, grid.ux, grid.uy, 1), "update_macros"); }

inline void stream_fvm_bardow(lattice_grid& grid) { check(plbm_stream_fvm_bardow(grid.dev), "stream_fvm_bardow"); }
inline void stream_fdm_bardow(lattice_grid& grid) { check(plbm_stream_fdm_bardow(grid.dev), "stream_fdm_bardow"); }
inline void stream_fdm_sofonea(lattice_grid& grid) { check(plbm_stream_fdm_sofonea(grid.dev), "stream_fdm_sofonea"); }

// ---- module periodic_lbm ---------------------------------------------------------------------------
inline void lbm_stream(lattice_grid& grid) { check(plbm_lbm_stream(grid.dev), "lbm_stream"); }

// ---- collision modules -----------------------------------------------------------------------------
inline void collide_bgk(lattice_grid& grid)
{
    push_omega(grid);
#ifdef SPLIT
    check(plbm_collide(grid.dev, PLBM_BGK_SPLIT), "collide_bgk");
#else
    check(plbm_collide(grid.dev, PLBM_BGK), "collide_bgk");
#endif
}
inline wp magic_number(wp le, wp ld) { return (wp(2) - le) * (wp(2) - ld) / (wp(4) * le * ld); }
inline wp lambda_d(wp omega, wp x) { return (wp(4) - wp(2) * omega) / (wp(4) * x * omega + wp(2) - omega); }
inline void collide_trt(lattice_grid& grid)
{
    push_omega(grid);
#ifdef SPLIT
    check(plbm_collide(grid.dev, PLBM_TRT_SPLIT), "collide_trt");
#else
    check(plbm_collide(grid.dev, PLBM_TRT), "collide_trt");
#endif
}
inline void collide_rr(lattice_grid& grid)
{
    push_omega(grid);
    check(plbm_collide(grid.dev, PLBM_RR), "collide_rr");
}
inline void collide_bgk_improved(lattice_grid& grid)
{
    push_omega(grid);
    check(plbm_collide(grid.dev, PLBM_BGK_IMPROVED), "collide_bgk_improved");
}

// ---- host output of the macroscopic fields (src/fvm_bardow.F90:895-1025, src/output/*) --------------------
// grid.rho/ux/uy are what the last update_macros left on the host; nothing below touches the device.
inline void set_output_folder(lattice_grid& grid, const std::string& foldername, bool verbose = false)
{
    const std::string cmd = std::string("mkdir -p ") + (verbose ? "-v " : "") + foldername;
    if (std::system(cmd.c_str()) != 0) throw error("[set_output_folder] error making directory " + foldername);
    grid.foldername = foldername;
}

namespace detail {
// the reference's edit descriptors: ES24.16E3 (double) / ES15.8E2 (single); buf holds 48 chars
inline void fmt_real(char* buf, wp v)
{
    if (sizeof(wp) == 4) {
        std::snprintf(buf, 48, "%15.8E", (double)v);
        return;
    }
    char t[40];
    std::snprintf(t, sizeof(t), "%.16E", (double)v);
    char* e = std::strchr(t, 'E');
    const int ex = std::atoi(e + 2);
    std::snprintf(e + 2, 8, "%03d", ex);  // C prints two exponent digits, the descriptor has three
    std::snprintf(buf, 48, "%24s", t);
}
inline std::string output_name(const lattice_grid& grid, const int* step, const char* ext)
{
    char istr[16] = "";
    if (step) std::snprintf(istr, sizeof(istr), "%09d", *step);
    return grid.foldername + "/" + grid.filename + istr + ext;
}
}  // namespace detail

// output_gnuplot -> output_gnuplot_grid (src/output/gnuplot.F90:21-39): "x y rho ux uy" per node, a blank line
// after every line x.  The reference never assigns its `ry` (it assigns `rx` twice), so its second column is
// undefined; the cell centres (x - 1/2, y - 1/2) it evidently intends are written here.
inline void output_gnuplot(const lattice_grid& grid, const int* step = nullptr)
{
    const std::string name = detail::output_name(grid, step, ".txt");
    std::FILE* fh = std::fopen(name.c_str(), "w");
    if (!fh) throw error("output_gnuplot: cannot open " + name);
    char b[5][48];
    for (int x = 0; x < grid.nx; ++x) {
        detail::fmt_real(b[0], wp(x) + wp(0.5));
        for (int y = 0; y < grid.ny; ++y) {
            const size_t m = (size_t)x * grid.ny + y;
            detail::fmt_real(b[1], wp(y) + wp(0.5));
            detail::fmt_real(b[2], grid.rho[m]);
            detail::fmt_real(b[3], grid.ux[m]);
            detail::fmt_real(b[4], grid.uy[m]);
            std::fprintf(fh, "%s %s %s %s %s\n", b[0], b[1], b[2], b[3], b[4]);
        }
        std::fputc('\n', fh);
    }
    std::fclose(fh);
}
inline void output_gnuplot(const lattice_grid& grid, int step) { output_gnuplot(grid, &step); }

// output_vtk -> output_vtk_structuredPoints (src/fvm_bardow.F90:960-997, src/output/vtk.F90:150-196)
inline void output_vtk(const lattice_grid& grid, const int* step = nullptr, bool binary = false)
{
    if (binary) {
        std::printf(" binary output not implemented\n");
        return;
    }
    const std::string name = detail::output_name(grid, step, ".vtk");
    std::FILE* fh = std::fopen(name.c_str(), "w");
    if (!fh) throw error("output_vtk: cannot open " + name);
    char z[48], o[48], a[48], b[48];
    detail::fmt_real(z, wp(0));
    detail::fmt_real(o, wp(1));
    std::fprintf(fh, "# vtk DataFile Version 3.0\nfluid\nASCII\nDATASET STRUCTURED_POINTS\n");
    std::fprintf(fh, "DIMENSIONS %d %d 2 \n", grid.nx + 1, grid.ny + 1);
    std::fprintf(fh, "ORIGIN  %s%s%s\nSPACING %s%s%s\n\n", z, z, z, o, o, o);
    std::fprintf(fh, "CELL_DATA %lld\nSCALARS Density float 1\nLOOKUP_TABLE default\n", (long long)grid.nx * grid.ny);
    for (int y = 0; y < grid.ny; ++y)
        for (int x = 0; x < grid.nx; ++x) {
            detail::fmt_real(a, grid.rho[(size_t)x * grid.ny + y]);
            std::fprintf(fh, "%s\n", a);
        }
    std::fprintf(fh, "\nVECTORS Velocity float\n");
    for (int y = 0; y < grid.ny; ++y)
        for (int x = 0; x < grid.nx; ++x) {
            detail::fmt_real(a, grid.ux[(size_t)x * grid.ny + y]);
            detail::fmt_real(b, grid.uy[(size_t)x * grid.ny + y]);
            std::fprintf(fh, "%s%s%s\n", a, b, z);
        }
    std::fclose(fh);
}
inline void output_vtk(const lattice_grid& grid, int step, bool binary = false) { output_vtk(grid, &step, binary); }

// output_npy -> output_fluid_npy (src/fvm_bardow.F90:929-958, src/output/npy.f90:14-29): mf(ny,nx,3) in Fortran
// order as a NumPy v1.0 file, what stdlib's save_npy writes
inline void output_npy(const lattice_grid& grid, const int* step = nullptr)
{
    const std::string name = detail::output_name(grid, step, ".npy");
    std::FILE* fh = std::fopen(name.c_str(), "wb");
    if (!fh) throw error("output_npy: cannot open " + name);
    char dict[160];
    std::snprintf(dict, sizeof(dict), "{'descr': '<f%d', 'fortran_order': True, 'shape': (%d, %d, 3), }", (int)sizeof(wp), grid.ny, grid.nx);
    std::string header(dict);
    while ((10 + header.size() + 1) % 64 != 0) header.push_back(' ');
    header.push_back('\n');
    const unsigned char magic[8] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
    const unsigned short hlen = (unsigned short)header.size();
    std::fwrite(magic, 1, 8, fh);
    std::fwrite(&hlen, 2, 1, fh);  // little endian hosts only (x86-64, aarch64)
    std::fwrite(header.data(), 1, header.size(), fh);
    std::fwrite(grid.mf.data(), sizeof(wp), grid.mf.size(), fh);
    std::fclose(fh);
}
inline void output_npy(const lattice_grid& grid, int step) { output_npy(grid, &step); }

namespace detail {
inline int collision_id(collision_interface c)
{
#ifdef SPLIT
    if (c == collide_bgk) return PLBM_BGK_SPLIT;
    if (c == collide_trt) return PLBM_TRT_SPLIT;
#else
    if (c == collide_bgk) return PLBM_BGK;
    if (c == collide_trt) return PLBM_TRT;
#endif
    if (c == collide_rr) return PLBM_RR;
    if (c == collide_bgk_improved) return PLBM_BGK_IMPROVED;
    return -1;
}
inline int streaming_id(streaming_interface s)
{
    if (s == lbm_stream) return PLBM_STREAM_LBM;
    if (s == stream_fvm_bardow) return PLBM_STREAM_FVM_BARDOW;
    if (s == stream_fdm_bardow) return PLBM_STREAM_FDM_BARDOW;
    if (s == stream_fdm_sofonea) return PLBM_STREAM_FDM_SOFONEA;
    return -1;
}
}  // namespace detail

// perform_lbm_step / perform_step: streaming(); collision(); swap.  For the procedure pairs the library
// fuses, ONE kernel does the whole step (n steps per call avoid a host round trip per step); any other
// (user-supplied) pair is called one after the other exactly like the reference.
inline void perform_lbm_step(lattice_grid& grid, int n = 1)
{
    const int cid = detail::collision_id(grid.collision), sid = detail::streaming_id(grid.streaming);
    push_omega(grid);
    if (cid >= 0 && sid >= 0) {
        check(plbm_perform_step(grid.dev, sid, cid, n), "perform_lbm_step");
    } else {
        for (int i = 0; i < n; ++i) {
            grid.streaming(grid);
            grid.collision(grid);
            check(plbm_swap(grid.dev), "swap");
        }
    }
    sync_indices(grid);
}
inline void perform_step(lattice_grid& grid, int n = 1) { perform_lbm_step(grid, n); }

inline void perform_triple_step(lattice_grid& grid, int n = 1)
{
    const int cid = detail::collision_id(grid.collision), sid = detail::streaming_id(grid.streaming);
    if (cid < 0 || sid < 0) throw error("perform_triple_step: unknown streaming/collision procedure");
    push_omega(grid);
    check(plbm_perform_triple_step(grid.dev, sid, cid, n), "perform_triple_step");
    sync_indices(grid);
}

// ---- module periodic_dugks -------------------------------------------------------------------------
inline void dugks_collide(lattice_grid& grid) { check(plbm_dugks_collide(grid.dev, grid.dugks), "dugks_collide"); }
inline void dugks_stream(lattice_grid& grid) { check(plbm_dugks_stream(grid.dev, grid.dugks), "dugks_stream"); }
inline void perform_dugks_step(lattice_grid& grid, int n = 1)
{
    push_omega(grid);
    const bool fused = (!grid.collision || grid.collision == dugks_collide) && (!grid.streaming || grid.streaming == dugks_stream);
    if (fused) {
        check(plbm_perform_dugks_step(grid.dev, grid.dugks, n), "perform_dugks_step");
    } else {
        for (int i = 0; i < n; ++i) {
            grid.collision(grid);
            grid.streaming(grid);
            check(plbm_swap(grid.dev), "swap");
        }
    }
    sync_indices(grid);
}

// ---- module vorticity ------------------------------------------------------------------------------
// ux, uy, omega are (ny,nx) host arrays like the Fortran assumed-shape arguments
inline void vorticity_nth(int order, int nx, int ny, const wp* ux, const wp* uy, wp* omega)
{
    plbm_handle tmp = nullptr;
    check(plbm_alloc_grid(&tmp, nx, ny, 2, plbm_precision_id), "alloc_grid");
    const int rc = plbm_vorticity_host(tmp, order, ux, uy, omega);
    plbm_dealloc_grid(tmp);
    check(rc, "vorticity");
}
inline void vorticity_2nd(int nx, int ny, const wp* ux, const wp* uy, wp* omega) { vorticity_nth(2, nx, ny, ux, uy, omega); }
inline void vorticity_4th(int nx, int ny, const wp* ux, const wp* uy, wp* omega) { vorticity_nth(4, nx, ny, ux, uy, omega); }

// ---- module taylor_green ---------------------------------------------------------------------------
inline wp pi() { return wp(4) * std::atan(wp(1)); }

struct taylor_green_t {
    int nx = 0, ny = 0;
    wp kx = 0, ky = 0, umax = 0, nu = 0, td = 0;
    taylor_green_t() = default;
    taylor_green_t(int nx_, int ny_, wp kx_, wp ky_, wp umax_, wp nu_) : nx(nx_), ny(ny_), kx(kx_), ky(ky_), umax(umax_), nu(nu_)
    {
        td = (wp)plbm_case_tg_decay_time(plbm_precision_id, kx, ky, nu);
    }
    wp decay_time() const { return td; }
    void eval(wp t, wp* p, wp* ux, wp* uy) const
    {
        check(plbm_case_taylor_green(plbm_precision_id, nx, ny, kx, ky, umax, td, t, p, ux, uy), "taylor_green eval");
    }
};

// ---- module barotropic_vortex_case -----------------------------------------------------------------
struct vortex_case_t {
    wp U0 = 0, xc = 0, yc = 0, Rc = 0, eps = 0;
    wp rho0 = wp(1);
    wp csqr = wp(1) / wp(3);
    void eval(int nx, int ny, wp* rho, wp* ux, wp* uy) const
    {
        check(plbm_case_vortex(plbm_precision_id, nx, ny, U0, xc, yc, Rc, eps, rho0, csqr, rho, ux, uy), "vortex eval");
    }
};

}  // namespace plbm
