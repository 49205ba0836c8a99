/*
 * plbm_sim_plugin.h -- the reference's C plugin seam (sim/lbm.h:13-19) as exported by
 * libplbm_b200.so.  The reference loader (sim/cases.py:13-66) derives the symbol prefix from
 * the library file name, lib<name>.so -> c_<name>_{init,step,vars,free,norm}; with <name> =
 * "plbm" these are the concrete prototypes of the reference's `slbm` plugin
 * (sim/sim_slbm.F90:135-205).  Arrays are Fortran-contiguous rho(nx,ny), u(nx,ny,2) (x
 * fastest); `rho` carries PRESSURE on input (rho = 1 + p/cs^2, sim/sim.F90:181-199); dt must
 * be 1 (sim/sim_slbm.F90:48-51).  c_plbm_init returns NULL on failure instead of `error stop`
 * (message via plbm_last_error()).
 */
#ifndef PLBM_SIM_PLUGIN_H
#define PLBM_SIM_PLUGIN_H

#ifdef __cplusplus
extern "C" {
#endif

/* void *siminit(int nx, int ny, double dt, double *rho, double *u, double *sigma, void *params) */
void* c_plbm_init(int nx, int ny, double dt, const double* rho, const double* u, const double* sigma, void* params);
/* void simstep(void *sim, double omega): collide -> push-stream -> periodic fold, once */
void c_plbm_step(void* sim, double omega);
/* extension: n steps per call (the reference pays one ctypes round trip per step) */
void c_plbm_step_n(void* sim, double omega, int n);
/* void simvars(void *sim, double *rho, double *u) */
void c_plbm_vars(void* sim, double* rho, double* u);
/* void simfree(void *sim) */
void c_plbm_free(void* sim);
/* c_slbm_norm (sim/sim_slbm.F90:198-205): norm2(u - ua) / norm2(ua) over nx*ny host values */
double c_plbm_norm(int nx, int ny, const double* u, const double* ua);

/* The same entry points under the reference plugin's own name (sim/sim_slbm.F90:135-205), so that a
 * symlink libslbm.so -> libplbm_b200.so is a drop-in for the reference's libslbm.so. */
void* c_slbm_init(int nx, int ny, double dt, const double* rho, const double* u, const double* sigma, void* params);
void c_slbm_step(void* sim, double omega);
void c_slbm_vars(void* sim, double* rho, double* u);
void c_slbm_free(void* sim);
double c_slbm_norm(int nx, int ny, const double* u, const double* ua);

/* The reference's second-order Lax-Wendroff plugin `lw` (sim/sim_lw.F90: lw_stream :24-85, lw_collision
 * :87-166, lw_bc :169-197, exports :333-425), one fused kernel per step; any dt.  liblw.so drop-in. */
void* c_lw_init(int nx, int ny, double dt, const double* rho, const double* u, const double* sigma, void* params);
void c_lw_step(void* sim, double omega);
void c_lw_step_n(void* sim, double omega, int n);
void c_lw_vars(void* sim, double* rho, double* u);
void c_lw_free(void* sim);
double c_lw_norm(int nx, int ny, const double* u, const double* ua);

/* The fourth- and sixth-order Lax-Wendroff plugins `lw4` (sim/sim_lw4.F90: lw4_stream :26-115, lw4_collision
 * :118-195, lw4_bc :198-245, exports :370-477) and `lw6` (sim/sim_lw6.F90: lw6_stream :26-127, exports at the
 * end of the file): same collision as `lw`, wider streaming stencils (halo 2 / 3 in the reference, periodic
 * index arithmetic here).  liblw4.so / liblw6.so drop-ins. */
void* c_lw4_init(int nx, int ny, double dt, const double* rho, const double* u, const double* sigma, void* params);
void c_lw4_step(void* sim, double omega);
void c_lw4_step_n(void* sim, double omega, int n);
void c_lw4_vars(void* sim, double* rho, double* u);
void c_lw4_free(void* sim);
double c_lw4_norm(int nx, int ny, const double* u, const double* ua);
void* c_lw6_init(int nx, int ny, double dt, const double* rho, const double* u, const double* sigma, void* params);
void c_lw6_step(void* sim, double omega);
void c_lw6_step_n(void* sim, double omega, int n);
void c_lw6_vars(void* sim, double* rho, double* u);
void c_lw6_free(void* sim);
double c_lw6_norm(int nx, int ny, const double* u, const double* ua);

/* The Heun finite-volume plugin `fvm` (sim/sim_fvm.F90: fvm_predict_hc :67-100, fvm_correct_hc :103-136,
 * fvm_collision :139-188, fvm_bc :191-218, sim_fvm%step :282-322, exports :340-432): three kernels per step. */
void* c_fvm_init(int nx, int ny, double dt, const double* rho, const double* u, const double* sigma, void* params);
void c_fvm_step(void* sim, double omega);
void c_fvm_step_n(void* sim, double omega, int n);
void c_fvm_vars(void* sim, double* rho, double* u);
void c_fvm_free(void* sim);
double c_fvm_norm(int nx, int ny, const double* u, const double* ua);

#ifdef __cplusplus
}
#endif
#endif /* PLBM_SIM_PLUGIN_H */
