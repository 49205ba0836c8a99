/*
 * plbm_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU oracle for the periodic D2Q9 hot path of ivan-pi/periodic-lbm: a plain-C,
 * line-faithful restatement of the reference's Fortran kernels (file:line cited
 * per function in plbm_oracle_impl.h).  It exists to CHECK the CUDA library and
 * to serve as the CPU baseline in bench.py; nothing in the product path may
 * call it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs load this library.
 *
 * Parity status: the reference is Fortran and no Fortran compiler exists in the
 * build image, so oracle/_ref cannot be produced.  The oracle is
 * PINNED against the only golden data the reference ships for this path,
 * graphs/fvm_bardow_64.txt and graphs/fvm_dugks_64.txt (tests/test_oracle_golden.py),
 * which cover equilibrium, set_pdf_to_equilibrium, stream_fvm_bardow,
 * collide_bgk, kernel_bgk, dugks_collide, dugks_stream (+update_ew/ns),
 * update_macros (lagged), the Taylor-Green case and the L2 norm.
 * lbm_stream, collide_trt, collide_rr, vorticity_*, the finite-difference schemes, the
 * -DSPLIT collisions and every fp32 result are pinned by no reference ARTEFACT; they are
 * pinned on the reference's SOURCE: oracle/f90_exec.py executes the Fortran statements of
 * /root/reference/src one by one, and this oracle reproduces the outputs bit for bit in
 * fp64 and fp32 (tests/golden/refsrc_*.npz, tests/test_oracle_refsrc.py).  The reference
 * binary itself has never run here; the Taylor-Green / vortex field evaluations
 * (transcendental functions) rest on the restatement and the golden files only.
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off, OpenMP optional).
 */
#include <math.h>
#include <stddef.h>

#define REAL double
#define SFX(x) x##_f64
#define MSQRT sqrt
#define MSIN sin
#define MCOS cos
#define MEXP exp
#define MFABS fabs
#define MHYPOT hypot
#include "plbm_oracle_impl.h"
#undef REAL
#undef SFX
#undef MSQRT
#undef MSIN
#undef MCOS
#undef MEXP
#undef MFABS
#undef MHYPOT

#define REAL float
#define SFX(x) x##_f32
#define MSQRT sqrtf
#define MSIN sinf
#define MCOS cosf
#define MEXP expf
#define MFABS fabsf
#define MHYPOT hypotf
#include "plbm_oracle_impl.h"

#ifdef _OPENMP
#include <omp.h>
int orc_num_threads(void) { return omp_get_max_threads(); }
void orc_set_num_threads(int n) { omp_set_num_threads(n); }
#else
int orc_num_threads(void) { return 1; }
void orc_set_num_threads(int n) { (void)n; }
#endif
