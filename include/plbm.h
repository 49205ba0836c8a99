/*
 * plbm.h -- C ABI of libplbm_b200.so, the B200 (sm_100a) implementation of the periodic
 * D2Q9 hot path of ivan-pi/periodic-lbm.
 *
 * This is the drop-in boundary: every entry point below is what a Fortran
 * `bind(c)` interface block (periodic_lbm_b200/fortran/plbm_c.f90), a ctypes stub
 * (periodic_lbm_b200/capi.py) or a C/C++ host would bind.  Signatures use plain pointers
 * and sizes only.  Each entry cites the reference interface it replaces (file:line
 * relative to the reference root).
 *
 * Conventions
 *  - All functions return 0 on success, non-zero on failure; plbm_last_error() returns a
 *    thread-local message.  Nothing aborts the process (the reference uses `error stop`).
 *  - Precision is a property of the handle: PLBM_F64 <-> default build of the reference,
 *    PLBM_F32 <-> the -DPRECISION_SINGLE build (src/precision.F90:11-15).  `void*` array
 *    arguments point to double or float accordingly; scalar parameters are passed as
 *    double and rounded to the working precision on entry (exact when the caller computed
 *    them in that precision).
 *  - Host array layouts are the reference's: PDFs f(ld,nx,0:8) column-major
 *    (src/fvm_bardow.F90:144-153), ld = ny rounded up to a multiple of 16; macroscopic
 *    fields (ny,nx) column-major, unpadded (src/fvm_bardow.F90:155).
 *  - Lattice indices are 1-based like grid%iold / grid%inew (src/fvm_bardow.F90:57,171-177).
 *  - The PDF lattices live on the device for the lifetime of the handle; host pointers are
 *    borrowed for the duration of a call.  Calls on one handle are stream-ordered; entry
 *    points that return data to the host synchronise that stream.
 *  - There is no CPU fallback: every compute entry fails with PLBM_ERR_CUDA when no
 *    sm_100-class device is usable.
 */
#ifndef PLBM_H
#define PLBM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct plbm_grid_s* plbm_handle;

enum plbm_precision { PLBM_F64 = 0, PLBM_F32 = 1 };

/* collision operators: collide_bgk (src/collision_bgk.F90:17), collide_trt
 * (src/collision_trt.F90:41), collide_rr (src/collision_regularized.F90:21), and the
 * -DSPLIT re-associated BGK (src/collision_bgk.F90:84-176). */
enum plbm_collision {
    PLBM_BGK = 0, PLBM_TRT = 1, PLBM_RR = 2,
    PLBM_BGK_SPLIT = 3,    /* collide_bgk built with -DSPLIT (bgk_kernel_cache)              */
    PLBM_TRT_SPLIT = 4,    /* collide_trt built with -DSPLIT (src/collision_trt.F90:162-290) */
    PLBM_BGK_IMPROVED = 5  /* collide_bgk_improved (src/collision_bgk_improved.f90:16-107)   */
};

/* streaming schemes: lbm_stream (src/periodic_lbm.f90:32), stream_fvm_bardow
 * (src/fvm_bardow.F90:393), stream_fdm_bardow (:511, default build: Lax-Wendroff with plain central
 * differences; the FDM_WLS / FDM_ISO cpp variants are not built), stream_fdm_sofonea (:688). */
enum plbm_streaming {
    PLBM_STREAM_LBM = 0, PLBM_STREAM_FVM_BARDOW = 1, PLBM_STREAM_FDM_BARDOW = 2, PLBM_STREAM_FDM_SOFONEA = 3
};

enum plbm_status {
    PLBM_OK = 0,
    PLBM_ERR_ARG = 1,     /* bad argument / unsupported combination */
    PLBM_ERR_CUDA = 2,    /* CUDA runtime failure or no device */
    PLBM_ERR_STATE = 3,   /* call sequence error (e.g. properties not set) */
    PLBM_ERR_COMM = 4     /* multi-GPU exchange failure */
};

/* scalar diagnostics returned by plbm_diagnostics (app/main_taylor_green.f90:106,155) */
enum plbm_diag {
    PLBM_DIAG_MAX_SPEED = 0, /* maxval(hypot(ux,uy)) */
    PLBM_DIAG_MIN_SPEED = 1, /* minval(hypot(ux,uy)) */
    PLBM_DIAG_SUM_RHO = 2,   /* sum(rho)             */
    PLBM_DIAG_KINETIC = 3,   /* 1/2 sum(rho (ux^2+uy^2)) */
    PLBM_DIAG_COUNT = 4
};

const char* plbm_last_error(void);
int plbm_version(void);
/* number of usable CUDA devices (0 when none) -- does not fail without a GPU */
int plbm_device_count(void);

/* ---- lifecycle: alloc_grid / dealloc_grid (src/fvm_bardow.F90:129-202, 204-240) ------ */
/* nf = number of PDF lattices (2, or 3 for perform_triple_step).  Sets inew=1, iold=2. */
int plbm_alloc_grid(plbm_handle* grid, int nx, int ny, int nf, int precision);
/* same, on an explicit CUDA device ordinal */
int plbm_alloc_grid_on(plbm_handle* grid, int nx, int ny, int nf, int precision, int device);
int plbm_dealloc_grid(plbm_handle grid);

int plbm_get_dims(plbm_handle grid, int* nx, int* ny, int* ld, int* nf, int* precision);
/* grid%iold / grid%inew / grid%imid (1-based; imid = -1 when nf == 2) */
int plbm_get_indices(plbm_handle grid, int* iold, int* inew, int* imid);
/* Set the lattice roles directly (restoring a checkpoint: the rotations of perform_triple_step cannot be reached by
 * plbm_swap alone).  (iold, inew) must be a permutation of {1,2} for nf = 2 (imid ignored), (iold, inew, imid) of {1,2,3}
 * for nf = 3; the reference keeps these as plain public components, src/fvm_bardow.F90:57. */
int plbm_set_indices(plbm_handle grid, int iold, int inew, int imid);

/* ---- set_properties (src/fvm_bardow.F90:242-269) ------------------------------------ */
/* tau = nu/cs^2, omega = dt/(tau+dt/2), trt_magic = magic or (tau/dt)^2 */
int plbm_set_properties(plbm_handle grid, double nu, double dt, double magic, int has_magic);
/* out[0..5] = nu, dt, tau, omega, trt_magic, csqr (working precision values widened) */
int plbm_get_properties(plbm_handle grid, double out[6]);
/* grid%omega is a public component the drivers may overwrite */
int plbm_set_omega(plbm_handle grid, double omega);

/* ---- initial condition: set_pdf_to_equilibrium (src/fvm_bardow.F90:272-305) --------- */
/* host rho, ux, uy (ny,nx) -> device macroscopic fields -> f(:,:,:,iold) = feq */
int plbm_set_pdf_to_equilibrium(plbm_handle grid, const void* rho, const void* ux, const void* uy);

/* ---- time stepping ------------------------------------------------------------------- */
/* perform_lbm_step (src/periodic_lbm.f90:15-29): lbm_stream + collision + swap, fused
 * into one pull-scheme kernel; nsteps >= 1 steps per call.  A call of nsteps >= 3 advances two
 * steps per pass over HBM (the intermediate lattice lives in shared memory) and closes with a
 * single step, so that after the call BOTH lattices and the indices are bit-identical to nsteps
 * reference steps (lattice `inew` = state nsteps-1, what the lagged update_macros reads):
 * batch the steps between two outputs into one call. */
int plbm_perform_lbm_step(plbm_handle grid, int collision, int nsteps);
/* Deferred stepping, for callers that keep the reference's call pattern of ONE step per call
 * (app/main_taylor_green.f90:98-119: `call perform_lbm_step(grid)` in a do loop, update_macros every nprint
 * steps).  With max_pending > 0, perform_lbm_step calls of fewer than max_pending steps are only counted; the
 * steps run as one batched call -- two steps per pass over HBM -- when max_pending steps have accumulated, when
 * the collision operator changes, or when ANY other entry point looks at or changes the grid (update_macros,
 * download_f, diagnostics, set_properties, a changed set_omega, synchronize, ...).  Results, lattice roles and
 * plbm_get_indices are those of eager stepping; only the moment the kernels are launched moves.  0 (default)
 * = eager.  Not applied under a slab decomposition. */
int plbm_set_step_deferral(plbm_handle grid, int max_pending);
/* perform_step (src/fvm_bardow.F90:307-320) with streaming = stream_fvm_bardow, stream_fdm_bardow
 * or stream_fdm_sofonea (PLBM_STREAM_LBM forwards to plbm_perform_lbm_step). */
int plbm_perform_step(plbm_handle grid, int streaming, int collision, int nsteps);
/* perform_triple_step (src/fvm_bardow.F90:322-340), needs nf = 3: like perform_step but the
 * streamed, pre-collision PDFs are kept in lattice `iold` before the indices rotate
 * (iold, inew, imid) <- (inew, imid, iold). */
int plbm_perform_triple_step(plbm_handle grid, int streaming, int collision, int nsteps);
/* perform_dugks_step (src/periodic_dugks.F90:25-38): dugks_collide + dugks_stream + swap.
 * dugks != 0 selects the -DDUGKS branch (half-step + face relaxation), 0 the default
 * build (degenerates to Bardow's scheme). */
int plbm_perform_dugks_step(plbm_handle grid, int dugks, int nsteps);

/* The reference's separately public kernels, for callers that assign grid%streaming /
 * grid%collision to something this library cannot fuse.  They act on iold/inew exactly
 * like the Fortran procedures and do NOT swap. */
int plbm_lbm_stream(plbm_handle grid);                       /* src/periodic_lbm.f90:32  */
int plbm_stream_fvm_bardow(plbm_handle grid);                /* src/fvm_bardow.F90:393   */
int plbm_stream_fdm_bardow(plbm_handle grid);                /* src/fvm_bardow.F90:511   */
int plbm_stream_fdm_sofonea(plbm_handle grid);               /* src/fvm_bardow.F90:688   */
int plbm_collide(plbm_handle grid, int collision);           /* collide_bgk/trt/rr       */
int plbm_dugks_collide(plbm_handle grid, int dugks);         /* src/periodic_dugks.F90:46  */
int plbm_dugks_stream(plbm_handle grid, int dugks);          /* src/periodic_dugks.F90:172 */
int plbm_swap(plbm_handle grid);                             /* the swap block, e.g. src/periodic_lbm.f90:22-27 */

/* ---- diagnostics ----------------------------------------------------------------------*/
/* update_macros (src/fvm_bardow.F90:343-390).  lagged != 0 reads f(:,:,:,inew) exactly
 * like the reference does after the swap (the state before the last step);
 * lagged == 0 reads f(:,:,:,iold), the current state.  Any of rho/ux/uy may be NULL. */
int plbm_update_macros(plbm_handle grid, void* rho, void* ux, void* uy, int lagged);
/* vorticity_2nd / vorticity_4th (src/vorticity.f90:13-87) of the device-resident ux,uy
 * last produced by plbm_update_macros / plbm_set_pdf_to_equilibrium; order = 2 or 4.
 * The 4th-order weights reproduce the reference. omega may be NULL. */
int plbm_vorticity(plbm_handle grid, int order, void* omega);
/* same operator on host arrays, like the Fortran signature vorticity_2nd(ux,uy,omega) */
int plbm_vorticity_host(plbm_handle grid, int order, const void* ux, const void* uy, void* omega);
/* warp-shuffle reductions over the device-resident macroscopic fields */
int plbm_diagnostics(plbm_handle grid, double out[PLBM_DIAG_COUNT]);
/* calc_L2_norm (app/main_taylor_green.f90:174-212) against host analytic fields uxa, uya:
 * out[0] = sum |u-ua|^2, out[1] = sum |ua|^2 ; L2 = sqrt(out[0]/out[1]) */
int plbm_l2_sums(plbm_handle grid, const void* uxa, const void* uya, double out[2]);

/* ---- raw PDF access (checkpoint / tests) ---------------------------------------------- */
/* which = 1-based lattice index; host buffer is f(ld,nx,0:8) */
/* 64-bit checksum of lattice `which` (rows 1..ny of every (x,q) line; padding rows excluded), computed on the device:
 * position-dependent and independent of the reduction order, so equal lattices <=> equal checksums.  Lets a caller compare
 * device-resident states (two runs, the slabs of a ring against a single-GPU run) without downloading the PDFs. */
int plbm_lattice_hash(plbm_handle grid, int which, unsigned long long* out);
int plbm_upload_f(plbm_handle grid, int which, const void* host_f);
int plbm_download_f(plbm_handle grid, int which, void* host_f);

/* ---- execution control ---------------------------------------------------------------- */
/* run on a caller-owned cudaStream_t (e.g. a torch stream); NULL = back to a library-owned
 * stream.  To use the legacy default stream pass cudaStreamLegacy ((void*)1), not 0. */
int plbm_set_stream(plbm_handle grid, void* cuda_stream);
int plbm_synchronize(plbm_handle grid);
/* kernels launched by this process through the library since load (for bench accounting) */
long long plbm_launch_count(void);
/* select a kernel variant (tuning / A-B measurements); 0 = default everywhere.
 *   perform_lbm_step : 0 = default: calls of >= 3 steps advance THREE steps per pass over HBM (k_lbm3_ws / k_lbmn_bulk, see
 *                      plbm_lbm_triple_kernel: bgk / trt / rr and the -DSPLIT / improved operators, fp64 and fp32, from 512^2
 *                      nodes; env PLBM_TRIPLES=0: never, 2: at every size).  With the third lattice buffer (plbm_lbm_closing_triple)
 *                      a call closes with a triple that stores states n-1 and n, preceded by one pair when two steps are left
 *                      over; without it the last one to four steps go as pairs (k_lbm2_bulk on large grids, k_lbm2 on smaller
 *                      ones; env PLBM_PAIR_BULK=0 / 2: k_lbm2 / k_lbm2_bulk everywhere) and one closing single step (k_lbm, direct
 *                      128-bit loads); grids that fit in the shared memory of one cluster: all steps in one launch
 *                      (cluster-resident kernel).  One step per launch:
 *                      1 warp-shuffle shifts, 2 scalar, 3 TMA-staged tile, 4 streaming hints;
 *                      5 = like 0 but never the cluster kernel (tests of the two-step kernel on small grids);
 *                      6 / 7 = like 5 with the two-step kernel's raw columns fetched by per-thread loads (k_lbm2) /
 *                      by bulk async copies (k_lbm2_bulk); 8 = 7 issued as the three line ranges of the slab schedule;
 *                      9 / 10 = the depth-generic kernel k_lbmn_bulk forced on every grid it applies to: pairs / triples
 *                      (10 is what variant 0 does from 512^2 nodes);
 *                      11 = EXPERIMENTAL the two-step kernels compiled with FMA contraction (one GPU; within 1e-12 /
 *                      1e-5 relative of the non-FMA result, NOT bit-identical);
 *                      12 = like 5, fp32 collisions in scalar instead of packed (two nodes per FFMA2) form: same bits, A/B only
 *   perform_step (fvm/fdm) : 0 TMA + mbarrier pipelined tile kernel, 2 plain-load tile kernel
 *   perform_dugks_step     : 0 TMA-pipelined fused kernel, 1 the reference's two passes, 2 plain-load fused
 *   both                   : 3 = opt-in: the TMA-pipelined kernel compiled with FMA contraction (fewer fp64
 *                            instructions; within 1e-12 / 1e-5 relative of the non-FMA result, NOT bit-identical; one GPU);
 *                            4 = opt-in: marching kernel with shared cell faces (k_fv_march / k_fv_march_s, csrc/plbm_fvm_march.cu:
 *                            every face reconstructed and relaxed once, FMA contraction; same tolerance gate, NOT bit-identical;
 *                            DUGKS fp64 2048^2 24.4 GLUPS against 14.4 of the default) */
int plbm_set_variant(plbm_handle grid, int variant);
/* which kernel perform_lbm_step(nsteps >= 3) advances this grid with: 0 = one step per launch (k_lbm),
 * 1 = two steps per launch, raw columns by per-thread loads (k_lbm2), 2 = two steps per launch, raw columns by
 * bulk async copies (k_lbm2_bulk).  Grids that fit in the shared memory of one cluster are advanced by the cluster-
 * resident kernel instead when variant == 0 and nsteps >= 4; this query does not look at that.  For bench accounting;
 * -1 on a null handle. */
int plbm_lbm_pair_kernel(plbm_handle grid);
/* how many time steps one pass over HBM advances in a perform_lbm_step call of many steps with this collision: 3 = k_lbmn_bulk
 * (three fused steps: fp64 BGK / TRT on large grids, see csrc/plbm_api.cu lbm_triples_wanted), 2 = k_lbm2 / k_lbm2_bulk, 1 = k_lbm.
 * For bench accounting; -1 on a null handle. */
int plbm_lbm_steps_per_pass(plbm_handle grid, int collision);
/* 1 if a perform_lbm_step call of three or more steps with this collision closes with a fused triple that stores the states after
 * its second AND third step (lattice `inew` = state n-1, `iold` = state n, what the reference's last single step leaves), 0 if it
 * closes with a single-step launch.  The closing triple needs a third, hidden lattice buffer: allocated at the first query / call when
 * three steps per pass apply and the GPU has room (environment PLBM_SPARE_LATTICE: 0 never, 1 default -- only if a quarter of the
 * buffer + 1 GB stays free --, 2 whenever the allocation succeeds); under a slab decomposition every rank must have it.  Results are
 * bit-identical either way.  For bench accounting; -1 on a null handle. */
int plbm_lbm_closing_triple(plbm_handle grid, int collision);
/* which kernel runs the three-step launches of this collision that read no halo lines (every launch on one GPU, the interior
 * launches of a slab): 0 = k_lbmn_bulk (csrc/plbm_lbmn.cu: levels one after the other, a barrier after each), 1 = k_lbm3_ws
 * (csrc/plbm_lbm3w.cu: a producer warp, levels skewed by two columns, one barrier per column; the default except for the
 * two-relaxation-time collisions).  Environment PLBM_TRIPLE_WS = 0 / 1 forces one of them; bit-identical.  -1 on a null handle. */
int plbm_lbm_triple_kernel(plbm_handle grid, int collision);
/* derivative stencil of stream_fdm_bardow: the reference selects it at compile time with -DFDM_WLS,
 * -DFDM_WLS_GAUSS_V1, -DFDM_WLS_GAUSS_V2 or -DFDM_ISO (src/fvm_bardow.F90:591-660); default = none of them. */
enum plbm_fdm_stencil { PLBM_FDM_DEFAULT = 0, PLBM_FDM_WLS = 1, PLBM_FDM_WLS_GAUSS_V1 = 2, PLBM_FDM_WLS_GAUSS_V2 = 3, PLBM_FDM_ISO = 4 };
int plbm_set_fdm_stencil(plbm_handle grid, int stencil);

/* ---- flow cases (host side, O(N) input generators) ------------------------------------- */
/* taylor_green_t%decay_time / %eval (src/benchmarks/taylor_green.f90:31-84): fills host
 * arrays p, ux, uy (ny,nx) at time t, cell centres (x-1/2, y-1/2), working precision. */
double plbm_case_tg_decay_time(int precision, double kx, double ky, double nu);
int plbm_case_taylor_green(int precision, int nx, int ny, double kx, double ky, double umax, double td,
                           double t, void* p, void* ux, void* uy);
/* same, for the nx lines of a slab whose first line is global line x0 (multi-GPU ICs) */
int plbm_case_taylor_green_slab(int precision, int nx, int x0, int ny, double kx, double ky, double umax,
                                double td, double t, void* p, void* ux, void* uy);
/* vortex_case_t%eval (src/benchmarks/barotropic_vortex_case.F90:34-83) */
int plbm_case_vortex(int precision, int nx, int ny, double U0, double xc, double yc, double Rc, double eps,
                     double rho0, double csqr, void* rho, void* ux, void* uy);

/* ---- multi-GPU slabs (new functionality) ----------------------------------- */
/* The global grid is nx_global x ny; this handle owns the slab of nx lines starting at
 * x_offset along the slow index.  Ranks form a periodic ring.  The caller obtains a
 * 128-byte NCCL unique id on rank 0 (plbm_comm_unique_id), distributes it by any means
 * (torch.distributed / MPI_Bcast) and every rank calls plbm_comm_init. */
int plbm_comm_unique_id(void* id128);
int plbm_comm_init(plbm_handle grid, const void* id128, int rank, int nranks, int nx_global, int x_offset);
/* halo transport in use: 1 = p2p (CUDA IPC stores over NVLink + stream wait-value flags, default),
 * 0 = nccl (grouped ncclSend/ncclRecv; PLBM_HALO=nccl or no IPC), -1 = no ring */
int plbm_comm_transport(plbm_handle grid);
int plbm_comm_finalize(plbm_handle grid);

#ifdef __cplusplus
}
#endif
#endif /* PLBM_H */
