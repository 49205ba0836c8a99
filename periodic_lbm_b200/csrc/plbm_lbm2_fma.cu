// plbm_lbm2_fma.cu -- the two-step LBM kernels (k_lbm2_bulk, k_lbm2) with FMA contraction (opt-in, variant 11).
//
// The library is built with -fmad=false so that every result is bit-identical to the reference's non-FMA CPU
// arithmetic ("Parity and FMA").  With two steps per pass over HBM the BGK and TRT kernels sit on the HBM
// roof (83 GLUPS fp64 at 8192^2) but the recursive-regularized collision is bound by the fp64 pipe (68.5 GLUPS,
// ~200 flop per node and step): contracting a*b+c removes a good part of its instructions.  This translation unit
// is plbm_lbm2.cu compiled with -fmad=true (Makefile), exported as launch_lbm_pair_fma: results then differ from
// the non-FMA arithmetic in the last bits -- inside the tolerance BASELINE.json states (1e-12 relative fp64, 1e-5
// fp32), no longer bit-identical -- so it is never the default: plbm_set_variant(grid, 11) selects it for
// perform_lbm_step on one GPU (the closing single step of a call stays the non-FMA k_lbm).
// EXPERIMENTAL: written after round 1's GPU budget was spent; parity gate tests/test_gpu_zz_round1_late.py
// (PLBM_TEST_EXPERIMENTAL=1), A/B tools/pair_ab.py --variants 0,11.
#define PLBM_FMA_BUILD 1
#include "plbm_lbm2.cu"
