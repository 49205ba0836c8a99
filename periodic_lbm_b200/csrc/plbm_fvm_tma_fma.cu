// plbm_fvm_tma_fma.cu -- the TMA-pipelined FVM / DUGKS tile kernels with FMA contraction (opt-in, variant 3).
//
// The library is built with -fmad=false so that every result is bit-identical to the reference's non-FMA CPU
// arithmetic ("Parity and FMA").  The memory-bound kernels lose nothing by that; the DUGKS step is
// bound by the fp64 pipe (726 fp64 instructions per node, 419 DADD + 280 DMUL), and contracting a*b+c into one
// instruction removes a good part of them.  This translation unit is the same source compiled with -fmad=true
// (Makefile): results then differ from the non-FMA arithmetic in the last bits -- inside the tolerance BASELINE.json states
// (1e-12 relative fp64, 1e-5 fp32), no longer bit-identical -- so it is never the default:
// perform_dugks_step / perform_step(fvm, fdm) select it with plbm_set_variant(grid, 3).
// EXPERIMENTAL: written after round 1's GPU budget was spent; parity gate tests/test_gpu_zz_round1_late.py
// (PLBM_TEST_EXPERIMENTAL=1), A/B tools/kbench.py --case dugks,f64,bgk,3.
#define PLBM_FMA_BUILD 1
#include "plbm_fvm_tma.cu"
