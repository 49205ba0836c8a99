// TEST ONLY -- not a product path.  The per-node arithmetic headers of the CUDA library (csrc/plbm_math.cuh, csrc/plbm_fv.cuh) are
// `__device__` templates; with the stub cuda_runtime.h next to this file they compile as plain host C++ (g++ -ffp-contract=off: the
// same individually rounded operations the device build gets from -fmad=false).  tests/test_host_math_vs_refsrc.py feeds them the
// inputs of tests/golden/refsrc_*.npz node by node and demands the outputs of the reference's executed Fortran source, bit for bit:
// a pin of the PRODUCT's arithmetic source on the reference's source that needs no GPU.  What it cannot see: the kernels around
// these functions (indexing, tiling, schedules) -- that is what the -m gpu parity tests are for.  The packed fp32 type F2 is inline
// PTX and is not exercised here.
#include "plbm_math.cuh"
#include "plbm_fv.cuh"

using namespace plbm;

namespace {

template <typename T> void collide_any(int model, T (&g)[9], const CollideParams<T>& p, bool pre)
{
#define HM_CASE(M)                                              \
    case M:                                                     \
        if (pre) {                                              \
            T n[1][9], inv[1];                                  \
            for (int q = 0; q < 9; ++q) n[0][q] = g[q];         \
            node_reciprocals<T, M, 1>(n, inv);                  \
            collide_nodes<T, M, 1, false>(n, p, inv);           \
            for (int q = 0; q < 9; ++q) g[q] = n[0][q];         \
        } else {                                                \
            collide<T, M>(g, p);                                \
        }                                                       \
        break;
    switch (model) {
        HM_CASE(M_BGK)
        HM_CASE(M_TRT)
        HM_CASE(M_RR)
        HM_CASE(M_BGK_SPLIT)
        HM_CASE(M_TRT_SPLIT)
        HM_CASE(M_BGK_IMPROVED)
    }
#undef HM_CASE
}

// neighbourhood tile of one node: t[q * 9 + (dx + 1) * 3 + (dy + 1)], the layout plbm_fv.cuh addresses with PITCH = 3, PLANE = 9
template <typename T> void fv_any(int mode, int stencil, const T* tile, T dt, T omega_face, T* fp_io)
{
    T fp[9];
    for (int q = 0; q < 9; ++q) fp[q] = fp_io[q];
    const T* c0 = tile + 4;  // centre of population 0
    switch (mode) {
    case 0: flux_update<T, true, 3, 9>(c0, dt, omega_face, fp); break;    // DUGKS: faces + face relaxation + flux update
    case 1: flux_update<T, false, 3, 9>(c0, dt, omega_face, fp); break;   // Bardow FVM / periodic_dugks without -DDUGKS
    case 2: fdm_update<T, false, 3, 9>(c0, dt, fp); break;                // stream_fdm_bardow, default build
    case 3: fdm_update<T, true, 3, 9>(c0, dt, fp); break;                 // stream_fdm_sofonea
    case 4:
        switch (stencil) {
        case 1: fdm_stencil_update<T, 1, 3, 9>(c0, dt, fp); break;
        case 2: fdm_stencil_update<T, 2, 3, 9>(c0, dt, fp); break;
        case 3: fdm_stencil_update<T, 3, 3, 9>(c0, dt, fp); break;
        case 4: fdm_stencil_update<T, 4, 3, 9>(c0, dt, fp); break;
        }
        break;
    }
    for (int q = 0; q < 9; ++q) fp_io[q] = fp[q];
}

}  // namespace

#define HM_EXPORT(T, SFX)                                                                                              \
    extern "C" void hm_collide_##SFX(int model, T* f, T omega, T lambda_d, int pre)                                    \
    {                                                                                                                  \
        T g[9];                                                                                                        \
        for (int q = 0; q < 9; ++q) g[q] = f[q];                                                                       \
        collide_any<T>(model, g, CollideParams<T>{omega, lambda_d}, pre != 0);                                         \
        for (int q = 0; q < 9; ++q) f[q] = g[q];                                                                       \
    }                                                                                                                  \
    extern "C" void hm_equilibrium_##SFX(T rho, T ux, T uy, T* out)                                                    \
    {                                                                                                                  \
        T feq[9];                                                                                                      \
        equilibrium<T>(rho, ux, uy, feq);                                                                              \
        for (int q = 0; q < 9; ++q) out[q] = feq[q];                                                                   \
    }                                                                                                                  \
    extern "C" void hm_macros_##SFX(const T* f, T* out)                                                                \
    {                                                                                                                  \
        T g[9];                                                                                                        \
        for (int q = 0; q < 9; ++q) g[q] = f[q];                                                                       \
        macros<T>(g, out[0], out[1], out[2]);                                                                          \
    }                                                                                                                  \
    extern "C" void hm_fv_##SFX(int mode, int stencil, const T* tile, T dt, T omega_face, T* fp) { fv_any<T>(mode, stencil, tile, dt, omega_face, fp); }

HM_EXPORT(double, f64)
HM_EXPORT(float, f32)
