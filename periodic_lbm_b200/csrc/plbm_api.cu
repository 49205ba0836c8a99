// plbm_api.cu -- the C ABI of libplbm_b200.so (include/plbm.h): handle lifecycle, property
// derivation, step orchestration (iold/inew bookkeeping identical to the reference) and
// host <-> device staging of the macroscopic fields.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include "plbm_internal.h"

namespace plbm {

std::atomic<long long> g_launches{0};
static thread_local std::string t_error;

void set_error(const std::string& msg) { t_error = msg; }
int cuda_fail(cudaError_t e, const char* what)
{
    t_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    return PLBM_ERR_CUDA;
}

// The hidden third lattice buffer (Grid::spare).  PLBM_SPARE_LATTICE: 0 never, 1 (default) when the GPU keeps at least a quarter of
// the buffer's size + 1 GB free after it and the grid has 512^2 nodes or more (no triples below that by default), 2 whenever
// cudaMalloc succeeds (tests: with PLBM_TRIPLES=2 the closing dual triple then runs on small grids too).
void* lbm_spare(Grid& g)
{
    if (g.spare_state != 0) return g.spare_state > 0 ? g.spare : nullptr;
    static const int mode = getenv("PLBM_SPARE_LATTICE") && *getenv("PLBM_SPARE_LATTICE") ? atoi(getenv("PLBM_SPARE_LATTICE")) : 1;
    g.spare_state = -1;
    if (mode <= 0 || (mode == 1 && (long long)g.nx * g.ny < 512LL * 512LL)) return nullptr;
    const size_t bytes = g.lattice_elems() * g.esize();
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (mode == 1 && free_b < bytes + bytes / 4 + ((size_t)1 << 30)) return nullptr;
    if (cudaMalloc(&g.spare, bytes) != cudaSuccess) {
        cudaGetLastError();  // not an error of the call: the schedule without the spare is used
        g.spare = nullptr;
        return nullptr;
    }
    g.spare_state = 1;
    return g.spare;
}

void lbm_adopt_spare_as_inew(Grid& g)
{
    std::swap(g.f[g.inew - 1], g.spare);
    if (g.tmap_ok) make_tensor_maps(g);  // the descriptors name the buffers
}

namespace {

int check_noflush(plbm_handle g)
{
    if (!g) {
        set_error("null grid handle");
        return PLBM_ERR_ARG;
    }
    cudaError_t e = cudaSetDevice(g->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    return PLBM_OK;
}

// Deferred stepping (plbm_set_step_deferral): perform_lbm_step calls are only counted; the steps run, batched into
// one stepping call (two steps per pass over HBM), as soon as any other entry point looks at or changes the grid.
int flush_deferred(Grid& g);

// prologue of every entry point except the deferring ones: valid handle, its device current, no steps pending
int check(plbm_handle g)
{
    int rc = check_noflush(g);
    if (rc) return rc;
    return flush_deferred(*g);
}

int need_props(plbm_handle g)
{
    if (!g->props_set) {
        set_error("set_properties has not been called on this grid");
        return PLBM_ERR_STATE;
    }
    return PLBM_OK;
}

bool valid_model(int m) { return m >= PLBM_BGK && m <= PLBM_BGK_IMPROVED; }

void swap_lattices(Grid& g)
{
    int t = g.iold;
    g.iold = g.inew;
    g.inew = t;
}

// src/collision_trt.F90:30-34 lambda_d(omega, x), in working precision
template <typename T> T lambda_d(T omega, T x) { return (T(4) - T(2) * omega) / (T(4) * x * omega + T(2) - omega); }

template <typename T> CollideParams<T> collide_params(const Grid& g, int model)
{
    CollideParams<T> cp;
    cp.omega = (T)g.omega;
    cp.lambda_d = (model == PLBM_TRT || model == PLBM_TRT_SPLIT) ? lambda_d<T>((T)g.omega, (T)g.trt_magic) : T(0);
    return cp;
}

template <typename T> LbmArgs<T> lbm_args(const Grid& g, int src, int dst, int model)
{
    LbmArgs<T> a;
    a.src = g.lat<T>(src);
    a.dst = g.lat<T>(dst);
    a.nx = g.nx;
    a.ny = g.ny;
    a.ld = g.ld;
    a.x_begin = 0;
    a.x_end = g.nx;
    a.halo_lo = nullptr;
    a.halo_hi = nullptr;
    a.cp = collide_params<T>(g, model);
    return a;
}

// DUGKS relaxation rates, src/periodic_dugks.F90:40-44, 53, 60, 68-71, 180-181
template <typename T> void dugks_rates(const Grid& g, bool dugks, T& omega_full, T& omega_half, T& omega_face)
{
    const T tau_d = (T)g.tau / (T)g.dt;
    omega_full = (T)g.omega;  // first kernel_bgk call uses grid%omega
    T omega = T(1) / (tau_d + T(0.5));
    if (dugks) omega = T(0.75) * omega;
    omega_half = omega;
    omega_face = T(1) / (T(4) * tau_d + T(1));
}

template <typename T> int set_properties_t(Grid& g, double nu_, double dt_, double magic_, int has_magic)
{
    // src/fvm_bardow.F90:242-269, all arithmetic in working precision
    const T nu = (T)nu_, dt = (T)dt_;
    const T csqr = T(1) / T(3);
    const T invcsqr = T(1) / csqr;
    const T tau = invcsqr * nu;
    g.nu = nu;
    g.dt = dt;
    g.csqr = csqr;
    g.tau = tau;
    g.omega = dt / (tau + T(0.5) * dt);
    g.trt_magic = has_magic ? (T)magic_ : (tau / dt) * (tau / dt);
    g.props_set = true;
    return PLBM_OK;
}

// scratch fields, allocated on first use
int need_aux(Grid& g, int count)
{
    const size_t bytes = (size_t)g.nx * g.ny * g.esize();
    if (!g.aux) PLBM_CUDA(cudaMalloc(&g.aux, bytes));
    if (count > 1 && !g.aux2) PLBM_CUDA(cudaMalloc(&g.aux2, bytes));
    return PLBM_OK;
}

// Apply the deferred half-step collision of the last fused DUGKS step to lattice `inew`.
int materialize_inew(Grid& g)
{
    if (!g.dugks_pending) return PLBM_OK;
    g.dugks_pending = false;
    if (g.prec == PLBM_F64) {
        LbmArgs<double> a = lbm_args<double>(g, g.inew, g.inew, PLBM_BGK);
        a.cp.omega = g.dugks_pending_omega;
        return launch_lbm<double>(a, M_BGK_SPLIT, false, g.variant, g.stream);
    }
    LbmArgs<float> a = lbm_args<float>(g, g.inew, g.inew, PLBM_BGK);
    a.cp.omega = (float)g.dugks_pending_omega;
    return launch_lbm<float>(a, M_BGK_SPLIT, false, g.variant, g.stream);
}

template <typename T> int step_lbm_t(Grid& g, int model, int nsteps)
{
    g.dugks_pending = false;  // lattice inew is overwritten below
    if (g.comm) return comm_lbm_steps<T>(g, model, collide_params<T>(g, model), nsteps);
    if (g.variant == 0 && nsteps >= 4 && !lbm_triples_forced()) {
        // small grids: all the steps in ONE launch, lattices resident in the shared memory of a cluster
        bool done = false;
        int rc = try_lbm_cluster_steps<T>(g, g.lat<T>(g.iold), g.lat<T>(g.inew), model, collide_params<T>(g, model), nsteps, &done, g.stream);
        if (rc) return rc;
        if (done) {
            if (nsteps & 1) swap_lattices(g);
            return PLBM_OK;
        }
    }
    int s = 0;
    const bool triples = g.variant == 10 || lbm_triples_wanted(g, lbm_triples_level(g), model);
    if (g.variant == 0 && triples && nsteps >= 3 && lbm_multi_shape_is_default() && lbm_spare(g)) {
        // the default schedule when a third lattice buffer is available: triples, closed by a triple that stores the states after
        // its second and third step (lattice `inew` = state n-1, `iold` = state n, exactly what a closing single step leaves)
        const CollideParams<T> cp = collide_params<T>(g, model);
        const bool pairs = lbm_pair_applicable(g);
        while (s < nsteps) {
            const LbmLaunch L = lbm_next_launch(nsteps - s, true, pairs, true);
            int rc;
            if (L.depth == 3) {
                rc = launch_lbm_multi<T>(g, g.lat<T>(g.iold), g.lat<T>(g.inew), 0, g.nx, model, cp, 3, g.stream, nullptr, nullptr,
                                         L.dual ? static_cast<T*>(g.spare) : nullptr);
                if (rc) return rc;
                swap_lattices(g);
                if (L.dual) lbm_adopt_spare_as_inew(g);
            } else if (L.depth == 2) {
                if ((rc = launch_lbm_pair<T>(g, g.lat<T>(g.iold), g.lat<T>(g.inew), 0, g.nx, nullptr, nullptr, model, cp, g.stream))) return rc;
                std::swap(g.f[g.iold - 1], g.f[g.inew - 1]);
                for (int b = 0; b < 128; ++b) std::swap(g.tmap[g.iold - 1][b], g.tmap[g.inew - 1][b]);
            } else {
                LbmArgs<T> a = lbm_args<T>(g, g.iold, g.inew, model);
                if ((rc = launch_lbm<T>(a, model, true, g.variant, g.stream))) return rc;
                swap_lattices(g);
            }
            s += L.depth;
        }
        return PLBM_OK;
    }
    if ((g.variant == 9 || triples) && lbm_multi_applicable(g, model, triples ? 3 : 2)) {
        // the depth-generic multi-step kernel.  9: pairs through its NSTEP = 2 instance (measurement);
        // triples (three reference swaps = one swap of the indices; the result sits in lattice `inew`), then pairs.
        const CollideParams<T> cp = collide_params<T>(g, model);
        if (triples) {
            const bool pairs_follow = g.variant == 10 || (lbm_pair_variant(g.variant) && lbm_pair_applicable(g));
            for (; lbm_next_depth(nsteps - 1 - s, true, pairs_follow) == 3; s += 3) {
                int rc = launch_lbm_multi<T>(g, g.lat<T>(g.iold), g.lat<T>(g.inew), 0, g.nx, model, cp, 3, g.stream);
                if (rc) return rc;
                swap_lattices(g);
            }
        }
        // (operators the NSTEP = 2 instance is not built for fall through to the two-step kernels below)
        for (; (g.variant == 9 || g.variant == 10) && lbm_multi_applicable(g, model, 2) && s + 2 < nsteps; s += 2) {
            int rc = launch_lbm_multi<T>(g, g.lat<T>(g.iold), g.lat<T>(g.inew), 0, g.nx, model, cp, 2, g.stream);
            if (rc) return rc;
            std::swap(g.f[g.iold - 1], g.f[g.inew - 1]);
            for (int b = 0; b < 128; ++b) std::swap(g.tmap[g.iold - 1][b], g.tmap[g.inew - 1][b]);
        }
    }
    if (lbm_pair_variant(g.variant) && lbm_pair_applicable(g)) {  // 5..10: also on grids the cluster kernel would take
        // two steps per pass over HBM.  The last step stays single so that lattice `inew` ends up holding
        // state n-1 exactly as in the reference (what the lagged update_macros reads).  A pair leaves its
        // result in the buffer that was `inew`; two reference swaps leave the indices unchanged, so the
        // buffers trade places instead.
        const CollideParams<T> cp = collide_params<T>(g, model);
        for (; s + 2 < nsteps; s += 2) {
            const T* src = g.lat<T>(g.iold);
            T* dst = g.lat<T>(g.inew);
            int rc;
            if (g.variant == 8 && g.nx >= 8) {
                // test knob: the launch sequence of the slab decomposition (boundary lines, then the interior) on one GPU
                rc = launch_lbm_pair_boundaries<T>(g, src, dst, 2, nullptr, nullptr, model, cp, g.stream);
                if (!rc) rc = launch_lbm_pair<T>(g, src, dst, 2, g.nx - 2, nullptr, nullptr, model, cp, g.stream);
            } else if (g.variant == 11) {
                rc = launch_lbm_pair_fma<T>(g, src, dst, 0, g.nx, nullptr, nullptr, model, cp, g.stream);  // opt-in, FMA-contracted
            } else {
                rc = launch_lbm_pair<T>(g, src, dst, 0, g.nx, nullptr, nullptr, model, cp, g.stream);
            }
            if (rc) return rc;
            std::swap(g.f[g.iold - 1], g.f[g.inew - 1]);
            for (int b = 0; b < 128; ++b) std::swap(g.tmap[g.iold - 1][b], g.tmap[g.inew - 1][b]);
        }
    }
    for (; s < nsteps; ++s) {
        if (g.variant == 3 && g.tmap_ok && model <= PLBM_RR) {  // measurement variant: TMA-staged halo tile
            int rc = launch_fv_tma<T>(g, g.iold, g.lat<T>(g.iold), g.lat<T>(g.inew), 3, model, T(0), T(0), T(0), T(0),
                                      collide_params<T>(g, model), g.stream);
            if (rc) return rc;
            swap_lattices(g);
            continue;
        }
        LbmArgs<T> a = lbm_args<T>(g, g.iold, g.inew, model);
        int rc = launch_lbm<T>(a, model, true, g.variant, g.stream);
        if (rc) return rc;
        swap_lattices(g);
    }
    return PLBM_OK;
}

int flush_deferred(Grid& g)
{
    if (g.deferred_steps == 0) return PLBM_OK;
    const int n = g.deferred_steps, model = g.deferred_model;
    g.deferred_steps = 0;
    return g.prec == PLBM_F64 ? step_lbm_t<double>(g, model, n) : step_lbm_t<float>(g, model, n);
}

// perform_lbm_step behind the deferral: count the steps, run them when the budget is reached or the collision changes
int lbm_steps_entry(Grid& g, int collision, int nsteps)
{
    int rc;
    if (g.defer_max > 0 && !g.comm && nsteps > 0 && nsteps < g.defer_max) {
        if (g.deferred_steps > 0 && g.deferred_model != collision && (rc = flush_deferred(g))) return rc;
        g.deferred_model = collision;
        g.deferred_steps += nsteps;
        return g.deferred_steps >= g.defer_max ? flush_deferred(g) : PLBM_OK;
    }
    if ((rc = flush_deferred(g))) return rc;
    return g.prec == PLBM_F64 ? step_lbm_t<double>(g, collision, nsteps) : step_lbm_t<float>(g, collision, nsteps);
}

// kernel mode of a finite-volume / finite-difference streaming scheme (plbm_fvm*.cu)
int fv_mode(int streaming) { return streaming == PLBM_STREAM_FVM_BARDOW ? 2 : (streaming == PLBM_STREAM_FDM_BARDOW ? 4 : 5); }

// One sweep iold -> inew of stream_fvm_bardow / stream_fdm_bardow / stream_fdm_sofonea, fused with
// the collision `model` (M_NONE = streaming only).  Collisions the tile kernels do not instantiate for
// a scheme are applied by a second, in-place launch.
template <typename T> int fv_sweep(Grid& g, int streaming, int model, const CollideParams<T>& cp)
{
    int rc;
    const int mode = fv_mode(streaming);
    const bool fusable = model == M_NONE || mode == 2 || model <= PLBM_RR;
    const int kmodel = fusable ? model : (int)M_NONE;
    if (streaming == PLBM_STREAM_FDM_BARDOW && g.fdm_stencil != 0) {
        // the reference's -DFDM_WLS* / -DFDM_ISO builds of stream_fdm_bardow: plain-load tile kernel, collision as a second launch
        if (g.comm) {
            set_error("stream_fdm_bardow: the alternative stencils are not available under a slab decomposition");
            return PLBM_ERR_ARG;
        }
        rc = launch_fvm_bardow<T>(g, g.lat<T>(g.iold), g.lat<T>(g.inew), (T)g.dt, M_NONE, cp, g.stream, 5 + g.fdm_stencil);
        if (rc || model == M_NONE) return rc;
        LbmArgs<T> c = lbm_args<T>(g, g.inew, g.inew, model);
        return launch_lbm<T>(c, model, false, g.variant, g.stream);
    }
    if (g.comm && (rc = comm_fv_exchange<T>(g, g.lat<T>(g.iold)))) return rc;  // slab: neighbours' boundary lines
    if (g.variant == 4 && fv_march_applicable(mode, kmodel))  // opt-in: marching kernel with shared faces (within tolerance, not bit-identical)
        rc = launch_fv_march<T>(g, g.lat<T>(g.iold), g.lat<T>(g.inew), mode, kmodel, (T)g.dt, T(0), T(0), T(0), cp, g.stream);
    else if (g.variant == 3 && !g.comm && g.tmap_ok)  // opt-in: the tile kernel with FMA contraction (within tolerance, not bit-identical)
        rc = launch_fv_tma_fma<T>(g, g.iold, g.lat<T>(g.iold), g.lat<T>(g.inew), mode, kmodel, (T)g.dt, T(0), T(0), T(0), cp, g.stream);
    else if ((g.variant == 0 || g.comm) && g.tmap_ok)  // TMA + mbarrier pipelined tile kernel
        rc = launch_fv_tma<T>(g, g.iold, g.lat<T>(g.iold), g.lat<T>(g.inew), mode, kmodel, (T)g.dt, T(0), T(0), T(0), cp, g.stream);
    else
        rc = launch_fvm_bardow<T>(g, g.lat<T>(g.iold), g.lat<T>(g.inew), (T)g.dt, kmodel, cp, g.stream, mode);
    if (rc) return rc;
    if (!fusable) {
        LbmArgs<T> c = lbm_args<T>(g, g.inew, g.inew, model);
        rc = launch_lbm<T>(c, model, false, g.variant, g.stream);
    }
    return rc;
}

template <typename T> int step_fvm_t(Grid& g, int streaming, int model, int nsteps)
{
    const CollideParams<T> cp = collide_params<T>(g, model);
    g.dugks_pending = false;  // lattice inew is overwritten below
    for (int s = 0; s < nsteps; ++s) {
        int rc = fv_sweep<T>(g, streaming, model, cp);
        if (rc) return rc;
        swap_lattices(g);
    }
    return PLBM_OK;
}

template <typename T> int step_dugks_t(Grid& g, bool dugks, int nsteps)
{
    T of, oh, oc;
    dugks_rates<T>(g, dugks, of, oh, oc);
    for (int s = 0; s < nsteps; ++s) {
        int rc;
        g.dugks_pending = false;  // lattice inew is overwritten below
        if (g.comm && (rc = comm_fv_exchange<T>(g, g.lat<T>(g.iold)))) return rc;  // slab: neighbours' boundary lines
        if (g.variant == 1 && !g.comm) {  // reference structure: collide pass + stream pass
            rc = launch_dugks_collide<T>(g, g.lat<T>(g.iold), g.lat<T>(g.inew), of, oh, g.stream);
            if (rc) return rc;
            rc = launch_dugks_stream<T>(g, g.lat<T>(g.iold), g.lat<T>(g.inew), (T)g.dt, oc, dugks, g.stream);
        } else if (g.variant == 4) {
            // opt-in: marching kernel, every face reconstructed and relaxed once and shared by its two cells (within tolerance
            // of the reference, not bit-identical; plbm_fvm_march.cu)
            rc = launch_fv_march<T>(g, g.lat<T>(g.iold), g.lat<T>(g.inew), dugks ? 0 : 1, M_NONE, (T)g.dt, of, oh, oc,
                                    CollideParams<T>{T(0), T(0)}, g.stream);
        } else if (g.variant == 3 && !g.comm && g.tmap_ok) {
            // opt-in: the same kernel with FMA contraction (within tolerance of the non-FMA result, not bit-identical)
            rc = launch_fv_tma_fma<T>(g, g.iold, g.lat<T>(g.iold), g.lat<T>(g.inew), dugks ? 0 : 1, M_NONE, (T)g.dt, of, oh, oc,
                                      CollideParams<T>{T(0), T(0)}, g.stream);
        } else if ((g.variant == 0 || g.comm) && g.tmap_ok) {
            // fused, TMA-pipelined: lattice iold (ftilde^n) is left untouched, inew receives ftilde^{n+1}.
            rc = launch_fv_tma<T>(g, g.iold, g.lat<T>(g.iold), g.lat<T>(g.inew), dugks ? 0 : 1, M_NONE, (T)g.dt, of, oh, oc,
                                  CollideParams<T>{T(0), T(0)}, g.stream);
        } else {
            // fused, plain loads (variant 2, or no TMA descriptor)
            rc = launch_dugks_fused<T>(g, g.lat<T>(g.iold), g.lat<T>(g.inew), (T)g.dt, of, oh, oc, dugks, g.stream);
        }
        if (rc) return rc;
        swap_lattices(g);
        if (!(g.variant == 1 && !g.comm)) {
            g.dugks_pending = true;
            g.dugks_pending_omega = (double)oh;
        }
    }
    return PLBM_OK;
}

// perform_triple_step (src/fvm_bardow.F90:322-340): stream iold -> inew; copy inew -> iold (keep the
// pre-collision PDFs); collide inew; rotate (iold, inew, imid) <- (inew, imid, iold).
template <typename T> int step_triple_t(Grid& g, int streaming, int model, int nsteps)
{
    g.dugks_pending = false;
    const CollideParams<T> cp = collide_params<T>(g, model);
    for (int s = 0; s < nsteps; ++s) {
        int rc;
        if (streaming == PLBM_STREAM_LBM) {
            // fused: post-collision -> inew, pre-collision -> the spare lattice (imid); the spare and
            // the source then trade places so that index `iold` ends up holding the pre-collision copy
            LbmArgs<T> a = lbm_args<T>(g, g.iold, g.inew, model);
            a.pre = g.lat<T>(g.imid);
            if ((rc = launch_lbm<T>(a, model, true, g.variant, g.stream))) return rc;
            std::swap(g.f[g.iold - 1], g.f[g.imid - 1]);
            for (int b = 0; b < 128; ++b) std::swap(g.tmap[g.iold - 1][b], g.tmap[g.imid - 1][b]);
        } else {
            if ((rc = fv_sweep<T>(g, streaming, M_NONE, cp))) return rc;
            PLBM_CUDA(cudaMemcpyAsync(g.f[g.iold - 1], g.f[g.inew - 1], g.lattice_elems() * g.esize(), cudaMemcpyDeviceToDevice, g.stream));
            if ((rc = launch_lbm<T>(lbm_args<T>(g, g.inew, g.inew, model), model, false, g.variant, g.stream))) return rc;
        }
        const int t = g.iold;
        g.iold = g.inew;
        g.inew = g.imid;
        g.imid = t;
    }
    return PLBM_OK;
}

template <typename T> int upload_field(Grid& g, T* dev, const void* host)
{
    PLBM_CUDA(cudaMemcpyAsync(dev, host, sizeof(T) * (size_t)g.nx * g.ny, cudaMemcpyHostToDevice, g.stream));
    return PLBM_OK;
}
template <typename T> int download_field(Grid& g, void* host, const T* dev)
{
    if (!host) return PLBM_OK;
    PLBM_CUDA(cudaMemcpyAsync(host, dev, sizeof(T) * (size_t)g.nx * g.ny, cudaMemcpyDeviceToHost, g.stream));
    return PLBM_OK;
}

template <typename T> int set_pdf_to_equilibrium_t(Grid& g, const void* rho, const void* ux, const void* uy)
{
    int rc;
    if ((rc = upload_field<T>(g, g.rho<T>(), rho))) return rc;
    if ((rc = upload_field<T>(g, g.ux<T>(), ux))) return rc;
    if ((rc = upload_field<T>(g, g.uy<T>(), uy))) return rc;
    return launch_init_eq<T>(g, g.lat<T>(g.iold), g.stream);
}

template <typename T> static int vorticity_t(Grid& g, int order, void* omega)
{
    int rc;
    const T *lo = nullptr, *hi = nullptr;
    // slab decomposition: the neighbours' two nearest lines of uy (every rank of the ring calls this)
    if (g.comm && (rc = comm_field_exchange2<T>(g, g.uy<T>(), &lo, &hi))) return rc;
    if ((rc = launch_vorticity<T>(g, order, g.ux<T>(), g.uy<T>(), (T*)g.aux, g.stream, lo, hi))) return rc;
    return download_field<T>(g, omega, (T*)g.aux);
}

}  // namespace
}  // namespace plbm

using namespace plbm;

#define DISPATCH(g, expr_d, expr_f) ((g)->prec == PLBM_F64 ? (expr_d) : (expr_f))

extern "C" {

const char* plbm_last_error(void) { return t_error.c_str(); }
int plbm_version(void) { return 100; }
long long plbm_launch_count(void) { return g_launches.load(); }

int plbm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int plbm_alloc_grid_on(plbm_handle* out, int nx, int ny, int nf, int precision, int device)
{
    if (!out || nx < 1 || ny < 1 || (nf != 2 && nf != 3) || (precision != PLBM_F64 && precision != PLBM_F32)) {
        set_error("alloc_grid: bad argument (need nx,ny >= 1, nf in {2,3}, precision in {F64,F32})");
        return PLBM_ERR_ARG;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device available: libplbm_b200 has no CPU fallback");
        return PLBM_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) {
        set_error("alloc_grid: device ordinal out of range");
        return PLBM_ERR_ARG;
    }
    PLBM_CUDA(cudaSetDevice(device));
    plbm_grid_s* g = new plbm_grid_s();
    g->nx = nx;
    g->ny = ny;
    g->nx_global = nx;
    // round the unit-stride dimension up to a multiple of 16 (src/fvm_bardow.F90:144-147)
    g->ld = (ny + 15) / 16 * 16;
    g->nf = nf;
    g->prec = precision;
    g->device = device;
    g->inew = 1;
    g->iold = 2;
    g->imid = nf > 2 ? 3 : -1;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) g->sm_count = prop.multiProcessorCount;
    const size_t es = g->esize();
    const size_t nm = (size_t)nx * ny;
    auto fail = [&](cudaError_t err, const char* what) {
        int rc = cuda_fail(err, what);
        plbm_dealloc_grid(g);
        return rc;
    };
    for (int i = 0; i < nf; ++i) {
        e = cudaMalloc(&g->f[i], g->lattice_elems() * es);
        if (e != cudaSuccess) return fail(e, "cudaMalloc(lattice)");
    }
    if ((e = cudaMalloc(&g->mf, 3 * nm * es)) != cudaSuccess) return fail(e, "cudaMalloc(mf)");
    // aux / aux2 (vorticity output, analytic fields) are allocated on first use: a 32768^2 fp64 grid
    // fills the GPU with its two lattices + macroscopic fields (180.4 GB of 192 GB)
    g->npartial = 4 * g->sm_count;
    if ((e = cudaMalloc(&g->partial, 32 * (size_t)g->npartial)) != cudaSuccess) return fail(e, "cudaMalloc(partial)");
    if ((e = cudaMallocHost(&g->partial_host, 32 * (size_t)g->npartial)) != cudaSuccess) return fail(e, "cudaMallocHost");
    if ((e = cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate");
    g->own_stream = true;
    make_tensor_maps(*g);
    *out = g;
    return PLBM_OK;
}

int plbm_alloc_grid(plbm_handle* out, int nx, int ny, int nf, int precision)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        dev = 0;
    }
    return plbm_alloc_grid_on(out, nx, ny, nf, precision, dev);
}

int plbm_dealloc_grid(plbm_handle g)
{
    if (!g) return PLBM_OK;
    cudaSetDevice(g->device);
    if (g->comm) comm_finalize(*g);
    if (g->stream) cudaStreamSynchronize(g->stream);
    for (int i = 0; i < 3; ++i)
        if (g->f[i]) cudaFree(g->f[i]);
    if (g->spare) cudaFree(g->spare);
    if (g->mf) cudaFree(g->mf);
    if (g->aux) cudaFree(g->aux);
    if (g->aux2) cudaFree(g->aux2);
    if (g->partial) cudaFree(g->partial);
    if (g->partial_host) cudaFreeHost(g->partial_host);
    if (g->own_stream && g->stream) cudaStreamDestroy(g->stream);
    delete g;
    return PLBM_OK;
}

int plbm_get_dims(plbm_handle g, int* nx, int* ny, int* ld, int* nf, int* precision)
{
    if (!g) {
        set_error("null grid handle");
        return PLBM_ERR_ARG;
    }
    if (nx) *nx = g->nx;
    if (ny) *ny = g->ny;
    if (ld) *ld = g->ld;
    if (nf) *nf = g->nf;
    if (precision) *precision = g->prec;
    return PLBM_OK;
}

int plbm_get_indices(plbm_handle g, int* iold, int* inew, int* imid)
{
    if (!g) {
        set_error("null grid handle");
        return PLBM_ERR_ARG;
    }
    // an odd number of deferred steps swaps the roles of the two lattices, like the steps themselves will
    const bool odd = (g->deferred_steps & 1) != 0;
    if (iold) *iold = odd ? g->inew : g->iold;
    if (inew) *inew = odd ? g->iold : g->inew;
    if (imid) *imid = g->imid;
    return PLBM_OK;
}

int plbm_set_indices(plbm_handle g, int iold, int inew, int imid)
{
    int rc = check(g);
    if (rc) return rc;
    // a permutation of the lattices the grid owns: (iold, inew) of {1, 2} for nf = 2 (imid ignored, stays -1),
    // (iold, inew, imid) of {1, 2, 3} for nf = 3 -- the rotations perform_triple_step reaches and the others alike
    const bool ok2 = g->nf == 2 && ((iold == 1 && inew == 2) || (iold == 2 && inew == 1));
    const bool ok3 = g->nf == 3 && iold >= 1 && iold <= 3 && inew >= 1 && inew <= 3 && imid >= 1 && imid <= 3 && iold != inew && iold != imid &&
                     inew != imid;
    if (!ok2 && !ok3) {
        set_error("set_indices: (iold, inew[, imid]) must be a permutation of the grid's lattice numbers");
        return PLBM_ERR_ARG;
    }
    if ((rc = materialize_inew(*g))) return rc;  // the pending half-step collision belongs to the OLD inew
    g->iold = iold;
    g->inew = inew;
    if (g->nf == 3) g->imid = imid;
    comm_invalidate_halo(*g);
    return PLBM_OK;
}

int plbm_set_properties(plbm_handle g, double nu, double dt, double magic, int has_magic)
{
    if (!g) {
        set_error("null grid handle");
        return PLBM_ERR_ARG;
    }
    if (g->deferred_steps > 0) {  // pending steps use the old relaxation rates
        int rc = check(g);
        if (rc) return rc;
    }
    if (!(dt > 0) || !(nu >= 0)) {
        set_error("set_properties: need dt > 0 and nu >= 0");
        return PLBM_ERR_ARG;
    }
    return DISPATCH(g, set_properties_t<double>(*g, nu, dt, magic, has_magic), set_properties_t<float>(*g, nu, dt, magic, has_magic));
}

int plbm_get_properties(plbm_handle g, double out[6])
{
    if (!g || !out) {
        set_error("get_properties: null argument");
        return PLBM_ERR_ARG;
    }
    int rc = need_props(g);
    if (rc) return rc;
    out[0] = g->nu;
    out[1] = g->dt;
    out[2] = g->tau;
    out[3] = g->omega;
    out[4] = g->trt_magic;
    out[5] = g->csqr;
    return PLBM_OK;
}

int plbm_set_omega(plbm_handle g, double omega)
{
    if (!g) {
        set_error("null grid handle");
        return PLBM_ERR_ARG;
    }
    const double w = g->prec == PLBM_F64 ? omega : (double)(float)omega;
    if (w != g->omega && g->deferred_steps > 0) {  // pending steps use the old rate (the Fortran shim re-sends omega before every step)
        int rc = check(g);
        if (rc) return rc;
    }
    g->omega = w;
    return PLBM_OK;
}

int plbm_set_pdf_to_equilibrium(plbm_handle g, const void* rho, const void* ux, const void* uy)
{
    int rc = check(g);
    if (rc) return rc;
    if (!rho || !ux || !uy) {
        set_error("set_pdf_to_equilibrium: null field pointer");
        return PLBM_ERR_ARG;
    }
    if ((rc = materialize_inew(*g))) return rc;
    comm_invalidate_halo(*g);
    rc = DISPATCH(g, set_pdf_to_equilibrium_t<double>(*g, rho, ux, uy), set_pdf_to_equilibrium_t<float>(*g, rho, ux, uy));
    if (rc) return rc;
    // the host buffers are borrowed for the call only
    PLBM_CUDA(cudaStreamSynchronize(g->stream));
    return PLBM_OK;
}

int plbm_perform_lbm_step(plbm_handle g, int collision, int nsteps)
{
    int rc = check_noflush(g);
    if (rc) return rc;
    if ((rc = need_props(g))) return rc;
    if (!valid_model(collision) || nsteps < 0) {
        set_error("perform_lbm_step: bad collision id or nsteps");
        return PLBM_ERR_ARG;
    }
    return lbm_steps_entry(*g, collision, nsteps);
}

int plbm_set_step_deferral(plbm_handle g, int max_pending)
{
    int rc = check(g);  // runs what is pending under the old setting
    if (rc) return rc;
    if (max_pending < 0) {
        set_error("set_step_deferral: max_pending must be >= 0");
        return PLBM_ERR_ARG;
    }
    g->defer_max = max_pending;
    return PLBM_OK;
}

int plbm_perform_step(plbm_handle g, int streaming, int collision, int nsteps)
{
    if (streaming == PLBM_STREAM_LBM) return plbm_perform_lbm_step(g, collision, nsteps);
    int rc = check(g);
    if (rc) return rc;
    if ((rc = need_props(g))) return rc;
    if (streaming < PLBM_STREAM_FVM_BARDOW || streaming > PLBM_STREAM_FDM_SOFONEA || !valid_model(collision) || nsteps < 0) {
        set_error("perform_step: bad streaming/collision id or nsteps");
        return PLBM_ERR_ARG;
    }
    if (g->comm && !g->tmap_ok) {
        set_error("perform_step(fvm_bardow): the slab decomposition needs the TMA tile kernel (no tensor map on this device)");
        return PLBM_ERR_ARG;
    }
    return DISPATCH(g, step_fvm_t<double>(*g, streaming, collision, nsteps), step_fvm_t<float>(*g, streaming, collision, nsteps));
}

int plbm_perform_triple_step(plbm_handle g, int streaming, int collision, int nsteps)
{
    int rc = check(g);
    if (rc) return rc;
    if ((rc = need_props(g))) return rc;
    if (g->nf < 3) {
        set_error("perform_triple_step: the grid was allocated with nf = 2");
        return PLBM_ERR_STATE;
    }
    if (streaming < PLBM_STREAM_LBM || streaming > PLBM_STREAM_FDM_SOFONEA || !valid_model(collision) || nsteps < 0) {
        set_error("perform_triple_step: bad streaming/collision id or nsteps");
        return PLBM_ERR_ARG;
    }
    if (g->comm) {
        set_error("perform_triple_step: slab decomposition not supported for this orchestrator");
        return PLBM_ERR_ARG;
    }
    return DISPATCH(g, step_triple_t<double>(*g, streaming, collision, nsteps), step_triple_t<float>(*g, streaming, collision, nsteps));
}

int plbm_perform_dugks_step(plbm_handle g, int dugks, int nsteps)
{
    int rc = check(g);
    if (rc) return rc;
    if ((rc = need_props(g))) return rc;
    if (nsteps < 0) {
        set_error("perform_dugks_step: bad nsteps");
        return PLBM_ERR_ARG;
    }
    if (g->comm && !g->tmap_ok) {
        set_error("perform_dugks_step: the slab decomposition needs the TMA tile kernel (no tensor map on this device)");
        return PLBM_ERR_ARG;
    }
    return DISPATCH(g, step_dugks_t<double>(*g, dugks != 0, nsteps), step_dugks_t<float>(*g, dugks != 0, nsteps));
}

int plbm_lbm_stream(plbm_handle g)
{
    int rc = check(g);
    if (rc) return rc;
    if (g->comm) {
        set_error("lbm_stream: unfused entry points are single-GPU only");
        return PLBM_ERR_ARG;
    }
    g->dugks_pending = false;  // lattice inew is overwritten
    if (g->prec == PLBM_F64) return launch_lbm<double>(lbm_args<double>(*g, g->iold, g->inew, PLBM_BGK), M_NONE, true, g->variant, g->stream);
    return launch_lbm<float>(lbm_args<float>(*g, g->iold, g->inew, PLBM_BGK), M_NONE, true, g->variant, g->stream);
}

static int stream_only(plbm_handle g, int streaming)
{
    int rc = check(g);
    if (rc) return rc;
    if ((rc = need_props(g))) return rc;
    if (g->comm && !g->tmap_ok) {
        set_error("streaming: the slab decomposition needs the TMA tile kernel");
        return PLBM_ERR_ARG;
    }
    g->dugks_pending = false;  // lattice inew is overwritten
    if (g->prec == PLBM_F64) return fv_sweep<double>(*g, streaming, M_NONE, CollideParams<double>{0, 0});
    return fv_sweep<float>(*g, streaming, M_NONE, CollideParams<float>{0, 0});
}

int plbm_stream_fvm_bardow(plbm_handle g) { return stream_only(g, PLBM_STREAM_FVM_BARDOW); }
int plbm_stream_fdm_bardow(plbm_handle g) { return stream_only(g, PLBM_STREAM_FDM_BARDOW); }
int plbm_stream_fdm_sofonea(plbm_handle g) { return stream_only(g, PLBM_STREAM_FDM_SOFONEA); }

int plbm_collide(plbm_handle g, int collision)
{
    int rc = check(g);
    if (rc) return rc;
    if ((rc = need_props(g))) return rc;
    if (!valid_model(collision)) {
        set_error("collide: bad collision id");
        return PLBM_ERR_ARG;
    }
    if ((rc = materialize_inew(*g))) return rc;
    // in place on lattice inew, like collide_bgk/trt/rr
    if (g->prec == PLBM_F64) return launch_lbm<double>(lbm_args<double>(*g, g->inew, g->inew, collision), collision, false, g->variant, g->stream);
    return launch_lbm<float>(lbm_args<float>(*g, g->inew, g->inew, collision), collision, false, g->variant, g->stream);
}

int plbm_dugks_collide(plbm_handle g, int dugks)
{
    int rc = check(g);
    if (rc) return rc;
    if ((rc = need_props(g))) return rc;
    g->dugks_pending = false;  // lattice inew is overwritten
    if (g->prec == PLBM_F64) {
        double of, oh, oc;
        dugks_rates<double>(*g, dugks != 0, of, oh, oc);
        return launch_dugks_collide<double>(*g, g->lat<double>(g->iold), g->lat<double>(g->inew), of, oh, g->stream);
    }
    float of, oh, oc;
    dugks_rates<float>(*g, dugks != 0, of, oh, oc);
    return launch_dugks_collide<float>(*g, g->lat<float>(g->iold), g->lat<float>(g->inew), of, oh, g->stream);
}

int plbm_dugks_stream(plbm_handle g, int dugks)
{
    int rc = check(g);
    if (rc) return rc;
    if ((rc = need_props(g))) return rc;
    // after a fused perform_dugks_step lattice inew still holds ftilde^n: the reference has fbar^+ there, and an
    // unfused dugks_stream reads and updates exactly that lattice
    if ((rc = materialize_inew(*g))) return rc;
    if (g->prec == PLBM_F64) {
        double of, oh, oc;
        dugks_rates<double>(*g, dugks != 0, of, oh, oc);
        return launch_dugks_stream<double>(*g, g->lat<double>(g->iold), g->lat<double>(g->inew), (double)g->dt, oc, dugks != 0, g->stream);
    }
    float of, oh, oc;
    dugks_rates<float>(*g, dugks != 0, of, oh, oc);
    return launch_dugks_stream<float>(*g, g->lat<float>(g->iold), g->lat<float>(g->inew), (float)g->dt, oc, dugks != 0, g->stream);
}

int plbm_swap(plbm_handle g)
{
    if (!g) {
        set_error("null grid handle");
        return PLBM_ERR_ARG;
    }
    int rc = check(g);
    if (rc) return rc;
    if ((rc = materialize_inew(*g))) return rc;
    swap_lattices(*g);
    return PLBM_OK;
}

int plbm_update_macros(plbm_handle g, void* rho, void* ux, void* uy, int lagged)
{
    int rc = check(g);
    if (rc) return rc;
    const int which = lagged ? g->inew : g->iold;
    if (lagged && (rc = materialize_inew(*g))) return rc;
    if (g->prec == PLBM_F64) {
        if ((rc = launch_macros<double>(*g, g->lat<double>(which), g->stream))) return rc;
        if ((rc = download_field<double>(*g, rho, g->rho<double>()))) return rc;
        if ((rc = download_field<double>(*g, ux, g->ux<double>()))) return rc;
        if ((rc = download_field<double>(*g, uy, g->uy<double>()))) return rc;
    } else {
        if ((rc = launch_macros<float>(*g, g->lat<float>(which), g->stream))) return rc;
        if ((rc = download_field<float>(*g, rho, g->rho<float>()))) return rc;
        if ((rc = download_field<float>(*g, ux, g->ux<float>()))) return rc;
        if ((rc = download_field<float>(*g, uy, g->uy<float>()))) return rc;
    }
    if (rho || ux || uy) PLBM_CUDA(cudaStreamSynchronize(g->stream));
    return PLBM_OK;
}

int plbm_vorticity(plbm_handle g, int order, void* omega)
{
    int rc = check(g);
    if (rc) return rc;
    if (order != 2 && order != 4) {
        set_error("vorticity: order must be 2 or 4");
        return PLBM_ERR_ARG;
    }
    if ((rc = need_aux(*g, 1))) return rc;
    if ((rc = DISPATCH(g, vorticity_t<double>(*g, order, omega), vorticity_t<float>(*g, order, omega)))) return rc;
    if (omega) PLBM_CUDA(cudaStreamSynchronize(g->stream));
    return PLBM_OK;
}

int plbm_vorticity_host(plbm_handle g, int order, const void* ux, const void* uy, void* omega)
{
    int rc = check(g);
    if (rc) return rc;
    if (!ux || !uy || !omega) {
        set_error("vorticity_host: null pointer");
        return PLBM_ERR_ARG;
    }
    // stage through the device macroscopic fields (overwrites ux, uy)
    if (g->prec == PLBM_F64) {
        if ((rc = upload_field<double>(*g, g->ux<double>(), ux))) return rc;
        if ((rc = upload_field<double>(*g, g->uy<double>(), uy))) return rc;
    } else {
        if ((rc = upload_field<float>(*g, g->ux<float>(), ux))) return rc;
        if ((rc = upload_field<float>(*g, g->uy<float>(), uy))) return rc;
    }
    return plbm_vorticity(g, order, omega);
}

int plbm_diagnostics(plbm_handle g, double out[PLBM_DIAG_COUNT])
{
    int rc = check(g);
    if (rc) return rc;
    if (!out) {
        set_error("diagnostics: null output");
        return PLBM_ERR_ARG;
    }
    return DISPATCH(g, launch_diagnostics<double>(*g, out, g->stream), launch_diagnostics<float>(*g, out, g->stream));
}

int plbm_l2_sums(plbm_handle g, const void* uxa, const void* uya, double out[2])
{
    int rc = check(g);
    if (rc) return rc;
    if (!uxa || !uya || !out) {
        set_error("l2_sums: null pointer");
        return PLBM_ERR_ARG;
    }
    if ((rc = need_aux(*g, 2))) return rc;
    if (g->prec == PLBM_F64) {
        if ((rc = upload_field<double>(*g, (double*)g->aux, uxa))) return rc;
        if ((rc = upload_field<double>(*g, (double*)g->aux2, uya))) return rc;
        return launch_l2_sums<double>(*g, (const double*)g->aux, (const double*)g->aux2, out, g->stream);
    }
    if ((rc = upload_field<float>(*g, (float*)g->aux, uxa))) return rc;
    if ((rc = upload_field<float>(*g, (float*)g->aux2, uya))) return rc;
    return launch_l2_sums<float>(*g, (const float*)g->aux, (const float*)g->aux2, out, g->stream);
}

int plbm_lattice_hash(plbm_handle g, int which, unsigned long long* out)
{
    int rc = check(g);
    if (rc) return rc;
    if (which < 1 || which > g->nf || !out) {
        set_error("lattice_hash: bad lattice index or null pointer");
        return PLBM_ERR_ARG;
    }
    if (which == g->inew && (rc = materialize_inew(*g))) return rc;
    return DISPATCH(g, launch_lattice_hash<double>(*g, g->lat<double>(which), out, g->stream),
                    launch_lattice_hash<float>(*g, g->lat<float>(which), out, g->stream));
}

int plbm_upload_f(plbm_handle g, int which, const void* host_f)
{
    int rc = check(g);
    if (rc) return rc;
    if (which < 1 || which > g->nf || !host_f) {
        set_error("upload_f: bad lattice index or null pointer");
        return PLBM_ERR_ARG;
    }
    if (which == g->inew) g->dugks_pending = false;
    else if ((rc = materialize_inew(*g))) return rc;
    comm_invalidate_halo(*g);
    PLBM_CUDA(cudaMemcpyAsync(g->f[which - 1], host_f, g->lattice_elems() * g->esize(), cudaMemcpyHostToDevice, g->stream));
    PLBM_CUDA(cudaStreamSynchronize(g->stream));
    return PLBM_OK;
}

int plbm_download_f(plbm_handle g, int which, void* host_f)
{
    int rc = check(g);
    if (rc) return rc;
    if (which < 1 || which > g->nf || !host_f) {
        set_error("download_f: bad lattice index or null pointer");
        return PLBM_ERR_ARG;
    }
    if (which == g->inew && (rc = materialize_inew(*g))) return rc;
    PLBM_CUDA(cudaMemcpyAsync(host_f, g->f[which - 1], g->lattice_elems() * g->esize(), cudaMemcpyDeviceToHost, g->stream));
    PLBM_CUDA(cudaStreamSynchronize(g->stream));
    return PLBM_OK;
}

int plbm_set_stream(plbm_handle g, void* cuda_stream)
{
    int rc = check(g);
    if (rc) return rc;
    PLBM_CUDA(cudaStreamSynchronize(g->stream));
    if (g->own_stream) {
        cudaStreamDestroy(g->stream);
        g->own_stream = false;
    }
    if (cuda_stream) {
        g->stream = static_cast<cudaStream_t>(cuda_stream);
    } else {
        PLBM_CUDA(cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
        g->own_stream = true;
    }
    return PLBM_OK;
}

int plbm_synchronize(plbm_handle g)
{
    int rc = check(g);
    if (rc) return rc;
    PLBM_CUDA(cudaStreamSynchronize(g->stream));
    return PLBM_OK;
}

int plbm_set_fdm_stencil(plbm_handle g, int stencil)
{
    int rc = check(g);
    if (rc) return rc;
    if (stencil < PLBM_FDM_DEFAULT || stencil > PLBM_FDM_ISO) {
        set_error("set_fdm_stencil: unknown stencil");
        return PLBM_ERR_ARG;
    }
    g->fdm_stencil = stencil;
    return PLBM_OK;
}

int plbm_set_variant(plbm_handle g, int variant)
{
    if (!g) {
        set_error("null grid handle");
        return PLBM_ERR_ARG;
    }
    if (g->deferred_steps > 0) {
        int rc = check(g);
        if (rc) return rc;
    }
    g->variant = variant;
    return PLBM_OK;
}

int plbm_lbm_pair_kernel(plbm_handle g)
{
    if (!g) {
        set_error("null grid handle");
        return -1;
    }
    if (g->comm && !comm_pairs_agreed(*g)) return 0;
    return lbm_pair_flavour(*g);
}

int plbm_lbm_steps_per_pass(plbm_handle g, int collision)
{
    if (!g) {
        set_error("null grid handle");
        return -1;
    }
    int level = lbm_triples_level(*g);
    if (g->comm && !comm_triples_level(*g, &level)) level = -1;
    if (level >= 0 && lbm_multi_applicable(*g, collision, 3) && (g->variant == 10 || lbm_triples_wanted(*g, level, collision))) return 3;
    if (g->comm && !comm_pairs_agreed(*g)) return 1;
    return lbm_pair_flavour(*g) ? 2 : 1;
}

int plbm_lbm_closing_triple(plbm_handle g, int collision)
{
    if (!g) {
        set_error("null grid handle");
        return -1;
    }
    if (plbm_lbm_steps_per_pass(g, collision) != 3 || g->variant != 0 || !lbm_multi_shape_is_default()) return 0;
    if (g->comm) return comm_dual_agreed(*g) && g->spare ? 1 : 0;
    return lbm_spare(*g) ? 1 : 0;
}

int plbm_lbm_triple_kernel(plbm_handle g, int collision)
{
    if (!g) {
        set_error("null grid handle");
        return -1;
    }
    return lbm_triple_ws_wanted(collision) ? 1 : 0;
}

int plbm_comm_unique_id(void* id128) { return comm_unique_id(id128); }

int plbm_comm_init(plbm_handle g, const void* id128, int rank, int nranks, int nx_global, int x_offset)
{
    int rc = check(g);
    if (rc) return rc;
    return comm_init(*g, id128, rank, nranks, nx_global, x_offset);
}

int plbm_comm_transport(plbm_handle g)
{
    if (!g) return -1;
    return g->comm ? (comm_transport_is_p2p(*g) ? 1 : 0) : -1;
}

int plbm_comm_finalize(plbm_handle g)
{
    int rc = check(g);
    if (rc) return rc;
    return comm_finalize(*g);
}

}  // extern "C"
