// plbm_internal.h -- grid state and launcher prototypes shared by the translation units of
// libplbm_b200.so.  Not part of the public ABI (that is include/plbm.h).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstddef>
#include <cstdint>
#include <string>

#include "../../include/plbm.h"
#include "plbm_math.cuh"

namespace plbm {

extern std::atomic<long long> g_launches;  // kernels launched through the library
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

#define PLBM_CUDA(call)                                          \
    do {                                                         \
        cudaError_t _e = (call);                                 \
        if (_e != cudaSuccess) return plbm::cuda_fail(_e, #call); \
    } while (0)

struct Comm;  // multi-GPU ring (plbm_comm.cu)

// Device-resident mirror of the reference's `lattice_grid` (src/fvm_bardow.F90:37-69).
struct Grid {
    int nx = 0, ny = 0, ld = 0, nf = 2;
    int prec = PLBM_F64;
    int device = 0;
    void* f[3] = {nullptr, nullptr, nullptr};  // PDF lattices, each f(ld,nx,0:8)
    void* mf = nullptr;                        // rho, ux, uy: 3 x (ny,nx)
    void* aux = nullptr;                       // scratch field (ny,nx): vorticity / analytic upload
    void* aux2 = nullptr;
    void* partial = nullptr;                   // reduction partials (device)
    void* partial_host = nullptr;              // pinned host mirror
    int npartial = 0;
    int iold = 2, inew = 1, imid = -1;         // 1-based like the reference
    // properties in working precision, widened to double for storage
    double nu = 0, dt = 0, tau = 0, omega = 0, trt_magic = 0, csqr = 0;
    bool props_set = false;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int variant = 0;
    // deferred stepping (plbm_set_step_deferral): perform_lbm_step calls of fewer than defer_max steps are only
    // counted; they run as ONE batched call when defer_max is reached or any other entry point touches the grid
    int defer_max = 0, deferred_steps = 0, deferred_model = -1;
    int fdm_stencil = 0;  // stream_fdm_bardow derivative stencil: 0 default, 1 WLS, 2/3 WLS-Gauss v1/v2, 4 isotropic
    int sm_count = 148;
    Comm* comm = nullptr;
    // After a fused DUGKS step lattice `inew` still holds ftilde^n, whereas the reference leaves
    // fbar^{+,n} = BGK(ftilde^n, omega_half) there (what the lagged update_macros reads).  The
    // half-step collision is applied lazily, in place, the first time somebody looks at `inew`.
    bool dugks_pending = false;
    double dugks_pending_omega = 0;
    // TMA descriptors (CUtensorMap, 128 B each) of the lattices viewed as a (ny, nx, 9) tensor with
    // a (FY+2, FX+2, 9) box: used by the pipelined FVM/DUGKS tile kernel (plbm_fvm_tma.cu)
    alignas(64) unsigned char tmap[3][128];
    bool tmap_ok = false;
    // neighbours' boundary lines of all nine populations ([9][ld]) for the FVM/DUGKS tile kernel under a
    // slab decomposition; nullptr = periodic self-wrap
    const void* fv_halo_lo = nullptr;
    const void* fv_halo_hi = nullptr;
    // slab decomposition (single GPU: nx_global == nx, x_offset == 0)
    int nx_global = 0, x_offset = 0;
    // A third, hidden lattice buffer for the launch that CLOSES a perform_lbm_step call (lbm_spare, plbm_api.cu): lattice `inew`
    // must end up holding state n-1 and lattice `iold` state n, and with two buffers only a single-step launch can produce that
    // (47 GLUPS against 106 for the triples).  With a third buffer the closing launch is a triple that stores the states after
    // its second AND third step; the buffer it read becomes the spare.  Allocated on first use when the GPU has room for it.
    void* spare = nullptr;
    int spare_state = 0;  // 0 not tried yet, 1 allocated, -1 unavailable (no room, switched off, or the ring did not agree)

    size_t lattice_elems() const { return (size_t)ld * nx * 9; }
    size_t esize() const { return prec == PLBM_F64 ? 8 : 4; }
    template <typename T> T* lat(int which) const { return static_cast<T*>(f[which - 1]); }
    template <typename T> T* rho() const { return static_cast<T*>(mf); }
    template <typename T> T* ux() const { return static_cast<T*>(mf) + (size_t)nx * ny; }
    template <typename T> T* uy() const { return static_cast<T*>(mf) + 2 * (size_t)nx * ny; }
};

// Arguments of the fused pull-stream + collide kernel family (plbm_lbm.cu).
template <typename T> struct LbmArgs {
    const T* src;
    T* dst;
    int nx, ny, ld;        // lines in this slab, rows, leading dimension
    int x_begin, x_end;    // lines updated by this launch
    // two line ranges in one launch (both boundaries of a slab): lines at or beyond x_split are shifted by x_skip, i.e. the
    // launch covers [x_begin, x_split) and [x_split + x_skip, x_end); no split: x_split = INT_MAX, x_skip = 0
    int x_split = 0x7fffffff, x_skip = 0;
    // the ring neighbours' two nearest lines, all nine populations, [2][9][ld] each (lo: lines -2, -1;
    // hi: lines nx, nx+1); nullptr = periodic self-wrap by index arithmetic (single GPU)
    const T* halo_lo;
    const T* halo_hi;
    CollideParams<T> cp;
    T* pre = nullptr;      // optional: also store the streamed, pre-collision PDFs (perform_triple_step)
    T* mid = nullptr;      // slab schedule, closing triple: also store the state after its second step here (Grid::spare)
};

// ---- launchers (all asynchronous on `s`) -------------------------------------------------
template <typename T> int launch_lbm(const LbmArgs<T>& a, int model, bool stream_pdfs, int variant, cudaStream_t s);
template <typename T> int launch_init_eq(const Grid& g, T* f, cudaStream_t s);
template <typename T> int launch_macros(const Grid& g, const T* f, cudaStream_t s);
template <typename T>
int launch_vorticity(const Grid& g, int order, const T* ux, const T* uy, T* out, cudaStream_t s, const T* uy_lo = nullptr,
                     const T* uy_hi = nullptr);  // uy_lo / uy_hi: the ring neighbours' two nearest lines of uy, [2][ny]
template <typename T> int launch_diagnostics(Grid& g, double out[PLBM_DIAG_COUNT], cudaStream_t s);
template <typename T> int launch_l2_sums(Grid& g, const T* uxa, const T* uya, double out[2], cudaStream_t s);
template <typename T> int launch_lattice_hash(Grid& g, const T* f, unsigned long long* out, cudaStream_t s);
template <typename T>
int launch_fvm_bardow(const Grid& g, const T* fold, T* fnew, T dt, int model, const CollideParams<T>& cp, cudaStream_t s,
                      int mode = 2 /* 2 Bardow FVM, 4 Bardow FDM (Lax-Wendroff), 5 Sofonea FDM */);
template <typename T> int launch_dugks_collide(const Grid& g, T* fold, T* fnew, T omega_full, T omega_half, cudaStream_t s);
template <typename T> int launch_dugks_stream(const Grid& g, const T* ft, T* fp, T dt, T omega_face, bool dugks, cudaStream_t s);
template <typename T>
int launch_dugks_fused(const Grid& g, const T* fin, T* fout, T dt, T omega_full, T omega_half, T omega_face, bool dugks,
                       cudaStream_t s);
// cluster-resident multi-step kernel for grids that fit in the shared memory of one cluster (plbm_small.cu)
template <typename T>
int try_lbm_cluster_steps(const Grid& g, T* f_iold, T* f_inew, int model, const CollideParams<T>& cp, int nsteps, bool* done,
                          cudaStream_t s);
// two steps per pass over HBM through a shared-memory ring (plbm_lbm2.cu)
bool lbm_pair_applicable(const Grid& g);
// variants of perform_lbm_step that advance two steps per launch: 0 default, 5 never the cluster kernel,
// 6 per-thread loads forced, 7 bulk async copies forced, 8 = 7 issued as the slab schedule's three x ranges
// (9 / 10: the experimental depth-generic kernel of plbm_lbmn.cu on one GPU, pairs / triples; like 5 otherwise)
// (11: the default two-step kernels compiled with FMA contraction, plbm_lbm2_fma.cu, one GPU; like 5 otherwise)
// (12: like 5 with the OTHER fp32 collision code: scalar where packed FFMA2 pairs are the default, plbm_f32x2.cuh)
inline bool lbm_pair_variant(int variant) { return variant == 0 || (variant >= 5 && variant <= 12); }
int lbm_pair_flavour(const Grid& g);  // 0 one step per launch, 1 k_lbm2, 2 k_lbm2_bulk
template <typename T>
int launch_lbm_pair(const Grid& g, const T* src, T* dst, int x_begin, int x_end, const T* halo_lo, const T* halo_hi, int model,
                    const CollideParams<T>& cp, cudaStream_t s);
template <typename T>
int launch_lbm_pair_boundaries(const Grid& g, const T* src, T* dst, int nb, const T* halo_lo, const T* halo_hi, int model,
                               const CollideParams<T>& cp, cudaStream_t s);  // columns [0, nb) and [nx - nb, nx) in one launch
template <typename T>
int launch_lbm_pair_fma(const Grid& g, const T* src, T* dst, int x_begin, int x_end, const T* halo_lo, const T* halo_hi, int model,
                        const CollideParams<T>& cp, cudaStream_t s);  // plbm_lbm2_fma.cu: within tolerance, not bit-identical
// EXPERIMENTAL: `nstep` (2 or 3) fused steps per pass over HBM, depth-generic form of k_lbm2_bulk (plbm_lbmn.cu)
bool lbm_multi_applicable(const Grid& g, int model, int nstep);
template <typename T>
int launch_lbm_multi(const Grid& g, const T* src, T* dst, int x_begin, int x_end, int model, const CollideParams<T>& cp, int nstep,
                     cudaStream_t s, const T* halo_lo = nullptr, const T* halo_hi = nullptr,
                     T* dst_mid = nullptr);  // dst_mid (three steps, default shape): the state after step 2 as well
bool lbm_multi_shape_is_default();
bool lbm_triple_ws_wanted(int model);  // the triples that read no halo lines go to k_lbm3_ws (plbm_lbm3w.cu); PLBM_TRIPLE_WS overrides
// three steps per pass, warp-specialised and skewed (plbm_lbm3w.cu): no halo lines; dst_mid != nullptr: the state after step 2 too
template <typename T>
int launch_lbm_triple_ws(const Grid& g, const T* src, T* dst, T* dst_mid, int x_begin, int x_end, int model, const CollideParams<T>& cp,
                         cudaStream_t s);
// Halo message of the slab decomposition: the PLBM_HALO_LINES nearest lines of each neighbour, all nine populations,
// [PLBM_HALO_LINES][9][ld].  halo_lo holds lines -2, -1, -3 in that order (the first two are what one step / a fused pair read --
// k_lbm, k_lbm2 --, the third was appended for three steps per pass); halo_hi holds lines nx, nx+1, nx+2.
constexpr int PLBM_HALO_LINES = 3;
__host__ __device__ constexpr int halo_lo_index(int col) { return col == -3 ? 2 : col + 2; }          // col in [-3, -1]
__host__ __device__ constexpr int halo_lo_source_line(int l, int nx) { return l < 2 ? nx - 2 + l : nx - 3; }  // sender's line of slot l
// Three steps per pass (k_lbmn_bulk), see plbm_lbmn.cu.  Level of a grid / slab: -1 the kernel does not apply, 0 below 512^2
// nodes, 1 from 512^2 (a ring agrees on the minimum over its slabs: every rank must issue the same launches);
// lbm_triples_wanted: does this collision / variant take triples at that level.
int lbm_triples_level(const Grid& g);
// Steps the next launch of a call advances when `rem` steps remain BEFORE the closing single step (the last step of a call stays
// single so that lattice `inew` ends up holding state n-1 like the reference): triples, except that four remaining steps go as two
// pairs (a triple would leave one more single step: 3.67 + 2.82 ms against 2 x 3.1 ms on the bench slab), then pairs, then singles.
inline int lbm_next_depth(int rem, bool triples, bool pairs)
{
    if (triples && rem >= 3 && !(pairs && rem == 4)) return 3;
    return pairs && rem >= 2 ? 2 : 1;
}
bool lbm_triples_wanted(const Grid& g, int level, int model);
bool lbm_triples_forced();  // PLBM_TRIPLES=2
// The next launch of a call with `rem` steps left (the closing one included).  dual: a third lattice buffer is available, so the
// call may close with a triple that also stores the state after its second step (lattice `inew` = state n-1); otherwise the call
// closes with a single step as lbm_next_depth says.  With a closing dual triple: 3 -> it; 5 -> a pair first; 4 -> triple + single;
// 6 and more -> a triple.  Every rank of a ring computes the same sequence.
struct LbmLaunch {
    int depth;
    bool dual;
};
inline LbmLaunch lbm_next_launch(int rem, bool triples, bool pairs, bool dual)
{
    if (dual && triples) {
        if (rem == 3) return {3, true};
        if (rem == 5 && pairs) return {2, false};
        if (rem >= 4) return {3, false};
        return {1, false};
    }
    return {lbm_next_depth(rem - 1, triples, pairs), false};
}
// the hidden third lattice buffer of a grid (see Grid::spare): allocates it on first use; nullptr if unavailable
void* lbm_spare(Grid& g);
void lbm_adopt_spare_as_inew(Grid& g);  // after a dual triple + index swap: lattice `inew` <- spare (state n-1), spare <- the old source

bool comm_triples_level(const Grid& g, int* level);  // the ring's agreed level
// TMA + mbarrier pipelined tile kernel (plbm_fvm_tma.cu); `which` = 1-based source lattice
int make_tensor_maps(Grid& g);
template <typename T>
int launch_fv_tma(const Grid& g, int which_src, const T* fin, T* fout, int mode, int model, T dt, T omega_full, T omega_half,
                  T omega_face, const CollideParams<T>& cp, cudaStream_t s);
// the same kernels compiled with FMA contraction (plbm_fvm_tma_fma.cu): within tolerance, not bit-identical; variant 3
template <typename T>
int launch_fv_tma_fma(const Grid& g, int which_src, const T* fin, T* fout, int mode, int model, T dt, T omega_full, T omega_half,
                      T omega_face, const CollideParams<T>& cp, cudaStream_t s);
// OPT-IN (variant 4), tolerance-gated: marching DUGKS / Bardow-FVM kernel with shared cell faces (plbm_fvm_march.cu)
bool fv_march_applicable(int mode, int model);
template <typename T>
int launch_fv_march(const Grid& g, const T* fin, T* fout, int mode, int model, T dt, T omega_full, T omega_half, T omega_face,
                    const CollideParams<T>& cp, cudaStream_t s);
template <typename T> int launch_halo_pack(const Grid& g, const T* f, T* send_lo, T* send_hi, cudaStream_t s);

// multi-GPU ring exchange (plbm_comm.cu)
int comm_unique_id(void* id128);
int comm_init(Grid& g, const void* id128, int rank, int nranks, int nx_global, int x_offset);
int comm_finalize(Grid& g);
int comm_transport_is_p2p(const Grid& g);
bool comm_pairs_agreed(const Grid& g);  // every slab of the ring can run the two-step kernel
bool comm_dual_agreed(const Grid& g);   // every rank of the ring holds the third lattice buffer (closing dual triple)
void comm_invalidate_halo(Grid& g);  // the lattices were modified behind the ring's back
// ring-wide reduction of n doubles in place (every rank calls it; NaN propagates through max / min as through sum)
enum { PLBM_REDUCE_SUM = 0, PLBM_REDUCE_MAX = 2, PLBM_REDUCE_MIN = 3 };  // = ncclSum, ncclMax, ncclMin
int comm_allreduce(Grid& g, double* values, int n, int op);
template <typename T> int comm_lbm_steps(Grid& g, int model, const CollideParams<T>& cp, int nsteps);
// one FVM/DUGKS halo exchange: all nine populations of lines 0 and nx-1 of `f` -> g.fv_halo_lo/hi
template <typename T> int comm_fv_exchange(Grid& g, const T* f);
template <typename T> int launch_halo_pack9(const Grid& g, const T* f, T* send_lo, T* send_hi, cudaStream_t s);
// two lines of a macroscopic field ((nx, ny), unpadded) per direction: lines 0, 1 -> rank lo's *hi, lines nx-2, nx-1 -> rank hi's *lo
template <typename T> int comm_field_exchange2(Grid& g, const T* field, const T** lo, const T** hi);

}  // namespace plbm

struct plbm_grid_s : plbm::Grid {};
