// plbm_fvm_march.cu -- DUGKS / Bardow-FVM step as a MARCHING kernel with shared cell faces (sm_100a).
//
// OPT-IN (plbm_set_variant(h, 4)), TOLERANCE-GATED: not bit-identical to the reference, see "Arithmetic" below.
// The default, bit-identical kernel is k_fv_tma (plbm_fvm_tma.cu).
//
// Why another kernel.  k_fv_tma is bound by the fp64 pipe, not by HBM (ncu, round 1: 726 fp64 instructions per node,
// pipe 57 % busy, DRAM 135 B/node = 0.32 of the roofline): every node reconstructs and relaxes all FOUR of its faces, so
// every face of the grid is computed twice, and the half-step collision is redone on the halo ring of every 32 x 8 tile
// (340 / 256 nodes).  Here a block owns a strip of rows and marches along x (like k_lbm2): per column and node it
//   1. collides the raw node of column x + 1 (half-step -> fbar+, kept in a 3-column shared-memory ring; full-step -> the
//      node's own ftilde+, kept in registers for one iteration),                       src/periodic_dugks.F90:46-77
//   2. reconstructs and relaxes ONE vertical face, W(x + 1) -- which is the E face of column x and, one iteration later,
//      its own W face (registers) --, and ONE horizontal face, S(x, y), handing it to the node below through shared
//      memory as that node's N face,                                  src/periodic_dugks.F90:238-276, 310-434
//   3. applies the flux update to column x and stores it.                             src/periodic_dugks.F90:282-300
// Per node: one half-step + one full-step collision (sharing the moments), two face reconstructions, two face relaxations
// -- about 0.6 of the fp64 work of k_fv_tma -- and no recomputation along x (two halo rows per strip along y).
//
// Arithmetic.  The reference computes the east face of node x and the west face of node x + 1 separately:
//   cfe(x)   = p2*(fc+fe) - p2*cxq*(fe-fc) - p8*cyq*(((fne + fn ) - fse) - fs )
//   cfw(x+1) = p2*(fc+fw) - p2*cxq*(fc-fw) - p8*cyq*(((fnw + fn') - fsw) - fs')      (same four values, other order)
// The two agree except for the order of the last two subtractions, i.e. to one rounding of the cross term.  This kernel
// evaluates every face ONCE, in the cfw / cfs form, and uses it for both cells; it is also compiled with FMA contraction.
// Both change last bits only: tests/test_gpu_fast_variants.py gates the result at 1e-12 (fp64) / 1e-5 (fp32) relative to
// the reference arithmetic after N steps and over the golden Taylor-Green sweep (the tolerance the north-star states).
#include <cstdlib>

#include "plbm_internal.h"

namespace plbm {

namespace {

enum { MODE_DUGKS = 0, MODE_DUGKS_OFF = 1, MODE_BARDOW = 2 };

__device__ __forceinline__ int pmod(int i, int n)
{
    i %= n;
    return i < 0 ? i + n : i;
}

// compact index of the six populations with cy != 0 (2, 4, 5, 6, 7, 8) in the shared S-face buffer
__host__ __device__ constexpr int sidx(int q) { return q == 2 ? 0 : (q == 4 ? 1 : q - 3); }

template <typename T> struct MarchArgs {
    const T* fin;
    T* fout;
    int nx, ny, ld;
    int x_begin, x_end;
    int ty, nstrips, seglen;
    T dt, omega_full, omega_half, omega_face;
    CollideParams<T> cp;
    const T* hlo;  // ring neighbours' boundary lines, [9][ld] each (line -1 / line nx); nullptr = periodic self-wrap
    const T* hhi;
};

// row 0 of population 0 of column c (c in [-1, nx]) and the distance between populations
template <typename T> __device__ __forceinline__ const T* column_base(const MarchArgs<T>& a, int c, size_t& qstride)
{
    if (c < 0) {
        if (a.hlo) {
            qstride = (size_t)a.ld;
            return a.hlo;
        }
        c += a.nx;
    } else if (c >= a.nx) {
        if (a.hhi) {
            qstride = (size_t)a.ld;
            return a.hhi;
        }
        c -= a.nx;
    }
    qstride = (size_t)a.nx * a.ld;
    return a.fin + (size_t)c * a.ld;
}

// West face of column X at this thread's row: A = fbar column X - 1, B = fbar column X (pointers at [q = 0][row]; NT between
// populations).  ALL = every population (the moments of the relaxation need them), else the six with cx != 0.
template <typename T, int NT, bool ALL> __device__ __forceinline__ void west_face(const T* A, const T* B, T dt, T (&cf)[9])
{
    const T p2 = T(0.5);
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const int CX = cxi(q), CY = cyi(q);
        if (!ALL && CX == 0) continue;
        const T fc = B[q * NT], fw = A[q * NT];
        T v = p2 * (fc + fw);
        if (CX != 0) v = v - (p2 * (dt * T(CX))) * (fc - fw);
        if (CY != 0) {
            const T fnw = A[q * NT + 1], fn = B[q * NT + 1], fsw = A[q * NT - 1], fs = B[q * NT - 1];
            v = v - (T(0.125) * (dt * T(CY))) * (fnw + fn - fsw - fs);
        }
        cf[q] = v;
    }
}

// South face of node (X, row): L, C, R = fbar columns X - 1, X, X + 1.
template <typename T, int NT, bool ALL> __device__ __forceinline__ void south_face(const T* L, const T* C, const T* R, T dt, T (&cf)[9])
{
    const T p2 = T(0.5);
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const int CX = cxi(q), CY = cyi(q);
        if (!ALL && CY == 0) continue;
        const T fc = C[q * NT], fs = C[q * NT - 1];
        T v = p2 * (fc + fs);
        if (CY != 0) v = v - (p2 * (dt * T(CY))) * (fc - fs);
        if (CX != 0) {
            const T fse = R[q * NT - 1], fe = R[q * NT], fsw = L[q * NT - 1], fw = L[q * NT];
            v = v - (T(0.125) * (dt * T(CX))) * (fse + fe - fsw - fw);
        }
        cf[q] = v;
    }
}

template <typename T, int MODE, int MODEL, int NT, int MINB> __global__ void __launch_bounds__(NT, MINB) k_fv_march(const MarchArgs<T> a)
{
    constexpr bool HALF = MODE != MODE_BARDOW;  // faces from the half-step-collided state fbar+ (DUGKS), else from f^n
    constexpr bool RELAX = MODE == MODE_DUGKS;  // face relaxation (update_ew / update_ns)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* const ring = reinterpret_cast<T*>(smem_raw);  // [3 columns][9][NT]
    T* const sface = ring + 27 * NT;                 // [6][NT]: relaxed south faces of the column being updated

    const int strip = blockIdx.x % a.nstrips, seg = blockIdx.x / a.nstrips;
    const int y_lo = strip * a.ty;
    const int y_hi = min(y_lo + a.ty, a.ny);
    const int xs = a.x_begin + seg * a.seglen;
    const int xe = min(xs + a.seglen, a.x_end);
    const int t = threadIdx.x;
    const int yl = y_lo - 1 + t;                // logical row of this thread: y_lo - 1 .. y_lo + NT - 2
    const int yp = pmod(yl, a.ny);
    const bool act_f = yl <= y_hi;              // fbar: the strip and one halo row on each side
    const bool act_s = t >= 1 && yl <= y_hi;    // south faces: rows y_lo .. y_hi (the last one is the strip's top N face)
    const bool act_u = t >= 1 && yl < y_hi;     // updated nodes

    T raw[9];  // prefetched raw column
    auto fetch = [&](int c) {
        size_t qs;
        const T* base = column_base<T>(a, c, qs) + yp;
#pragma unroll
        for (int q = 0; q < 9; ++q) raw[q] = base[q * qs];
    };
    // raw column -> fbar into ring slot `slot`; returns the node's own post-collision state in fp
    auto stage1 = [&](int slot, T (&fp)[9]) {
        T b[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            b[q] = raw[q];
            fp[q] = raw[q];
        }
        if (HALF) {
            collide_bgk_split(b, a.omega_half);
            collide_bgk_split(fp, a.omega_full);
        }
        T* d = ring + slot * 9 * NT + t;
#pragma unroll
        for (int q = 0; q < 9; ++q) d[q * NT] = b[q];
    };

    int s_p1 = 0;  // ring slot of the newest column; the two older ones follow cyclically
    T fp_cur[9] = {}, fp_next[9] = {}, cw[9] = {}, ce[9] = {};

    // ---- warm-up: fbar of columns xs - 1 and xs, west face of column xs ----------------------------------------------
    if (act_f) {
        fetch(xs - 1);
        stage1(0, fp_next);  // fp of the halo column is not used
        fetch(xs);
        stage1(1, fp_cur);
        if (xs + 1 <= xe) fetch(xs + 1);  // column xe is needed too: it carries the east face of column xe - 1
    }
    s_p1 = 1;
    __syncthreads();
    if (act_u) {
        const T* A = ring + 0 * 9 * NT + t;
        const T* B = ring + 1 * 9 * NT + t;
        west_face<T, NT, RELAX>(A, B, a.dt, cw);
        if (RELAX) face_relax<T, true>(cw, a.omega_face);
    }

    for (int x = xs; x < xe; ++x) {
        // 1. column x + 1: fbar into the ring, own post-collision state into registers; prefetch column x + 2
        const int s_new = s_p1 == 2 ? 0 : s_p1 + 1;
        if (act_f) {
            stage1(s_new, fp_next);
            if (x + 2 <= xe) fetch(x + 2);
        }
        const int s_0 = s_p1, s_m1 = s_p1 == 0 ? 2 : s_p1 - 1;
        s_p1 = s_new;
        __syncthreads();
        const T* L = ring + s_m1 * 9 * NT + t;
        const T* C = ring + s_0 * 9 * NT + t;
        const T* R = ring + s_p1 * 9 * NT + t;
        // 2. east face of column x = west face of column x + 1
        if (act_u) {
            west_face<T, NT, RELAX>(C, R, a.dt, ce);
            if (RELAX) face_relax<T, true>(ce, a.omega_face);
        }
        // 3. south face of node (x, y); the node below reads it as its north face
        T cs[9] = {};
        if (act_s) {
            south_face<T, NT, RELAX>(L, C, R, a.dt, cs);
            if (RELAX) face_relax<T, false>(cs, a.omega_face);
#pragma unroll
            for (int q = 1; q < 9; ++q)
                if (cyi(q) != 0) sface[sidx(q) * NT + t] = cs[q];
        }
        __syncthreads();
        // 4. flux update of column x (src/periodic_dugks.F90:297), collision for the Bardow scheme, store
        if (act_u) {
#pragma unroll
            for (int q = 1; q < 9; ++q)
                if (cxi(q) != 0) fp_cur[q] = fp_cur[q] - (a.dt * T(cxi(q))) * (ce[q] - cw[q]);
#pragma unroll
            for (int q = 1; q < 9; ++q)
                if (cyi(q) != 0) fp_cur[q] = fp_cur[q] - (a.dt * T(cyi(q))) * (sface[sidx(q) * NT + t + 1] - cs[q]);
            if (MODE == MODE_BARDOW && MODEL != M_NONE) collide<T, MODEL>(fp_cur, a.cp);
            T* out = a.fout + (size_t)x * a.ld + yp;
            const size_t qs = (size_t)a.nx * a.ld;
#pragma unroll
            for (int q = 0; q < 9; ++q) out[q * qs] = fp_cur[q];
        }
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            cw[q] = ce[q];
            fp_cur[q] = fp_next[q];
        }
    }
}

// ---- k_fv_march_s: the same step with the face stencils written as sums -----------------------------------------------
// k_fv_march reads every operand of a face from the shared-memory ring: 42 loads per face, 105 shared-memory operations per node
// and column, and ncu (r02a) shows the shared-memory pipe 60 % busy next to the fp64 pipe at 50 % -- the two do not overlap.  Here
// a thread keeps its OWN row of the last two fbar columns in registers and the neighbours' contributions are exchanged as sums:
//   west face of column X:   fc + fw = SW(t),  cross term = SW(t+1) - SW(t-1),   SW(t) = fbar[X-1][t] + fbar[X][t]
//   south face of node (X,t): cross term = DS(t) + DS(t-1),                      DS(t) = fbar[X+1][t] - fbar[X-1][t]
// so that per node and column 27 values are stored and 33 loaded (60 operations instead of 105).  The sums re-associate
// the reference's cross terms (last-bit differences, like the shared faces themselves): same tolerance gate as k_fv_march.
__host__ __device__ constexpr int xidx(int q) { return q == 1 ? 0 : (q == 3 ? 1 : q - 3); }  // populations with cx != 0

template <typename T, int MODE, int MODEL, int NT, int MINB> __global__ void __launch_bounds__(NT, MINB) k_fv_march_s(const MarchArgs<T> a)
{
    constexpr bool HALF = MODE != MODE_BARDOW;
    constexpr bool RELAX = MODE == MODE_DUGKS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* const ring = reinterpret_cast<T*>(smem_raw);  // [2 columns][9][NT]: fbar, for the row below (fs of the south face)
    T* const sw = ring + 18 * NT;                    // [6][NT] SW of the populations with cy != 0
    T* const ds = sw + 6 * NT;                       // [6][NT] DS of the populations with cx != 0
    T* const sface = ds + 6 * NT;                    // [6][NT] relaxed south faces of the column being updated

    const int strip = blockIdx.x % a.nstrips, seg = blockIdx.x / a.nstrips;
    const int y_lo = strip * a.ty;
    const int y_hi = min(y_lo + a.ty, a.ny);
    const int xs = a.x_begin + seg * a.seglen;
    const int xe = min(xs + a.seglen, a.x_end);
    const int t = threadIdx.x;
    const int yl = y_lo - 1 + t;
    const int yp = pmod(yl, a.ny);
    const bool act_f = yl <= y_hi;
    const bool act_s = t >= 1 && yl <= y_hi;
    const bool act_u = t >= 1 && yl < y_hi;

    T raw[9];
    auto fetch = [&](int c) {
        size_t qs;
        const T* base = column_base<T>(a, c, qs) + yp;
#pragma unroll
        for (int q = 0; q < 9; ++q) raw[q] = base[q * qs];
    };
    // raw column -> fbar (b) and the node's own post-collision state (fp)
    auto stage1 = [&](T (&b)[9], T (&fp)[9]) {
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            b[q] = raw[q];
            fp[q] = raw[q];
        }
        if (HALF) {
            collide_bgk_split(b, a.omega_half);
            collide_bgk_split(fp, a.omega_full);
        }
    };
    // west face of the column whose fbar is bB, from the column before it (bA) and the SW sums of the rows above and below
    auto west = [&](const T (&bA)[9], const T (&bB)[9], const T (&swv)[9], T (&cf)[9]) {
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            const int CX = cxi(q), CY = cyi(q);
            if (!RELAX && CX == 0) continue;
            T v = T(0.5) * swv[q];
            if (CX != 0) v = v - (T(0.5) * (a.dt * T(CX))) * (bB[q] - bA[q]);
            if (CY != 0) v = v - (T(0.125) * (a.dt * T(CY))) * (sw[sidx(q) * NT + t + 1] - sw[sidx(q) * NT + t - 1]);
            cf[q] = v;
        }
        if (RELAX) face_relax<T, true>(cf, a.omega_face);
    };

    T bL[9] = {}, bC[9] = {}, b[9] = {}, swv[9] = {}, dsv[9] = {};
    T fp_cur[9] = {}, fp_next[9] = {}, cw[9] = {}, ce[9] = {};
    int slot_c = 0;  // ring slot of column x

    // ---- warm-up: fbar of columns xs - 1 (bL) and xs (bC), west face of column xs --------------------------------------
    if (act_f) {
        fetch(xs - 1);
        stage1(bL, fp_next);  // fp of the halo column is not used
        fetch(xs);
        stage1(bC, fp_cur);
        if (xs + 1 <= xe) fetch(xs + 1);
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            swv[q] = bL[q] + bC[q];
            if (cyi(q) != 0) sw[sidx(q) * NT + t] = swv[q];
            ring[(slot_c * 9 + q) * NT + t] = bC[q];
        }
    }
    __syncthreads();
    if (act_u) west(bL, bC, swv, cw);
    __syncthreads();  // the loop overwrites sw

    for (int x = xs; x < xe; ++x) {
        // 1. column x + 1: fbar and own post-collision state; what the neighbours need, as sums; prefetch column x + 2
        if (act_f) {
            stage1(b, fp_next);
            if (x + 2 <= xe) fetch(x + 2);
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                ring[((slot_c ^ 1) * 9 + q) * NT + t] = b[q];
                swv[q] = bC[q] + b[q];
                if (cyi(q) != 0) sw[sidx(q) * NT + t] = swv[q];
                if (cxi(q) != 0) {
                    dsv[q] = b[q] - bL[q];
                    ds[xidx(q) * NT + t] = dsv[q];
                }
            }
        }
        __syncthreads();
        // 2. east face of column x = west face of column x + 1
        if (act_u) west(bC, b, swv, ce);
        // 3. south face of node (x, y); the node below reads it as its north face
        T cs[9] = {};
        if (act_s) {
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                const int CX = cxi(q), CY = cyi(q);
                if (!RELAX && CY == 0) continue;
                const T fc = bC[q], fs = ring[(slot_c * 9 + q) * NT + t - 1];
                T v = T(0.5) * (fc + fs);
                if (CY != 0) v = v - (T(0.5) * (a.dt * T(CY))) * (fc - fs);
                if (CX != 0) v = v - (T(0.125) * (a.dt * T(CX))) * (dsv[q] + ds[xidx(q) * NT + t - 1]);
                cs[q] = v;
            }
            if (RELAX) face_relax<T, false>(cs, a.omega_face);
#pragma unroll
            for (int q = 1; q < 9; ++q)
                if (cyi(q) != 0) sface[sidx(q) * NT + t] = cs[q];
        }
        __syncthreads();
        // 4. flux update of column x (src/periodic_dugks.F90:297), collision for the Bardow scheme, store
        if (act_u) {
#pragma unroll
            for (int q = 1; q < 9; ++q)
                if (cxi(q) != 0) fp_cur[q] = fp_cur[q] - (a.dt * T(cxi(q))) * (ce[q] - cw[q]);
#pragma unroll
            for (int q = 1; q < 9; ++q)
                if (cyi(q) != 0) fp_cur[q] = fp_cur[q] - (a.dt * T(cyi(q))) * (sface[sidx(q) * NT + t + 1] - cs[q]);
            if (MODE == MODE_BARDOW && MODEL != M_NONE) collide<T, MODEL>(fp_cur, a.cp);
            T* out = a.fout + (size_t)x * a.ld + yp;
            const size_t qs = (size_t)a.nx * a.ld;
#pragma unroll
            for (int q = 0; q < 9; ++q) out[q * qs] = fp_cur[q];
        }
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            cw[q] = ce[q];
            fp_cur[q] = fp_next[q];
            bL[q] = bC[q];
            bC[q] = b[q];
        }
        slot_c ^= 1;
    }
}

int env_int(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

template <typename T, int MODE, int MODEL, int NT, int MINB, bool SUMS = false>
int launch_march(const Grid& g, const T* fin, T* fout, T dt, T of, T oh, T oc, const CollideParams<T>& cp, cudaStream_t s)
{
    constexpr size_t smem = (size_t)(SUMS ? 18 + 18 : 27 + 6) * NT * sizeof(T);
    void (*kern)(const MarchArgs<T>);
    if constexpr (SUMS) kern = k_fv_march_s<T, MODE, MODEL, NT, MINB>;
    else kern = k_fv_march<T, MODE, MODEL, NT, MINB>;
    static bool configured[64] = {false};
    if (g.device < 64 && !configured[g.device]) {
        PLBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // Carve-out left to the driver (it sizes it for three 128-thread blocks; the rest stays L1 for the raw-column loads).
        // Measured, DUGKS fp64 2048^2 (r02a / r02d): default carve-out 24.4 GLUPS; carve-out 100 % 22.5, and with it 4 x 128, 6 or 8 x 64,
        // 12 or 16 x 32 threads per SM all land on 22.3 - 23.6: the kernel is not bound by occupancy or by barrier coupling.
        if (env_int("PLBM_MARCH_CARVEOUT", 0) > 0)
            PLBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, env_int("PLBM_MARCH_CARVEOUT", 0)));
        configured[g.device] = true;
    }
    MarchArgs<T> a;
    a.fin = fin;
    a.fout = fout;
    a.nx = g.nx;
    a.ny = g.ny;
    a.ld = g.ld;
    a.x_begin = 0;
    a.x_end = g.nx;
    a.dt = dt;
    a.omega_full = of;
    a.omega_half = oh;
    a.omega_face = oc;
    a.cp = cp;
    a.hlo = static_cast<const T*>(g.fv_halo_lo);
    a.hhi = static_cast<const T*>(g.fv_halo_hi);
    // strips of at most NT - 2 rows; segments so that the blocks fill whole rounds of MINB blocks per SM (see the launcher
    // of k_lbm2_bulk), at least 8 columns each (two warm-up columns per segment)
    const int ty_max = NT - 2;
    const int slots = MINB * g.sm_count;
    a.nstrips = (g.ny + ty_max - 1) / ty_max;
    a.ty = (g.ny + a.nstrips - 1) / a.nstrips;
    a.nstrips = (g.ny + a.ty - 1) / a.ty;
    const int ncols = g.nx;
    int nseg = env_int("PLBM_MARCH_NSEG", 0);
    if (nseg <= 0) {
        const long long blocks64 = (long long)a.nstrips * ((ncols + 63) / 64);
        const long long rounds = blocks64 >= slots ? (blocks64 + slots - 1) / slots : 1;
        nseg = (int)(rounds * slots / a.nstrips);
    }
    if (nseg < 1) nseg = 1;
    a.seglen = (ncols + nseg - 1) / nseg;
    if (a.seglen < 8) a.seglen = ncols < 8 ? ncols : 8;
    nseg = (ncols + a.seglen - 1) / a.seglen;
    kern<<<(unsigned)(a.nstrips * nseg), NT, smem, s>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PLBM_CUDA(cudaGetLastError());
    return PLBM_OK;
}

template <typename T, int MODE, int MODEL>
int launch_march_shape(const Grid& g, const T* fin, T* fout, T dt, T of, T oh, T oc, const CollideParams<T>& cp, cudaStream_t s)
{
    // block shape: measurement knobs PLBM_MARCH_NT (128 / 256) and PLBM_MARCH_MINB
    static const int nt = env_int("PLBM_MARCH_NT", 128);
    static const int minb = env_int("PLBM_MARCH_MINB", sizeof(T) == 8 ? 3 : 4);
    // PLBM_MARCH_FORM: 0 = k_fv_march (operands from the ring), 1 = k_fv_march_s (faces from sums, own rows in registers; 128 threads
    // x 2 / 3 blocks, 64 threads x 5 blocks).  Measured at 2048^2 (r02e, GLUPS, form 0 -> form 1 with two blocks per SM): DUGKS fp32
    // 36.1 -> 40.3, Bardow fp32 31.6 -> 38.8, DUGKS fp64 24.3 -> 23.9 (226 registers: 18.2 when capped at 168 for three blocks),
    // Bardow fp64 23.5 -> 19.9.  Default: the sum form in fp32, the ring form in fp64.
    static const int form = env_int("PLBM_MARCH_FORM", sizeof(T) == 4 ? 1 : 0);
    if (form == 1) {
        static const int minb_s = env_int("PLBM_MARCH_MINB", 2);
        if (nt == 64) return launch_march<T, MODE, MODEL, 64, 5, true>(g, fin, fout, dt, of, oh, oc, cp, s);
        if (minb_s >= 3) return launch_march<T, MODE, MODEL, 128, 3, true>(g, fin, fout, dt, of, oh, oc, cp, s);
        return launch_march<T, MODE, MODEL, 128, 2, true>(g, fin, fout, dt, of, oh, oc, cp, s);
    }
    if (nt == 256) {
        if (minb >= 2) return launch_march<T, MODE, MODEL, 256, 2>(g, fin, fout, dt, of, oh, oc, cp, s);
        return launch_march<T, MODE, MODEL, 256, 1>(g, fin, fout, dt, of, oh, oc, cp, s);
    }
    // small blocks: the two barriers per column couple fewer warps, so the load and the arithmetic phases of different
    // blocks overlap on an SM
    if (nt == 64) {
        if (minb >= 8) return launch_march<T, MODE, MODEL, 64, 8>(g, fin, fout, dt, of, oh, oc, cp, s);
        return launch_march<T, MODE, MODEL, 64, 6>(g, fin, fout, dt, of, oh, oc, cp, s);
    }
    if (nt == 32) {
        if (minb >= 16) return launch_march<T, MODE, MODEL, 32, 16>(g, fin, fout, dt, of, oh, oc, cp, s);
        return launch_march<T, MODE, MODEL, 32, 12>(g, fin, fout, dt, of, oh, oc, cp, s);
    }
    if (minb >= 4) return launch_march<T, MODE, MODEL, 128, 4>(g, fin, fout, dt, of, oh, oc, cp, s);
    if (minb == 3) return launch_march<T, MODE, MODEL, 128, 3>(g, fin, fout, dt, of, oh, oc, cp, s);
    return launch_march<T, MODE, MODEL, 128, 2>(g, fin, fout, dt, of, oh, oc, cp, s);
}

}  // namespace

// mode: 0 DUGKS (-DDUGKS), 1 periodic_dugks without the macro, 2 Bardow FVM + collision `model` (bgk / trt / rr / none)
bool fv_march_applicable(int mode, int model)
{
    if (mode == MODE_DUGKS || mode == MODE_DUGKS_OFF) return true;
    return mode == MODE_BARDOW && (model == M_NONE || model == M_BGK || model == M_TRT || model == M_RR);
}

template <typename T>
int launch_fv_march(const Grid& g, const T* fin, T* fout, int mode, int model, T dt, T of, T oh, T oc, const CollideParams<T>& cp,
                    cudaStream_t s)
{
    if (mode == MODE_DUGKS) return launch_march_shape<T, MODE_DUGKS, M_NONE>(g, fin, fout, dt, of, oh, oc, cp, s);
    if (mode == MODE_DUGKS_OFF) return launch_march_shape<T, MODE_DUGKS_OFF, M_NONE>(g, fin, fout, dt, of, oh, oc, cp, s);
    switch (model) {
    case M_NONE: return launch_march_shape<T, MODE_BARDOW, M_NONE>(g, fin, fout, dt, of, oh, oc, cp, s);
    case M_BGK: return launch_march_shape<T, MODE_BARDOW, M_BGK>(g, fin, fout, dt, of, oh, oc, cp, s);
    case M_TRT: return launch_march_shape<T, MODE_BARDOW, M_TRT>(g, fin, fout, dt, of, oh, oc, cp, s);
    case M_RR: return launch_march_shape<T, MODE_BARDOW, M_RR>(g, fin, fout, dt, of, oh, oc, cp, s);
    }
    set_error("fv_march: collision not instantiated for the marching kernel");
    return PLBM_ERR_ARG;
}

template int launch_fv_march<double>(const Grid&, const double*, double*, int, int, double, double, double, double,
                                     const CollideParams<double>&, cudaStream_t);
template int launch_fv_march<float>(const Grid&, const float*, float*, int, int, float, float, float, float, const CollideParams<float>&,
                                    cudaStream_t);

}  // namespace plbm
