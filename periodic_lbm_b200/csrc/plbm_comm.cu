// plbm_comm.cu -- slab decomposition along the slow index across the GPUs of one box.
//
// The reference is single-process; this is new functionality whose correctness
// is defined as "bitwise identical to the single-GPU result".  Rank r owns nx lines starting
// at x_offset and forms a periodic ring with r-1 (lo) and r+1 (hi).  Pull streaming needs,
// per step, the last line of the populations moving in +x (q = 1,5,8) from lo and the first
// line of those moving in -x (q = 3,6,7) from hi: 3 x ld reals per direction.
//
// Per step (src -> dst), two streams:
//   main : wait(halo[src]) -> boundary lines 0 and nx-1 -> pack dst boundary -> record(packed)
//          -> interior lines 1 .. nx-2                       (overlaps the exchange)
//   comm : wait(packed) -> grouped ncclSend x2 / ncclRecv x2 into halo[dst] -> record(halo[dst])
// Halo buffers are double-buffered by step parity.  NCCL is resolved with dlopen at
// comm_init time so the single-GPU library has no link-time dependency on it.
//
// Two transports for the LBM halo (same schedule, same result):
//   p2p  (default): the neighbours' halo buffers are mapped into this process with CUDA IPC and a small
//         push kernel stores the three outgoing populations of each boundary line STRAIGHT into the
//         neighbour's halo slot over NVLink, then raises an epoch flag there; the consumer's stream
//         blocks on that flag with cuStreamWaitValue32 (no spinning kernel, no send/recv kernels, no
//         staging copies).  NCCL is only used to bootstrap (IPC handles, agreement, quiesce barrier).
//   nccl (PLBM_HALO=nccl, or when IPC is unavailable): pack kernel + grouped ncclSend/ncclRecv.
#include <cuda.h>
#include <dlfcn.h>

#include <cstdlib>

#include <cstring>

#include <utility>

#include "plbm_internal.h"

namespace plbm {

namespace {
struct NcclUniqueId {
    char internal[128];
};
typedef void* ncclComm_t;
typedef int ncclResult_t;
constexpr int kNcclInt8 = 0, kNcclInt32 = 2, kNcclFloat64 = 8, kNcclMin = 3;

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(NcclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl()
{
    if (g_nccl.lib) return PLBM_OK;
    // a process that already imported torch has its bundled libnccl.so.2 mapped: same soname
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names)
        if ((lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!lib) {
        set_error(std::string("cannot load NCCL: ") + dlerror());
        return PLBM_ERR_COMM;
    }
#define SYM(field, name)                                                   \
    *(void**)(&g_nccl.field) = dlsym(lib, name);                           \
    if (!g_nccl.field) {                                                   \
        set_error(std::string("NCCL symbol missing: ") + name);            \
        return PLBM_ERR_COMM;                                              \
    }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(AllReduce, "ncclAllReduce")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.lib = lib;
    return PLBM_OK;
}

int nccl_fail(ncclResult_t r, const char* what)
{
    set_error(std::string("NCCL error: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?") + " in " + what);
    return PLBM_ERR_COMM;
}
#define PLBM_NCCL(call)                                  \
    do {                                                 \
        ncclResult_t _r = (call);                        \
        if (_r != 0) return nccl_fail(_r, #call);        \
    } while (0)
}  // namespace

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1, lo = 0, hi = 0;
    cudaStream_t stream = nullptr;          // exchange stream
    cudaEvent_t ev_packed = nullptr;        // send buffers ready (main -> comm)
    cudaEvent_t ev_halo[2] = {nullptr, nullptr};  // halo[p] received (comm -> main)
    cudaEvent_t ev_consumed[2] = {nullptr, nullptr};  // halo[p] read by the boundary kernels (main -> comm)
    void* send_lo = nullptr;                // lines 0, 1       -> rank lo   ([2][9][ld], NCCL transport)
    void* send_hi = nullptr;                // lines nx-2, nx-1 -> rank hi
    void* halo_lo[2] = {nullptr, nullptr};  // from lo: its lines nx-2, nx-1 (our lines -2, -1)
    void* halo_hi[2] = {nullptr, nullptr};  // from hi: its lines 0, 1      (our lines nx, nx+1)
    // FVM / DUGKS: all nine populations of one line per direction (not overlapped: those kernels are
    // ALU-bound and the 9*ld-real message is microseconds)
    void* send9_lo = nullptr;
    void* send9_hi = nullptr;
    void* halo9_lo = nullptr;
    void* halo9_hi = nullptr;
    cudaEvent_t ev9_packed = nullptr, ev9_done = nullptr;
    // p2p transport: one IPC-exported block [halo_lo0 | halo_lo1 | halo_hi0 | halo_hi1 | flags]
    bool p2p = false;
    unsigned char* ipc_block = nullptr;
    unsigned char* peer_lo = nullptr;   // rank lo's block, mapped here (== ipc_block when nranks == 1)
    unsigned char* peer_hi = nullptr;
    bool opened_lo = false, opened_hi = false;
    size_t slot_bytes = 0;              // halo slot stride inside the block (256-B multiple)
    unsigned* ticket = nullptr;         // completion counter of the push kernel (local)
    unsigned epoch = 0;                 // number of exchanges issued; exchange e uses slot e & 1, flag value e
    cudaEvent_t ev_boundary = nullptr;  // boundary lines of dst written and pushed (boundary stream -> main)
    cudaEvent_t ev_interior = nullptr;  // interior lines of dst written (main -> boundary stream)
    cudaStream_t bstream = nullptr;     // high-priority stream of the boundary launches + halo push (p2p transport)
    size_t bytes = 0;   // one LBM halo message
    size_t bytes9 = 0;  // one FVM / DUGKS halo message
    int parity = 0;        // halo slot holding the neighbours' lines of lattice `iold`
    bool halo_valid = false;
    bool pairs_ok = false;  // every slab of the ring can run the two-step kernel (agreed at init)
    int triples_level = -1;  // minimum over the slabs of lbm_triples_level (agreed at init): three steps per pass
    bool dual_ok = false;    // every rank holds a third lattice buffer (agreed at init): a call may close with a dual triple
    int halo_of_lattice = 0;
};


namespace {
typedef CUresult (*WaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
WaitValue32Fn wait_value32()
{
    static WaitValue32Fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<WaitValue32Fn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

// flags live behind the four halo slots: u32 flag_lo[2], flag_hi[2]
__host__ __device__ inline size_t off_lo(size_t slot_bytes, int p) { return (size_t)p * slot_bytes; }
__host__ __device__ inline size_t off_hi(size_t slot_bytes, int p) { return (size_t)(2 + p) * slot_bytes; }
__host__ __device__ inline size_t off_flag_lo(size_t slot_bytes, int p) { return 4 * slot_bytes + 4 * (size_t)p; }
__host__ __device__ inline size_t off_flag_hi(size_t slot_bytes, int p) { return 4 * slot_bytes + 8 + 4 * (size_t)p; }

// Store the three boundary lines of each side of `f` (all nine populations, [PLBM_HALO_LINES][9][ld]) into the neighbours'
// halo slots over NVLink (peer pointers), then -- last block to finish -- publish the epoch in their flags.
//   lines 0, 1, 2 -> rank lo's halo_hi[slot]      lines nx-2, nx-1, nx-3 -> rank hi's halo_lo[slot]
// One step consumes the nearest line only, a fused pair of steps (plbm_lbm2.cu) two, a fused triple (plbm_lbmn.cu) all three.
template <typename T>
__global__ void __launch_bounds__(256)
    k_halo_push(const T* __restrict__ f, T* __restrict__ peer_lo_halo_hi, T* __restrict__ peer_hi_halo_lo, int nx, int ld,
                unsigned* peer_lo_flag_hi, unsigned* peer_hi_flag_lo, unsigned epoch, unsigned* ticket)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < PLBM_HALO_LINES * 9 * ld) {
        const int lq = i / ld, y = i - lq * ld;
        const int l = lq / 9, q = lq - 9 * l;
        const int line_lo = min(l, nx - 1), line_hi = max(halo_lo_source_line(l, nx), 0);  // thin slabs repeat a line (unused)
        peer_lo_halo_hi[i] = f[((size_t)q * nx + line_lo) * (size_t)ld + y];
        peer_hi_halo_lo[i] = f[((size_t)q * nx + line_hi) * (size_t)ld + y];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) {  // every block's stores are fenced before its ticket
            *ticket = 0;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned*>(peer_lo_flag_hi) = epoch;
            *reinterpret_cast<volatile unsigned*>(peer_hi_flag_lo) = epoch;
            __threadfence_system();
        }
    }
}
}  // namespace

// Neighbour barrier: both streams drained, then a 4-byte ring handshake.  Needed before an exchange that
// is not preceded by a consuming step (initial condition / upload), see comm_lbm_steps.
static int quiesce(Grid& g)
{
    Comm* c = g.comm;
    PLBM_CUDA(cudaStreamSynchronize(g.stream));
    PLBM_CUDA(cudaStreamSynchronize(c->stream));
    PLBM_NCCL(g_nccl.GroupStart());
    PLBM_NCCL(g_nccl.Send(c->send_lo, 4, kNcclInt8, c->lo, c->comm, c->stream));
    PLBM_NCCL(g_nccl.Send(c->send_hi, 4, kNcclInt8, c->hi, c->comm, c->stream));
    PLBM_NCCL(g_nccl.Recv((char*)c->send9_lo, 4, kNcclInt8, c->hi, c->comm, c->stream));
    PLBM_NCCL(g_nccl.Recv((char*)c->send9_hi, 4, kNcclInt8, c->lo, c->comm, c->stream));
    PLBM_NCCL(g_nccl.GroupEnd());
    PLBM_CUDA(cudaStreamSynchronize(c->stream));
    return PLBM_OK;
}

// Ring-wide minimum of a flag (every rank must call it).
static int agree_min(Grid& g, int* value)
{
    Comm* c = g.comm;
    int* flag = (int*)c->send9_hi;  // scratch device memory
    PLBM_CUDA(cudaMemcpy(flag, value, sizeof(int), cudaMemcpyHostToDevice));
    PLBM_NCCL(g_nccl.AllReduce(flag, flag, 1, kNcclInt32, kNcclMin, c->comm, c->stream));
    PLBM_CUDA(cudaStreamSynchronize(c->stream));
    PLBM_CUDA(cudaMemcpy(value, flag, sizeof(int), cudaMemcpyDeviceToHost));
    return PLBM_OK;
}

// Try to set up the p2p transport; on any failure (on ANY rank: agreed by an all-reduce) stay on NCCL.
static int p2p_setup(Grid& g)
{
    Comm* c = g.comm;
    const char* env = getenv("PLBM_HALO");
    int ok = !(env && env[0] == 'n') && wait_value32() != nullptr;
    c->slot_bytes = (c->bytes + 255) / 256 * 256;
    const size_t block_bytes = 4 * c->slot_bytes + 256;
    cudaIpcMemHandle_t mine, from_lo, from_hi;
    memset(&mine, 0, sizeof(mine));
    if (ok && cudaMalloc(&c->ipc_block, block_bytes) != cudaSuccess) ok = 0;
    if (ok && cudaMemset(c->ipc_block, 0, block_bytes) != cudaSuccess) ok = 0;
    if (ok && cudaMalloc(&c->ticket, 256) != cudaSuccess) ok = 0;
    if (ok && cudaMemset(c->ticket, 0, 256) != cudaSuccess) ok = 0;
    if (ok && cudaIpcGetMemHandle(&mine, c->ipc_block) != cudaSuccess) ok = 0;
    cudaGetLastError();
    // every rank takes part in the handle exchange, whatever its own state (keeps NCCL calls matched)
    char* stage = (char*)c->send9_lo;  // scratch device memory: [mine | from_lo | from_hi]
    PLBM_CUDA(cudaMemcpy(stage, &mine, sizeof(mine), cudaMemcpyHostToDevice));
    PLBM_NCCL(g_nccl.GroupStart());
    PLBM_NCCL(g_nccl.Send(stage, sizeof(mine), kNcclInt8, c->lo, c->comm, c->stream));
    PLBM_NCCL(g_nccl.Send(stage, sizeof(mine), kNcclInt8, c->hi, c->comm, c->stream));
    PLBM_NCCL(g_nccl.Recv(stage + 128, sizeof(mine), kNcclInt8, c->hi, c->comm, c->stream));
    PLBM_NCCL(g_nccl.Recv(stage + 256, sizeof(mine), kNcclInt8, c->lo, c->comm, c->stream));
    PLBM_NCCL(g_nccl.GroupEnd());
    PLBM_CUDA(cudaStreamSynchronize(c->stream));
    PLBM_CUDA(cudaMemcpy(&from_hi, stage + 128, sizeof(mine), cudaMemcpyDeviceToHost));
    PLBM_CUDA(cudaMemcpy(&from_lo, stage + 256, sizeof(mine), cudaMemcpyDeviceToHost));
    if (ok) {
        if (c->nranks == 1) {
            c->peer_lo = c->peer_hi = c->ipc_block;
        } else {
            if (cudaIpcOpenMemHandle((void**)&c->peer_lo, from_lo, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess)
                c->opened_lo = true;
            else
                ok = 0;
            if (ok && c->lo == c->hi) {
                c->peer_hi = c->peer_lo;  // two ranks: the same neighbour on both sides
            } else if (ok) {
                if (cudaIpcOpenMemHandle((void**)&c->peer_hi, from_hi, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess)
                    c->opened_hi = true;
                else
                    ok = 0;
            }
            cudaGetLastError();
        }
    }
    // agreement: p2p only if it works everywhere
    int rc = agree_min(g, &ok);
    if (rc) return rc;
    c->p2p = ok != 0;
    // fused pairs of steps only if every slab of the ring can run them: the ranks must issue the same
    // sequence of launches (one halo message per launch)
    int pairs = lbm_pair_applicable(g) ? 1 : 0;
    if ((rc = agree_min(g, &pairs))) return rc;
    c->pairs_ok = pairs != 0;
    int level = lbm_triples_level(g) + 1;  // agree_min works on non-negative flags
    if ((rc = agree_min(g, &level))) return rc;
    c->triples_level = level - 1;
    // a call may close with a triple that stores the states after its second and third step (Grid::spare) only if every rank has
    // the third buffer: the ranks must issue the same launches
    const bool triples_here = c->triples_level >= 1 || (c->triples_level == 0 && lbm_triples_forced());  // what lbm_triples_wanted says
    int dual = c->p2p && triples_here && lbm_multi_shape_is_default() && lbm_spare(g) ? 1 : 0;
    if ((rc = agree_min(g, &dual))) return rc;
    c->dual_ok = dual != 0;
    if (!c->dual_ok && g.spare) {
        cudaFree(g.spare);
        g.spare = nullptr;
        g.spare_state = -1;
    }
    if (c->p2p) {
        for (int p = 0; p < 2; ++p) {  // the halo slots now live inside the exported block
            cudaFree(c->halo_lo[p]);
            cudaFree(c->halo_hi[p]);
            c->halo_lo[p] = c->ipc_block + off_lo(c->slot_bytes, p);
            c->halo_hi[p] = c->ipc_block + off_hi(c->slot_bytes, p);
        }
    }
    return PLBM_OK;
}

// Ring-wide reduction of a few doubles (global diagnostics).  NCCL's max / min follow the IEEE maxNum rule and
// would drop a NaN, so a NaN flag is reduced alongside and re-applied: a rank that diverged poisons the global value.
int comm_allreduce(Grid& g, double* values, int n, int op)
{
    Comm* c = g.comm;
    if (!c || n < 1 || n > 8) {
        set_error("comm_allreduce: no ring or bad count");
        return PLBM_ERR_ARG;
    }
    double stage[9];
    double nan_flag = 0.0;
    for (int i = 0; i < n; ++i) {
        stage[i] = values[i];
        if (values[i] != values[i]) {
            nan_flag = 1.0;
            stage[i] = 0.0;
        }
    }
    double* dev = (double*)c->send9_hi;  // scratch device memory (>= 9 * ld reals)
    PLBM_CUDA(cudaStreamSynchronize(g.stream));
    PLBM_CUDA(cudaMemcpyAsync(dev, stage, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    PLBM_CUDA(cudaMemcpyAsync(dev + 8, &nan_flag, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    PLBM_NCCL(g_nccl.GroupStart());
    PLBM_NCCL(g_nccl.AllReduce(dev, dev, (size_t)n, kNcclFloat64, op, c->comm, c->stream));
    PLBM_NCCL(g_nccl.AllReduce(dev + 8, dev + 8, 1, kNcclFloat64, PLBM_REDUCE_MAX, c->comm, c->stream));
    PLBM_NCCL(g_nccl.GroupEnd());
    PLBM_CUDA(cudaMemcpyAsync(stage, dev, sizeof(double) * 9, cudaMemcpyDeviceToHost, c->stream));
    PLBM_CUDA(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; ++i) values[i] = stage[8] > 0.0 ? __builtin_nan("") : stage[i];
    return PLBM_OK;
}

int comm_transport_is_p2p(const Grid& g) { return g.comm && g.comm->p2p ? 1 : 0; }
bool comm_pairs_agreed(const Grid& g) { return g.comm && g.comm->pairs_ok; }
bool comm_dual_agreed(const Grid& g) { return g.comm && g.comm->dual_ok && g.comm->p2p; }
bool comm_triples_level(const Grid& g, int* level)
{
    if (!g.comm) return false;
    *level = g.comm->triples_level;
    return true;
}

int comm_unique_id(void* id128)
{
    if (!id128) {
        set_error("comm_unique_id: null pointer");
        return PLBM_ERR_ARG;
    }
    int rc = load_nccl();
    if (rc) return rc;
    NcclUniqueId id;
    PLBM_NCCL(g_nccl.GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof(id));
    return PLBM_OK;
}

int comm_init(Grid& g, const void* id128, int rank, int nranks, int nx_global, int x_offset)
{
    if (g.comm) {
        set_error("comm_init: already initialised");
        return PLBM_ERR_STATE;
    }
    if (!id128 || nranks < 1 || rank < 0 || rank >= nranks || x_offset < 0 || x_offset + g.nx > nx_global) {
        set_error("comm_init: bad argument");
        return PLBM_ERR_ARG;
    }
    if (g.nx < 2 && nranks > 1) {
        set_error("comm_init: every slab needs at least 2 lines");
        return PLBM_ERR_ARG;
    }
    int rc = load_nccl();
    if (rc) return rc;
    Comm* c = new Comm();
    c->rank = rank;
    c->nranks = nranks;
    c->lo = (rank + nranks - 1) % nranks;
    c->hi = (rank + 1) % nranks;
    c->bytes = (size_t)PLBM_HALO_LINES * 9 * g.ld * g.esize();   // [3 lines][9 populations][ld]
    c->bytes9 = 9 * (size_t)g.ld * g.esize();
    NcclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
    if (r != 0) {
        delete c;
        return nccl_fail(r, "ncclCommInitRank");
    }
    PLBM_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    PLBM_CUDA(cudaEventCreateWithFlags(&c->ev_packed, cudaEventDisableTiming));
    for (int p = 0; p < 2; ++p) {
        PLBM_CUDA(cudaEventCreateWithFlags(&c->ev_halo[p], cudaEventDisableTiming));
        PLBM_CUDA(cudaEventCreateWithFlags(&c->ev_consumed[p], cudaEventDisableTiming));
        PLBM_CUDA(cudaMalloc(&c->halo_lo[p], c->bytes));
        PLBM_CUDA(cudaMalloc(&c->halo_hi[p], c->bytes));
    }
    PLBM_CUDA(cudaMalloc(&c->send_lo, c->bytes));
    PLBM_CUDA(cudaMalloc(&c->send_hi, c->bytes));
    PLBM_CUDA(cudaMalloc(&c->send9_lo, c->bytes9));
    PLBM_CUDA(cudaMalloc(&c->send9_hi, c->bytes9));
    PLBM_CUDA(cudaMalloc(&c->halo9_lo, c->bytes9));
    PLBM_CUDA(cudaMalloc(&c->halo9_hi, c->bytes9));
    PLBM_CUDA(cudaEventCreateWithFlags(&c->ev9_packed, cudaEventDisableTiming));
    PLBM_CUDA(cudaEventCreateWithFlags(&c->ev9_done, cudaEventDisableTiming));
    PLBM_CUDA(cudaEventCreateWithFlags(&c->ev_boundary, cudaEventDisableTiming));
    PLBM_CUDA(cudaEventCreateWithFlags(&c->ev_interior, cudaEventDisableTiming));
    {
        int prio_lo = 0, prio_hi = 0;  // numerically lower = higher priority
        PLBM_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        PLBM_CUDA(cudaStreamCreateWithPriority(&c->bstream, cudaStreamNonBlocking, prio_hi));
    }
    g.comm = c;
    g.nx_global = nx_global;
    g.x_offset = x_offset;
    return p2p_setup(g);
}

int comm_finalize(Grid& g)
{
    Comm* c = g.comm;
    if (!c) return PLBM_OK;
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (g.stream) cudaStreamSynchronize(g.stream);
    if (c->p2p && c->comm) {  // every rank leaves the ring together: nobody unmaps while a neighbour may still push
        quiesce(g);
        if (c->opened_lo) cudaIpcCloseMemHandle(c->peer_lo);
        if (c->opened_hi) cudaIpcCloseMemHandle(c->peer_hi);
        quiesce(g);
        c->halo_lo[0] = c->halo_lo[1] = c->halo_hi[0] = c->halo_hi[1] = nullptr;  // inside ipc_block
    }
    if (c->comm) g_nccl.CommDestroy(c->comm);
    if (c->ipc_block) cudaFree(c->ipc_block);
    if (c->ticket) cudaFree(c->ticket);
    if (c->bstream) cudaStreamSynchronize(c->bstream);
    if (c->ev_boundary) cudaEventDestroy(c->ev_boundary);
    if (c->ev_interior) cudaEventDestroy(c->ev_interior);
    if (c->bstream) cudaStreamDestroy(c->bstream);
    for (int p = 0; p < 2; ++p) {
        if (c->halo_lo[p]) cudaFree(c->halo_lo[p]);
        if (c->halo_hi[p]) cudaFree(c->halo_hi[p]);
        if (c->ev_halo[p]) cudaEventDestroy(c->ev_halo[p]);
        if (c->ev_consumed[p]) cudaEventDestroy(c->ev_consumed[p]);
    }
    if (c->send_lo) cudaFree(c->send_lo);
    if (c->send_hi) cudaFree(c->send_hi);
    if (c->send9_lo) cudaFree(c->send9_lo);
    if (c->send9_hi) cudaFree(c->send9_hi);
    if (c->halo9_lo) cudaFree(c->halo9_lo);
    if (c->halo9_hi) cudaFree(c->halo9_hi);
    if (c->ev9_packed) cudaEventDestroy(c->ev9_packed);
    if (c->ev9_done) cudaEventDestroy(c->ev9_done);
    g.fv_halo_lo = g.fv_halo_hi = nullptr;
    if (c->ev_packed) cudaEventDestroy(c->ev_packed);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    g.comm = nullptr;
    g.nx_global = g.nx;
    g.x_offset = 0;
    return PLBM_OK;
}

void comm_invalidate_halo(Grid& g)
{
    if (g.comm) g.comm->halo_valid = false;
}

// Post the ring exchange of the packed boundary lines into halo slot p (on the comm stream).
static int exchange(Grid& g, int p)
{
    Comm* c = g.comm;
    PLBM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_packed, 0));
    // do not overwrite halo[p] before the boundary kernels of two steps ago have read it
    PLBM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_consumed[p], 0));
    PLBM_NCCL(g_nccl.GroupStart());
    // posting order matters when lo == hi (2 ranks): first message lands in the peer's halo_hi
    PLBM_NCCL(g_nccl.Send(c->send_lo, c->bytes, kNcclInt8, c->lo, c->comm, c->stream));
    PLBM_NCCL(g_nccl.Send(c->send_hi, c->bytes, kNcclInt8, c->hi, c->comm, c->stream));
    PLBM_NCCL(g_nccl.Recv(c->halo_hi[p], c->bytes, kNcclInt8, c->hi, c->comm, c->stream));
    PLBM_NCCL(g_nccl.Recv(c->halo_lo[p], c->bytes, kNcclInt8, c->lo, c->comm, c->stream));
    PLBM_NCCL(g_nccl.GroupEnd());
    PLBM_CUDA(cudaEventRecord(c->ev_halo[p], c->stream));
    return PLBM_OK;
}


// p2p transport of the step loop.  Exchange number e (1, 2, ...) targets slot e & 1 and flag value e.
template <typename T> static int p2p_push(Grid& g, const T* f, cudaStream_t s)
{
    Comm* c = g.comm;
    const unsigned e = ++c->epoch;
    const int slot = (int)(e & 1u);
    const size_t sb = c->slot_bytes;
    const int n = PLBM_HALO_LINES * 9 * g.ld;
    k_halo_push<T><<<(n + 255) / 256, 256, 0, s>>>(f, (T*)(c->peer_lo + off_hi(sb, slot)), (T*)(c->peer_hi + off_lo(sb, slot)), g.nx, g.ld,
                                                  (unsigned*)(c->peer_lo + off_flag_hi(sb, slot)),
                                                  (unsigned*)(c->peer_hi + off_flag_lo(sb, slot)), e, c->ticket);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PLBM_CUDA(cudaGetLastError());
    return PLBM_OK;
}

static int p2p_wait(Grid& g, unsigned e, cudaStream_t s)
{
    Comm* c = g.comm;
    const int slot = (int)(e & 1u);
    WaitValue32Fn wv = wait_value32();
    CUresult r1 = wv((CUstream)s, (CUdeviceptr)(c->ipc_block + off_flag_lo(c->slot_bytes, slot)), e, CU_STREAM_WAIT_VALUE_GEQ);
    CUresult r2 = wv((CUstream)s, (CUdeviceptr)(c->ipc_block + off_flag_hi(c->slot_bytes, slot)), e, CU_STREAM_WAIT_VALUE_GEQ);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
        set_error("cuStreamWaitValue32 failed");
        return PLBM_ERR_COMM;
    }
    return PLBM_OK;
}

// One launch over lines [x0, x1): the step kernel (depth 1), the two-step kernel (2) or the three-step kernel (3).
template <typename T> static int lbm_range(Grid& g, LbmArgs<T> a, int depth, int x0, int x1, int model, cudaStream_t s)
{
    if (x1 <= x0) return PLBM_OK;
    if (depth == 3) return launch_lbm_multi<T>(g, a.src, a.dst, x0, x1, model, a.cp, 3, s, a.halo_lo, a.halo_hi, a.mid);
    if (depth == 2) return launch_lbm_pair<T>(g, a.src, a.dst, x0, x1, a.halo_lo, a.halo_hi, model, a.cp, s);
    a.x_begin = x0;
    a.x_end = x1;
    return launch_lbm<T>(a, model, true, g.variant, s);
}

// Both boundaries of the slab, lines [0, nb) and [nx - nb, nx): ONE launch for depths 1 and 2 (they are two lines each: a launch
// per side would leave the GPU four fifths empty twice), one launch per side for depth 3.
template <typename T> static int lbm_boundaries(Grid& g, LbmArgs<T> a, int depth, int nb, int model, cudaStream_t s)
{
    if (2 * nb >= g.nx) return lbm_range<T>(g, a, depth, 0, g.nx, model, s);  // the boundaries are the whole slab
    if (depth == 3) {
        int rc = lbm_range<T>(g, a, 3, 0, nb, model, s);
        return rc ? rc : lbm_range<T>(g, a, 3, g.nx - nb, g.nx, model, s);
    }
    if (depth == 2) return launch_lbm_pair_boundaries<T>(g, a.src, a.dst, nb, a.halo_lo, a.halo_hi, model, a.cp, s);
    a.x_begin = 0;
    a.x_end = g.nx;
    a.x_split = nb;
    a.x_skip = g.nx - 2 * nb;
    return launch_lbm<T>(a, model, true, g.variant, s);
}

// Lattice roles after one step or a fused triple (an odd number of reference swaps = one index swap) or after a fused pair (two
// reference swaps = the indices stay, the result sits in the buffer that was `inew`: the buffers trade places).
static void finish_steps(Grid& g, int depth, bool dual = false)
{
    if (dual) {  // a closing triple that also stored state n-1: `inew` becomes the third buffer, the source becomes the spare
        std::swap(g.iold, g.inew);
        lbm_adopt_spare_as_inew(g);
    } else if (depth == 2) {
        std::swap(g.f[g.iold - 1], g.f[g.inew - 1]);
        for (int b = 0; b < 128; ++b) std::swap(g.tmap[g.iold - 1][b], g.tmap[g.inew - 1][b]);
    } else {
        std::swap(g.iold, g.inew);
    }
    g.comm->halo_of_lattice = g.iold;
}

// lines per side the boundary launches cover = lines of the halo message (slabs too thin for that: two, or the whole slab)
static int boundary_lines(const Grid& g) { return g.nx >= 2 * PLBM_HALO_LINES ? PLBM_HALO_LINES : (g.nx >= 4 ? 2 : g.nx); }

// How many steps the next launch of a call advances: three while more than three remain and the ring takes triples for this
// collision, two while more than two remain, else one (the last step stays single so that lattice `inew` ends up holding state
// n-1 like the reference, see step_lbm_t).  Every rank computes the same sequence.
static LbmLaunch next_launch(const Grid& g, int model, int s, int nsteps)
{
    const Comm* c = g.comm;
    const bool triples = c->triples_level >= 0 && lbm_multi_applicable(g, model, 3) &&
                         (g.variant == 10 || lbm_triples_wanted(g, c->triples_level, model));
    const bool pairs = lbm_pair_variant(g.variant) && c->pairs_ok;
    const bool dual = c->dual_ok && c->p2p && g.variant == 0 && g.spare != nullptr;
    return lbm_next_launch(nsteps - s, triples, pairs, dual);
}
static int next_depth(const Grid& g, int model, int s, int nsteps) { return next_launch(g, model, s, nsteps).depth; }

// p2p transport, two streams per rank:
//   B (high priority): wait(interior of the previous launch) -> wait(neighbours' flags) -> both boundaries, one launch
//                      -> push the fresh boundary lines into the neighbours' halo slots, raise their flags
//   M (the grid's)   : wait(boundaries of the previous launch) -> interior launch
// The interior does not read the halo, so it starts at once and the boundary launch (0.2 of one round of blocks), the two
// flag waits and the push run BESIDE it instead of ahead of it (round 1: serial on one stream, ~35 us per launch, 1.2 % at
// eight GPUs).  Slot reuse is safe by the handshake: a rank pushes epoch e + 1 after its boundary launch e, which waited for
// the neighbours' flags e, which they raised after their boundary launch e - 1 -- the last reader of slot (e + 1) & 1.
template <typename T> static int p2p_lbm_steps(Grid& g, int model, const CollideParams<T>& cp, int nsteps)
{
    Comm* c = g.comm;
    int rc;
    if (nsteps <= 0) return PLBM_OK;
    cudaStream_t M = g.stream, B = c->bstream;
    if (!c->halo_valid || c->halo_of_lattice != g.iold) {
        // not preceded by a consuming step: make sure no neighbour still reads the slot we are about to fill
        PLBM_CUDA(cudaStreamSynchronize(B));
        if ((rc = quiesce(g))) return rc;
        if ((rc = p2p_push<T>(g, g.lat<T>(g.iold), M))) return rc;
        c->halo_valid = true;
        c->halo_of_lattice = g.iold;
    }
    // everything enqueued on M so far (previous calls, the push above) precedes the first boundary launch
    PLBM_CUDA(cudaEventRecord(c->ev_interior, M));
    bool first = true;
    for (int s = 0; s < nsteps;) {
        const LbmLaunch L = next_launch(g, model, s, nsteps);
        const int depth = L.depth;
        const unsigned e = c->epoch;
        const int slot = (int)(e & 1u);
        LbmArgs<T> a;
        a.src = g.lat<T>(g.iold);
        a.dst = g.lat<T>(g.inew);
        a.mid = L.dual ? static_cast<T*>(g.spare) : nullptr;
        a.nx = g.nx;
        a.ny = g.ny;
        a.ld = g.ld;
        a.halo_lo = (const T*)c->halo_lo[slot];
        a.halo_hi = (const T*)c->halo_hi[slot];
        a.cp = cp;
        // THREE boundary lines per side whatever the launch advances (one step would need one, a pair two): they need the halo,
        // and they are what the neighbours get next -- the message always carries three lines so that any launch may follow, and
        // every line of it must come from the boundary launch, which is all the push waits for (r02n2: with two-line boundaries
        // after a single step the third line was still the interior launch's to write, and a following triple read it stale)
        const int nb = boundary_lines(g);
        PLBM_CUDA(cudaStreamWaitEvent(B, c->ev_interior, 0));                  // interior of the previous launch (wrote src, read dst)
        if (!first) PLBM_CUDA(cudaStreamWaitEvent(M, c->ev_boundary, 0));      // boundaries of the previous launch (wrote src)
        if ((rc = p2p_wait(g, e, B))) return rc;                               // the neighbours' lines of lattice `iold` have landed
        if ((rc = lbm_boundaries<T>(g, a, depth, nb, model, B))) return rc;
        if ((rc = p2p_push<T>(g, a.dst, B))) return rc;
        PLBM_CUDA(cudaEventRecord(c->ev_boundary, B));
        a.halo_lo = a.halo_hi = nullptr;
        if ((rc = lbm_range<T>(g, a, depth, nb, g.nx - nb, model, M))) return rc;
        PLBM_CUDA(cudaEventRecord(c->ev_interior, M));
        finish_steps(g, depth, L.dual);
        s += depth;
        first = false;
    }
    PLBM_CUDA(cudaStreamWaitEvent(M, c->ev_boundary, 0));  // whatever follows on the grid's stream sees the whole lattice
    return PLBM_OK;
}

template <typename T> int comm_lbm_steps(Grid& g, int model, const CollideParams<T>& cp, int nsteps)
{
    Comm* c = g.comm;
    int rc;
    if (c->p2p) return p2p_lbm_steps<T>(g, model, cp, nsteps);
    if (nsteps > 0 && (!c->halo_valid || c->halo_of_lattice != g.iold)) {
        // first step after an initial condition / upload: exchange the boundary lines of `iold`
        if ((rc = launch_halo_pack<T>(g, g.lat<T>(g.iold), (T*)c->send_lo, (T*)c->send_hi, g.stream))) return rc;
        PLBM_CUDA(cudaEventRecord(c->ev_packed, g.stream));
        if ((rc = exchange(g, c->parity))) return rc;
        c->halo_valid = true;
        c->halo_of_lattice = g.iold;
    }
    for (int s = 0; s < nsteps;) {
        const int depth = next_depth(g, model, s, nsteps);
        const int p = c->parity;
        LbmArgs<T> a;
        a.src = g.lat<T>(g.iold);
        a.dst = g.lat<T>(g.inew);
        a.nx = g.nx;
        a.ny = g.ny;
        a.ld = g.ld;
        a.halo_lo = (const T*)c->halo_lo[p];
        a.halo_hi = (const T*)c->halo_hi[p];
        a.cp = cp;
        const int nb = boundary_lines(g);
        // the boundary lines of each side first: they need the neighbours' lines, and produce what must be sent
        PLBM_CUDA(cudaStreamWaitEvent(g.stream, c->ev_halo[p], 0));
        if ((rc = lbm_boundaries<T>(g, a, depth, nb, model, g.stream))) return rc;
        PLBM_CUDA(cudaEventRecord(c->ev_consumed[p], g.stream));
        if ((rc = launch_halo_pack<T>(g, a.dst, (T*)c->send_lo, (T*)c->send_hi, g.stream))) return rc;
        PLBM_CUDA(cudaEventRecord(c->ev_packed, g.stream));
        if ((rc = exchange(g, p ^ 1))) return rc;
        // interior, overlapped with the exchange
        // (the send buffers are re-packed next step, ordered after ev_halo[p^1], recorded after the sends completed)
        a.halo_lo = a.halo_hi = nullptr;
        if ((rc = lbm_range<T>(g, a, depth, nb, g.nx - nb, model, g.stream))) return rc;
        finish_steps(g, depth);
        c->parity = p ^ 1;
        s += depth;
    }
    return PLBM_OK;
}

// Exchange the boundary lines of all nine populations of lattice `f` with the ring neighbours and
// publish them as g.fv_halo_lo / g.fv_halo_hi (stream-ordered: the caller's next kernel on g.stream
// sees them).  rank r's line 0 goes to r-1's halo_hi, its line nx-1 to r+1's halo_lo.
template <typename T> int comm_fv_exchange(Grid& g, const T* f)
{
    Comm* c = g.comm;
    int rc;
    const size_t bytes9 = c->bytes9;
    if ((rc = launch_halo_pack9<T>(g, f, (T*)c->send9_lo, (T*)c->send9_hi, g.stream))) return rc;
    PLBM_CUDA(cudaEventRecord(c->ev9_packed, g.stream));
    PLBM_CUDA(cudaStreamWaitEvent(c->stream, c->ev9_packed, 0));
    PLBM_NCCL(g_nccl.GroupStart());
    PLBM_NCCL(g_nccl.Send(c->send9_lo, bytes9, kNcclInt8, c->lo, c->comm, c->stream));
    PLBM_NCCL(g_nccl.Send(c->send9_hi, bytes9, kNcclInt8, c->hi, c->comm, c->stream));
    PLBM_NCCL(g_nccl.Recv(c->halo9_hi, bytes9, kNcclInt8, c->hi, c->comm, c->stream));
    PLBM_NCCL(g_nccl.Recv(c->halo9_lo, bytes9, kNcclInt8, c->lo, c->comm, c->stream));
    PLBM_NCCL(g_nccl.GroupEnd());
    PLBM_CUDA(cudaEventRecord(c->ev9_done, c->stream));
    PLBM_CUDA(cudaStreamWaitEvent(g.stream, c->ev9_done, 0));
    g.fv_halo_lo = c->halo9_lo;
    g.fv_halo_hi = c->halo9_hi;
    c->halo_valid = false;  // the LBM halo slots are stale after an FVM/DUGKS step
    return PLBM_OK;
}
// Two boundary lines of a macroscopic field per direction (vorticity: d(uy)/dx reaches two lines into the neighbours'
// slabs).  Lines of an (nx, ny) field are contiguous, so the sends read the field itself; stream-ordered like above.
template <typename T> int comm_field_exchange2(Grid& g, const T* field, const T** lo, const T** hi)
{
    Comm* c = g.comm;
    const size_t bytes = 2 * (size_t)g.ny * sizeof(T);
    if (bytes > c->bytes9 || g.nx < 2) {
        set_error("comm_field_exchange2: slab too thin");
        return PLBM_ERR_ARG;
    }
    PLBM_CUDA(cudaEventRecord(c->ev9_packed, g.stream));
    PLBM_CUDA(cudaStreamWaitEvent(c->stream, c->ev9_packed, 0));
    PLBM_NCCL(g_nccl.GroupStart());
    PLBM_NCCL(g_nccl.Send(field, bytes, kNcclInt8, c->lo, c->comm, c->stream));
    PLBM_NCCL(g_nccl.Send(field + (size_t)(g.nx - 2) * g.ny, bytes, kNcclInt8, c->hi, c->comm, c->stream));
    PLBM_NCCL(g_nccl.Recv(c->halo9_hi, bytes, kNcclInt8, c->hi, c->comm, c->stream));
    PLBM_NCCL(g_nccl.Recv(c->halo9_lo, bytes, kNcclInt8, c->lo, c->comm, c->stream));
    PLBM_NCCL(g_nccl.GroupEnd());
    PLBM_CUDA(cudaEventRecord(c->ev9_done, c->stream));
    PLBM_CUDA(cudaStreamWaitEvent(g.stream, c->ev9_done, 0));
    g.fv_halo_lo = g.fv_halo_hi = nullptr;  // the 9-population halo buffers were reused
    *lo = (const T*)c->halo9_lo;
    *hi = (const T*)c->halo9_hi;
    return PLBM_OK;
}
template int comm_field_exchange2<double>(Grid&, const double*, const double**, const double**);
template int comm_field_exchange2<float>(Grid&, const float*, const float**, const float**);

template int comm_fv_exchange<double>(Grid&, const double*);
template int comm_fv_exchange<float>(Grid&, const float*);

template int comm_lbm_steps<double>(Grid&, int, const CollideParams<double>&, int);
template int comm_lbm_steps<float>(Grid&, int, const CollideParams<float>&, int);

}  // namespace plbm
