// Development probe: which (box, coordinates, dtype) combinations does cp.async.bulk.tensor.3d accept?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int c2, unsigned bytes, double* out, int nout)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + 65536);
    unsigned b = (unsigned)__cvta_generic_to_shared(bar), d = (unsigned)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(d),
                     "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(b)
                     : "memory");
    }
    unsigned ok = 0;
    long long spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(0u) : "memory");
        if (!ok && ++spins > (1ll << 24)) { if (threadIdx.x == 0) out[0] = -12345.0; return; }
    } while (!ok);
    for (int i = threadIdx.x; i < nout; i += blockDim.x) out[i] = reinterpret_cast<double*>(smem)[i];
}

int main(int argc, char** argv)
{
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    if (!enc) { printf("no encode fn\n"); return 1; }
    const int ny = 64, nx = 64, ld = 64;
    std::vector<double> h((size_t)ld * nx * 9);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (double)i;
    double *d, *out;
    cudaMalloc(&d, h.size() * 8);
    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    cudaMalloc(&out, 65536);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 64);
    struct Case { int by, bx, bq, c0, c1, c2; const char* name; } cases[] = {
        {32, 8, 1, 0, 0, 0, "box 32x8x1 @0,0,0"},
        {32, 8, 9, 0, 0, 0, "box 32x8x9 @0,0,0"},
        {34, 10, 9, 0, 0, 0, "box 34x10x9 @0,0,0"},
        {34, 10, 9, 1, 1, 0, "box 34x10x9 @1,1,0"},
        {34, 10, 9, -1, -1, 0, "box 34x10x9 @-1,-1,0"},
        {34, 10, 9, 31, 23, 0, "box 34x10x9 @31,23,0"},
        {34, 10, 9, 31, 55, 0, "box 34x10x9 @31,55,0 (overhang)"},
        {36, 10, 9, -1, -1, 0, "box 36x10x9 @-1,-1,0"},
    };
    Case one = {34, 10, 9, 0, 0, 0, "cli"};
    if (argc >= 7) { one.by = atoi(argv[1]); one.bx = atoi(argv[2]); one.bq = atoi(argv[3]); one.c0 = atoi(argv[4]); one.c1 = atoi(argv[5]); one.c2 = atoi(argv[6]); }
    for (auto& c0_ : cases) {
        Case c = argc >= 7 ? one : c0_;
        CUtensorMap m;
        cuuint64_t gdim[3] = {(cuuint64_t)ny, (cuuint64_t)nx, 9}, gstr[2] = {(cuuint64_t)ld * 8, (cuuint64_t)ld * nx * 8};
        cuuint32_t box[3] = {(cuuint32_t)c.by, (cuuint32_t)c.bx, (cuuint32_t)c.bq}, es[3] = {1, 1, 1};
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%-40s encode failed %d\n", c.name, (int)r); continue; }
        unsigned bytes = (unsigned)(c.by * c.bx * c.bq * 8);
        cudaMemset(out, 0, 65536);
        probe<<<1, 128, 65536 + 64>>>(m, c.c0, c.c1, c.c2, bytes, out, 64);
        cudaError_t e = cudaDeviceSynchronize();
        double o[64] = {0};
        if (e == cudaSuccess) cudaMemcpy(o, out, sizeof(o), cudaMemcpyDeviceToHost);
        printf("%-40s bytes=%6u -> %s  first=%g second=%g  [by]=%g\n", c.name, bytes, cudaGetErrorString(e), o[0], o[1], o[c.by < 64 ? c.by : 0]);
        if (e != cudaSuccess) { printf("sticky error, stopping\n"); return 2; }
        if (argc >= 7) break;
    }
    return 0;
}
