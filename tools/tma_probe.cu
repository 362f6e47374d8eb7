// tma_probe.cu — which 3-D TMA box shapes / start coordinates work for Float64 tiles on sm_100a?
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void probe(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, int nelem, double* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    double* tile = reinterpret_cast<double*>(sm);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(sm + ((nelem * 8 + 127) / 128) * 128);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"((unsigned)__cvta_generic_to_shared(tile)), "l"(&map), "r"(c0), "r"(c1), "r"(c2), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(nelem * 8) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
    for (int i = threadIdx.x; i < nelem; i += blockDim.x) out[i] = tile[i];
}
int main() {
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    const int n = 128; const long N = (long)n * n * n;
    double* h = new double[N]; for (long i = 0; i < N; ++i) h[i] = (double)i;
    double *d, *out; cudaMalloc(&d, N * 8); cudaMalloc(&out, 64 * 64 * 8); cudaMemcpy(d, h, N * 8, cudaMemcpyHostToDevice);
    int tests[][5] = {{32, 16, 32, 32, 5}, {32, 16, 29, 29, 5}, {38, 22, 32, 32, 5}, {38, 22, 29, 29, 5}, {40, 22, 29, 29, 5}, {38, 16, 29, 29, 5}, {32, 22, 29, 29, 5}, {36, 22, 29, 29, 5}, {48, 22, 29, 29, 5}, {64, 22, 29, 29, 5}};
    for (auto& t : tests) {
        CUtensorMap m;
        cuuint64_t dims[3] = {n, n, n}, strides[2] = {n * 8ull, (cuuint64_t)n * n * 8ull};
        cuuint32_t box[3] = {(cuuint32_t)t[0], (cuuint32_t)t[1], 1}, es[3] = {1, 1, 1};
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        int nelem = t[0] * t[1];
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        probe<<<1, 128, ((nelem * 8 + 127) / 128) * 128 + 64>>>(m, t[2], t[3], t[4], nelem, out);
        cudaError_t e = cudaDeviceSynchronize();
        double v0 = -1, v1 = -1; if (e == cudaSuccess) { cudaMemcpy(&v0, out, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&v1, out + nelem - 1, 8, cudaMemcpyDeviceToHost); }
        double e0 = t[2] + (double)n * t[3] + (double)n * n * t[4], e1 = (t[2] + t[0] - 1) + (double)n * (t[3] + t[1] - 1) + (double)n * n * t[4];
        printf("box %2dx%2d at (%d,%d,%d): encode %d, run %s, first %.0f (exp %.0f) last %.0f (exp %.0f)\n", t[0], t[1], t[2], t[3], t[4], (int)r, cudaGetErrorString(e), v0, e0, v1, e1);
        if (e != cudaSuccess) { printf("stopping after sticky error\n"); break; }
    }
    return 0;
}
