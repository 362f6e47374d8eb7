// fp64_peak.cu — microbenchmark of the B200 FP64 pipe (DFMA / DADD / DMUL / ddiv / drcp+NR) and of
// a plain HBM copy, to place the second roof of the WENO5 stencil.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(256) pipe(double* out, double a, double b, int iters) {
    double x0 = a + threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        if (OP == 0) { x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a); x4 = fma(x4, b, a); x5 = fma(x5, b, a); x6 = fma(x6, b, a); x7 = fma(x7, b, a); }
        if (OP == 1) { x0 = x0 + b; x1 = x1 + b; x2 = x2 + b; x3 = x3 + b; x4 = x4 + b; x5 = x5 + b; x6 = x6 + b; x7 = x7 + b; }
        if (OP == 2) { x0 = x0 * b; x1 = x1 * b; x2 = x2 * b; x3 = x3 * b; x4 = x4 * b; x5 = x5 * b; x6 = x6 * b; x7 = x7 * b; }
        if (OP == 3) { x0 = a / x0; x1 = a / x1; x2 = a / x2; x3 = a / x3; x4 = a / x4; x5 = a / x5; x6 = a / x6; x7 = a / x7; }
        if (OP == 4) { x0 = fmax(x0, b) + a; x1 = fmax(x1, b) + a; x2 = fmax(x2, b) + a; x3 = fmax(x3, b) + a; x4 = fmax(x4, b) + a; x5 = fmax(x5, b) + a; x6 = fmax(x6, b) + a; x7 = fmax(x7, b) + a; }
        if (OP == 6) { x0 = (fabs(x0) > fabs(b) ? x0 : b) + a; x1 = (fabs(x1) > fabs(b) ? x1 : b) + a; x2 = (fabs(x2) > fabs(b) ? x2 : b) + a; x3 = (fabs(x3) > fabs(b) ? x3 : b) + a; x4 = (fabs(x4) > fabs(b) ? x4 : b) + a; x5 = (fabs(x5) > fabs(b) ? x5 : b) + a; x6 = (fabs(x6) > fabs(b) ? x6 : b) + a; x7 = (fabs(x7) > fabs(b) ? x7 : b) + a; }
        if (OP == 5) { x0 = sqrt(x0) + a; x1 = sqrt(x1) + a; x2 = sqrt(x2) + a; x3 = sqrt(x3) + a; x4 = sqrt(x4) + a; x5 = sqrt(x5) + a; x6 = sqrt(x6) + a; x7 = sqrt(x7) + a; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void copyk(const double2* __restrict__ a, double2* __restrict__ b, long n) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) b[i] = a[i];
}

template <int OP> void run(const char* name, double* d, int ops_per_iter) {
    const int iters = 4096, grid = 148 * 8, block = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    pipe<OP><<<grid, block>>>(d, 1.000001, 0.999999, 64);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); pipe<OP><<<grid, block>>>(d, 1.000001, 0.999999, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double ops = (double)grid * block * iters * 8.0 * ops_per_iter;
    printf("%-22s %8.3f ms  %8.2f Tops/s (lane-ops)  -> %6.2f lane-ops/clk/SM @1.965GHz\n", name, best, ops / best * 1e-9,
           ops / (best * 1e-3) / 148.0 / 1.965e9);
}

int main() {
    double* d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(double));
    run<0>("DFMA", d, 1);
    run<1>("DADD", d, 1);
    run<2>("DMUL", d, 1);
    run<3>("ddiv (a/x)", d, 1);
    run<4>("fmax+DADD", d, 2);
    run<5>("sqrt+DADD", d, 2);
    run<6>("(|x|>|b|?x:b)+DADD", d, 2);
    long n = 1L << 28;   // 4 GiB per buffer as double2
    double2 *a, *b; cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16); cudaMemset(a, 1, n * 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 6; ++r) {
        cudaEventRecord(e0); copyk<<<148 * 16, 512>>>(a, b, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    printf("copy 4 GiB -> 4 GiB     %8.3f ms  %8.1f GB/s (read+write)\n", best, 2.0 * n * 16 / best * 1e-6);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); printf("clockRate attr %d kHz\n", clk);
    return 0;
}
