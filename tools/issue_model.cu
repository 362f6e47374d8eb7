// issue_model.cu — does a DP instruction block the SMSP dispatch port for 2 cycles?  Interleave independent DFMA
// chains with independent integer (IMAD / LOP3) and FP32 (FFMA) chains at ratios 1:0, 1:1, 1:2 and time them.
#include <cstdio>
#include <cuda_runtime.h>
template <int NI, int KIND>
__global__ void __launch_bounds__(256) k(double* out, unsigned* iout, double a, double b, unsigned c, int iters) {
    double x0 = a + threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    unsigned i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3, i4 = i0 + 4, i5 = i0 + 5, i6 = i0 + 6, i7 = i0 + 7;
    unsigned j0 = i0 * 3, j1 = j0 + 1, j2 = j0 + 2, j3 = j0 + 3, j4 = j0 + 4, j5 = j0 + 5, j6 = j0 + 6, j7 = j0 + 7;
    float f0 = threadIdx.x, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3, f4 = f0 + 4, f5 = f0 + 5, f6 = f0 + 6, f7 = f0 + 7;
    float g0 = f0 * 3, g1 = g0 + 1, g2 = g0 + 2, g3 = g0 + 3, g4 = g0 + 4, g5 = g0 + 5, g6 = g0 + 6, g7 = g0 + 7;
    const float fb = (float)b, fa = (float)a;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a);
        x4 = fma(x4, b, a); x5 = fma(x5, b, a); x6 = fma(x6, b, a); x7 = fma(x7, b, a);
        if (NI >= 1) {
            if (KIND == 0) { i0 = (i0 ^ c) + it; i1 = (i1 ^ c) + it; i2 = (i2 ^ c) + it; i3 = (i3 ^ c) + it; i4 = (i4 ^ c) + it; i5 = (i5 ^ c) + it; i6 = (i6 ^ c) + it; i7 = (i7 ^ c) + it; }
            if (KIND == 1) { i0 = i0 * c + it; i1 = i1 * c + it; i2 = i2 * c + it; i3 = i3 * c + it; i4 = i4 * c + it; i5 = i5 * c + it; i6 = i6 * c + it; i7 = i7 * c + it; }
            if (KIND == 2) { f0 = fmaf(f0, fb, fa); f1 = fmaf(f1, fb, fa); f2 = fmaf(f2, fb, fa); f3 = fmaf(f3, fb, fa); f4 = fmaf(f4, fb, fa); f5 = fmaf(f5, fb, fa); f6 = fmaf(f6, fb, fa); f7 = fmaf(f7, fb, fa); }
        }
        if (NI >= 2) {
            if (KIND == 0) { j0 = (j0 ^ c) + it; j1 = (j1 ^ c) + it; j2 = (j2 ^ c) + it; j3 = (j3 ^ c) + it; j4 = (j4 ^ c) + it; j5 = (j5 ^ c) + it; j6 = (j6 ^ c) + it; j7 = (j7 ^ c) + it; }
            if (KIND == 1) { j0 = j0 * c + it; j1 = j1 * c + it; j2 = j2 * c + it; j3 = j3 * c + it; j4 = j4 * c + it; j5 = j5 * c + it; j6 = j6 * c + it; j7 = j7 * c + it; }
            if (KIND == 2) { g0 = fmaf(g0, fb, fa); g1 = fmaf(g1, fb, fa); g2 = fmaf(g2, fb, fa); g3 = fmaf(g3, fb, fa); g4 = fmaf(g4, fb, fa); g5 = fmaf(g5, fb, fa); g6 = fmaf(g6, fb, fa); g7 = fmaf(g7, fb, fa); }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7 + g0 + g1 + g2 + g3 + g4 + g5 + g6 + g7;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = i0 + i1 + i2 + i3 + i4 + i5 + i6 + i7 + j0 + j1 + j2 + j3 + j4 + j5 + j6 + j7;
}
template <int NI, int KIND> void run(const char* name, double* d, unsigned* di) {
    const int iters = 4096, grid = 148 * 8, block = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NI, KIND><<<grid, block>>>(d, di, 1.000001, 0.999999, 12345u, 64);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) { cudaEventRecord(e0); k<NI, KIND><<<grid, block>>>(d, di, 1.000001, 0.999999, 12345u, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    // cycles per (DFMA + NI other) group per warp per SMSP: 8 warps/SMSP resident (148*8 blocks*8 warps / (148*4))
    double groups = (double)iters * 8.0;                 // per warp
    double warps_per_smsp = 8.0 * 8.0 / 4.0;
    double cyc = best * 1e-3 * 1.965e9 / (groups * warps_per_smsp);
    printf("%-28s %8.3f ms   %.2f cycles per {1 DFMA + %d other} per SMSP\n", name, best, cyc, NI);
}
int main() {
    double* d; unsigned* di; cudaMalloc(&d, 148 * 8 * 256 * 8); cudaMalloc(&di, 148 * 8 * 256 * 4);
    run<0, 0>("DFMA only", d, di);
    run<1, 0>("DFMA + 1 LOP3/IADD (2 ALU)", d, di);
    run<2, 0>("DFMA + 2x(LOP3+IADD)", d, di);
    run<1, 1>("DFMA + 1 IMAD", d, di);
    run<2, 1>("DFMA + 2 IMAD", d, di);
    run<1, 2>("DFMA + 1 FFMA", d, di);
    run<2, 2>("DFMA + 2 FFMA", d, di);
    return 0;
}
