// FP64 pipe rate as a function of the operand pattern (B200, sm_100a): the usual peak microbenchmark
// x = fma(x, a, b) reads ONE changing register pair per instruction (a, b stay in the operand-reuse
// cache); the jet bodies of validate_kernel read three different pairs per DFMA (t[b], u[c], acc).
// 5 warps per scheduler (20 per SM), 8 independent chains per thread.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_operands dfma_operands.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double* out, long long* cyc, int iters, double a, double b) {
    double x[8], y[8], z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x + i; y[i] = 1.0 + 1e-9 * (threadIdx.x + i); z[i] = 1.0 - 1e-9 * (threadIdx.x * 3 + i); }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) x[i] = fma(x[i], a, b);                       // 1 changing operand
                if (MODE == 1) x[i] = fma(x[i], y[i], b);                    // 2 register operands
                if (MODE == 2) x[i] = fma(y[i], z[i], x[i]);                 // 3 different register pairs
                if (MODE == 3) x[i] = fma(y[(i + r) & 7], z[(i + 2 * r + 1) & 7], x[i]);   // 3 pairs, shuffled like a convolution
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + y[i] + z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* what) {
    const int warps = 20, iters = 4000;
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * 148 * warps * 32);
    cudaMalloc(&cyc, sizeof(long long));
    k<MODE><<<148, warps * 32>>>(out, cyc, iters, 1.0000001, 1e-9);
    k<MODE><<<148, warps * 32>>>(out, cyc, iters, 1.0000001, 1e-9);
    long long c = 0;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double per = (double)c / (iters * 32.0) / (warps / 4.0);
    printf("%-44s %.2f cycles per DFMA per scheduler  (%.1f %% of 64 DFMA/clk/SM)\n", what, per, 200.0 / per);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("x = fma(x, a, b)      (1 changing pair)");
    run<1>("x = fma(x, y, b)      (2 register pairs)");
    run<2>("x = fma(y, z, x)      (3 register pairs)");
    run<3>("x = fma(y[j], z[k], x) (3 pairs, shuffled)");
    return 0;
}
