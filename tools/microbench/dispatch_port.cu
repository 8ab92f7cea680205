// Does a half-rate FP64 warp instruction leave an issue slot for another pipe in its second cycle?
// (B200, sm_100a.)  Each loop iteration issues K independent DFMAs and M independent integer
// instructions (LOP3/IADD on the ALU pipe), 5 warps per scheduler.  Cycles per iteration per scheduler:
//     2K + M   if the FP64 instruction holds the dispatch port for two cycles (no overlap)
//     max(2K, K + M)   if the second cycle can issue to another pipe
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dispatch_port dispatch_port.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int K, int M>
__global__ void mix(double* out, unsigned* iout, long long* cyc, int iters, double a, double b, unsigned m) {
    double x[8];
    unsigned y[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { x[k] = threadIdx.x + k; y[k] = threadIdx.x * 7 + k; }
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int k = 0; k < K; ++k) x[k % 8] = fma(x[k % 8], a, b);
#pragma unroll
            for (int k = 0; k < M; ++k) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[k % 8]) : "r"(m), "r"(y[(k + 3) % 8]));
        }
    }
    long long t1 = clock64();
    double s = 0; unsigned u = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { s += x[k]; u ^= y[k]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = u;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int K, int M>
void run() {
    const int warps = 20;      // 5 per scheduler, like validate_kernel
    double* out; unsigned* iout; long long* cyc;
    cudaMalloc(&out, sizeof(double) * 148 * warps * 32);
    cudaMalloc(&iout, sizeof(unsigned) * 148 * warps * 32);
    cudaMalloc(&cyc, sizeof(long long));
    const int iters = 4000;
    mix<K, M><<<148, warps * 32>>>(out, iout, cyc, iters, 1.0000001, 1e-9, 0x9e3779b9u);
    mix<K, M><<<148, warps * 32>>>(out, iout, cyc, iters, 1.0000001, 1e-9, 0x9e3779b9u);
    long long c = 0;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double per_iter_smsp = (double)c / (iters * 4.0) / 1.0 / (warps / 4.0) * 1.0;   // one warp's time / iterations, x (1 / warps per scheduler)
    // all 5 warps of a scheduler run the same loop concurrently: time per iteration per scheduler = warp time / iterations * ... 
    const double cyc_per_iter = (double)c / (iters * 4.0);          // cycles one warp needs per (K DFMA + M ALU) group
    const double per_sched = cyc_per_iter / (warps / 4.0);          // the scheduler completes 5 groups in that time
    printf("K=%2d DFMA + M=%2d ALU: %.2f cycles per group per scheduler   (no overlap 2K+M = %d, overlap max(2K,K+M) = %d)\n",
           K, M, per_sched, 2 * K + M, (2 * K > K + M) ? 2 * K : K + M);
    (void)per_iter_smsp;
    cudaFree(out); cudaFree(iout); cudaFree(cyc);
}

int main() {
    run<8, 0>(); run<8, 4>(); run<8, 8>(); run<8, 16>(); run<8, 24>(); run<4, 16>(); run<0, 16>();
    return 0;
}
