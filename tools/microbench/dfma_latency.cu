// DFMA dependent-issue latency / throughput on one SM sub-partition (B200, sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_latency dfma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(double* out, long long* cyc, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x + k;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
#pragma unroll
            for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
void run(int warps_per_sm) {
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * 148 * 2048);
    cudaMalloc(&cyc, sizeof(long long));
    const int iters = 2000;
    chain<ILP><<<148, warps_per_sm * 32>>>(out, cyc, iters, 1.0000001, 1e-9);
    chain<ILP><<<148, warps_per_sm * 32>>>(out, cyc, iters, 1.0000001, 1e-9);
    long long c = 0;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double per_warp_dfma = (double)c / (iters * 16.0 * ILP);       // cycles between DFMA issues of one warp
    const double per_smsp = per_warp_dfma / (warps_per_sm / 4.0);          // cycles per DFMA per sub-partition
    printf("ILP %d warps/SM %2d: %.2f cycles per DFMA per warp, %.2f cycles per DFMA per SMSP (pipe floor 2.0)\n",
           ILP, warps_per_sm, per_warp_dfma, per_smsp);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {4, 8, 16, 32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
    return 0;
}
