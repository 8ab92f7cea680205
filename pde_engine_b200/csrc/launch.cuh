// launch.cuh -- launch plumbing of the stage-2 kernel, shared by the translation units that instantiate it
// (pde_b200.cu: the built-in residuals; program.cu: run-time residual programs).  Everything here is `static`:
// each unit uploads the __constant__ tables of ITS copy of validate.cuh.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include "common.h"
#include "validate.cuh"

namespace pde {

// Proposed rejections of the first pass (cleared survivor bits) -> index list for the confirmation pass.
// Order is whatever the atomics give (roughly the evaluation order of the first pass): every output of the confirmation
// pass is addressed by candidate.
static __global__ void __launch_bounds__(256)
compact_rejects_kernel(const unsigned* __restrict__ bits, long long n, const int* __restrict__ order, int* __restrict__ index,
                       unsigned long long* __restrict__ count) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long i = (t < n && order) ? (long long)order[t] : t;          // walk the batch in its evaluation order
    const bool rej = t < n && !((bits[i >> 5] >> (i & 31)) & 1u);
    const unsigned b = __ballot_sync(0xffffffffu, rej);
    if (!b) return;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(count, (unsigned long long)__popc(b));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (rej) index[base + __popc(b & ((1u << lane) - 1u))] = (int)i;
}


static int upload_tables(const pde_session* s, double tau, double t0, cudaStream_t st) {
    double cv[PDE_N_CONST], pv[PDE_N_POW];
    pde_session_tables(s, cv, nullptr, pv, nullptr);
    double rv[PDE_N_CONST];
    for (int i = 0; i < PDE_N_CONST; ++i) rv[i] = 1.0 / cv[i];
    PDE_CUDA(cudaMemcpyToSymbolAsync(c_const, cv, sizeof(cv), 0, cudaMemcpyHostToDevice, st));
    PDE_CUDA(cudaMemcpyToSymbolAsync(c_rconst, rv, sizeof(rv), 0, cudaMemcpyHostToDevice, st));
    PDE_CUDA(cudaMemcpyToSymbolAsync(c_pow, pv, sizeof(pv), 0, cudaMemcpyHostToDevice, st));
    // Taylor-ratio rows: x**k has f_{j+1}/f_j = (k - j)/(j + 1) / x_0
    double fr[kNRows][4];
    for (int sl = 0; sl < kNRows; ++sl) {
        for (int j = 0; j < 4; ++j) fr[sl][j] = (pv[sl] - j) / (j + 1);
    }
    PDE_CUDA(cudaMemcpyToSymbolAsync(c_frow, fr, sizeof(fr), 0, cudaMemcpyHostToDevice, st));
    // x**n, n = 0, 1, 2, ... <= 64: binomial coefficients C(n, 0..4) (the division-free Taylor coefficients)
    int pi[kNRows];
    double fb[kNRows][5];
    for (int sl = 0; sl < kNRows; ++sl) {
        const double k = pv[sl];
        pi[sl] = (k >= 0.0 && k <= 64.0 && k == floor(k)) ? (int)k : -1;
        fb[sl][0] = 1.0;
        for (int j = 0; j < 4; ++j) fb[sl][j + 1] = fb[sl][j] * (k - j) / (j + 1);
    }
    PDE_CUDA(cudaMemcpyToSymbolAsync(c_pow_int, pi, sizeof(pi), 0, cudaMemcpyHostToDevice, st));
    PDE_CUDA(cudaMemcpyToSymbolAsync(c_fbin, fb, sizeof(fb), 0, cudaMemcpyHostToDevice, st));
    // round-off majorants: theta_n / W = 2 eps n! / (t0^n tau) (+ 0.2 % for the float32 arithmetic of the majorants)
    const float t0f = (float)t0;
    double th[4], fact = 1.0, tp = 1.0;
    for (int n = 1; n <= 4; ++n) {
        fact *= n; tp *= (double)t0f;
        th[n - 1] = 2.0 * 2.220446049250313e-16 * fact / (tp * tau) * 1.002;
    }
    float cf[PDE_N_CONST], rf[PDE_N_CONST], pf[PDE_N_POW];
    for (int i = 0; i < PDE_N_CONST; ++i) { cf[i] = fmaxf((float)fabs(cv[i]), 1e-18f); rf[i] = fmaxf((float)fabs(rv[i]), 1e-18f); }
    for (int i = 0; i < PDE_N_POW; ++i) pf[i] = (float)pv[i];
    PDE_CUDA(cudaMemcpyToSymbolAsync(c_constf, cf, sizeof(cf), 0, cudaMemcpyHostToDevice, st));
    PDE_CUDA(cudaMemcpyToSymbolAsync(c_rconstf, rf, sizeof(rf), 0, cudaMemcpyHostToDevice, st));
    PDE_CUDA(cudaMemcpyToSymbolAsync(c_powf, pf, sizeof(pf), 0, cudaMemcpyHostToDevice, st));
    PDE_CUDA(cudaMemcpyToSymbolAsync(c_t0, &t0f, sizeof(t0f), 0, cudaMemcpyHostToDevice, st));
    PDE_CUDA(cudaMemcpyToSymbolAsync(c_theta, th, sizeof(th), 0, cudaMemcpyHostToDevice, st));
    return PDE_OK;
}

template <int PROBLEM, bool DUMP, int W, int NP, int MINB, bool MAJ>
static int launch_validate_cfg(const ValidateParams& vp, cudaStream_t st, bool* fits) {
    constexpr int N = Residual<PROBLEM>::N;
    const size_t smem = cta_smem_bytes<N, NP>(vp.L, vp.ns, W);
    auto kern = validate_kernel<PROBLEM, DUMP, W, NP, MINB, MAJ>;
    int dev = 0, sms = 0, occ = 0, max_smem = 0;
    PDE_CUDA(cudaGetDevice(&dev));
    PDE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    PDE_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (smem > (size_t)max_smem) { if (fits) { *fits = false; return PDE_OK; } set_error("validate kernel does not fit: smem %zu B per block", smem); return PDE_E_INVALID; }
    PDE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PDE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, W * 32, smem));
    if (occ < 1) { if (fits) { *fits = false; return PDE_OK; } set_error("validate kernel does not fit: smem %zu B per block", smem); return PDE_E_INVALID; }
    if (fits) *fits = true;
    const long long rounds = (vp.n + W - 1) / W;
    const long long resident = (long long)sms * occ;    // persistent grid: a multiple of the SM count
    int grid = (int)(rounds < resident ? rounds : resident);
    if (grid < 1) grid = 1;
    // the dynamic chunk counter: 8 bytes from the library's stream-ordered pool, zeroed and released on the stream
    ValidateParams v = vp;
    unsigned long long* ctr = nullptr;
    static const bool static_deal = getenv("PDE_B200_STATIC_DEAL") != nullptr;      // A/B switch: round-robin chunks
    if (!static_deal) {
        if (int rc = scratch_alloc(reinterpret_cast<void**>(&ctr), sizeof(unsigned long long), st)) return rc;
        PDE_CUDA(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), st));
    }
    v.chunk_counter = ctr;
    kern<<<grid, W * 32, smem, st>>>(v);
    count_launch();
    const cudaError_t le = cudaGetLastError();
    scratch_free(ctr, st);
    if (le != cudaSuccess) return cuda_fail((int)le, "validate_kernel launch");
    return PDE_OK;
}


// The two-pass filter of pde_validate (include/pde_b200.h).  LAUNCHER::launch<MAJ>(vp, stream) starts the reduce-mode
// kernel of one residual with / without the round-off majorants.
template <class LAUNCHER>
static int run_two_pass(ValidateParams vp, const pde_validate_out* out, int confirm_points, cudaStream_t st) {
    // evaluation order of the batch (common.h: candidate_order; null for small batches)
    int* order = nullptr;
    int rc = candidate_order(vp.code, vp.row_off, vp.len, vp.n, vp.L, &order, st);
    if (rc) return rc;
    vp.index = order; vp.n_index = nullptr; vp.is_confirm = 0;
    if (confirm_points == 0) {
        // one pass over the whole grid with the majorants carried
        rc = LAUNCHER::template launch<true>(vp, st);
        scratch_free(order, st);
        return rc;
    }
    // pass 1: all P points, no majorants -- proposes rejections
    rc = LAUNCHER::template launch<false>(vp, st);
    if (rc) { scratch_free(order, st); return rc; }
    // pass 2: the proposed rejections again on the first confirm_points points WITH the majorants; a rejection
    // stands only if this pass votes it too (include/pde_b200.h)
    unsigned long long* cnt = reinterpret_cast<unsigned long long*>(out->scratch);
    int* index = out->scratch + 2;
    PDE_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), st));
    if (out->confirm) PDE_CUDA(cudaMemsetAsync(out->confirm, 0xff, sizeof(int32_t) * 2 * (size_t)vp.n, st));   // -1: not re-examined
    compact_rejects_kernel<<<(unsigned)((vp.n + 255) / 256), 256, 0, st>>>(out->survivor_bits, vp.n, order, index, cnt);
    count_launch();
    scratch_free(order, st);
    PDE_CUDA(cudaGetLastError());
    vp.index = index; vp.n_index = cnt; vp.is_confirm = 1; vp.confirm = out->confirm; vp.P_eval = confirm_points;
    return LAUNCHER::template launch<true>(vp, st);
}

}  // namespace pde
