// order.cu -- evaluation order for validate_kernel (see common.h).  Which group evaluates a candidate, and when, never
// changes its outputs (every output is addressed by candidate), so the order is free: sorting the batch by the op signature of
// its leading tokens puts candidates with the same micro-op prefix next to each other, the 20 warps of an SM then sit in the
// same interpreter bodies more often, and the instruction caches (L0 ~6 KB per scheduler, L1.5 32 KB; the interpreter's
// hot code is 26 KB) miss less: -3.0 % kernel time on the synthetic depth-5 batch for a sort that costs 0.2 %
// (tools/sorted_order_ab.py).  The sort itself is cub's radix sort (library plumbing, like the allocator).
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>
#include <stdlib.h>
#include "common.h"

namespace pde {

static __global__ void __launch_bounds__(256)
order_keys_kernel(const uint8_t* __restrict__ code, const unsigned* __restrict__ row_off, const uint8_t* __restrict__ len,
                  long long n, int L, unsigned long long* __restrict__ keys, int* __restrict__ idx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(row_off ? code + (size_t)row_off[i] * 16 : code + (size_t)i * L);
    int m = len[i];
    if (m > L) m = 0;                               // malformed rows (the kernel reports them) sort first
    // The key is the program's OP SIGNATURE: 12 leading tokens at 5 bits each, token 0 in the most significant bits.
    // Leaves are collapsed to their kind (coordinate / PRIM / constant) and POW(k) to one class: which coordinate or
    // constant a leaf is does not change the interpreter body it runs (138.3 ms as given, 134.2 sorted by the 7 leading
    // raw bytes, 133.6 by the signature; tools/sorted_order_ab.py)
    unsigned w[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) w[k] = (4 * k < m) ? __ldg(src + k) : 0u;
    unsigned long long key = 0;
#pragma unroll
    for (int t = 0; t < 12; ++t) {
        const unsigned c = (t < m) ? ((w[t >> 2] >> (8 * (t & 3))) & 0xffu) : 0u;
        unsigned cls;
        if (c == 0u) cls = 0u;
        else if (c < 0x08u) cls = 1u;                       // coordinate
        else if (c < 0x10u) cls = 2u;                       // PRIM(p)
        else if (c < 0x18u) cls = 4u + (c & 7u);            // ADD SUB MUL DIV
        else if (c < 0x20u) cls = 12u + (c & 7u);           // NEG ABS SQRT EXP
        else if (c < 0x40u) cls = 20u + (c & 7u);           // neg inv square pow_3_2 pow_neg_3_2 exp_neg
        else if (c < 0x80u) cls = 28u;                      // POW(k)
        else cls = 3u;                                      // CONST(k)
        key = (key << 5) | cls;
    }
    keys[i] = key << 4;
    idx[i] = (int)i;
}

int candidate_order(const uint8_t* code, const unsigned* row_off, const uint8_t* len, long long n, int L, int** order, void* stream) {
    *order = nullptr;
    static const bool off = getenv("PDE_B200_NO_ORDER") != nullptr;         // A/B switch
    // Large batches only: the sort's fixed cost (a dozen launches) is 0.15 ms, and batches that come out of the enumerator
    // or the normaliser are ordered by construction already (depth-4: 258 285 raw candidates 32.04 ms sorted vs 31.91 as
    // given) -- the gain is for long unordered streams like the synthetic depth-5 batch
    if (off || n < (1LL << 18) || n > 0x7fffffffLL || L < 12) return PDE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *k_in = nullptr, *k_out = nullptr;
    int *i_in = nullptr, *i_out = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    PDE_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, i_in, i_out, (int)n, 4, 64, st));
    int rc;
    if ((rc = scratch_alloc(reinterpret_cast<void**>(&k_in), sizeof(unsigned long long) * n, st))) return rc;
    if ((rc = scratch_alloc(reinterpret_cast<void**>(&k_out), sizeof(unsigned long long) * n, st))) return rc;
    if ((rc = scratch_alloc(reinterpret_cast<void**>(&i_in), sizeof(int) * n, st))) return rc;
    if ((rc = scratch_alloc(reinterpret_cast<void**>(&i_out), sizeof(int) * n, st))) return rc;
    if ((rc = scratch_alloc(&tmp, tmp_bytes, st))) return rc;
    order_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(code, row_off, len, n, L, k_in, i_in);
    count_launch();
    PDE_CUDA(cudaGetLastError());
    // bits 4..63 (the 12 tokens): a stable sort, so equal signatures keep the caller's order
    PDE_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, i_in, i_out, (int)n, 4, 64, st));
    scratch_free(k_in, st); scratch_free(k_out, st); scratch_free(i_in, st); scratch_free(tmp, st);
    *order = i_out;
    return PDE_OK;
}

}  // namespace pde
