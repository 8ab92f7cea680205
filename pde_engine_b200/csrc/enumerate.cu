// enumerate.cu -- stage 1 of the hot path: the combinatorial generator.
//
// Replaces the candidate loops of FastExpressionGenerator.stream_generate
// (lean_normalizer/lean_bridge_fixed.py:139-195).  The reference's loop nest
//     unary : for expr in E[d-1]: for op in 8 unary ops               LBF:142-153
//     binary: for d1: for e1 in E[d1]: for e2 in E[d-d1]: for op      LBF:155-195
// is flattened into a dense slot index; every slot evaluates the reference's
// prune predicates on precomputed string attributes + lexicographic ranks
// (LBF:134-136,143-152,162-195), an order-preserving compaction (block scan +
// scanned block sums) assigns the candidate index the reference's list would
// have, and the kept slots splice their operands' term-structured bytecode
// exactly as the reference's *textual* templates parse (LBF:170-195):
//     add   terms(a) ++ terms(b)
//     sub   terms(a) ++ [-t1(b)] ++ terms(b)[1:]
//     mul   terms(a)[:-1] ++ [last(a) * first(b)] ++ terms(b)[1:]
//     div   terms(a)[:-1] ++ [last(a) / (b)]
//     geom  terms(a)[:-1] ++ [last(a) / (1 - t1(b) +- t2(b) ...)]
//     unary op(whole(a))
// HBM-bound: 69 algorithmic bytes per candidate (L = 48): code L + len 1 +
// hash 8 + triple 12.  Rows are assembled in shared memory and written with
// coalesced 16-byte stores.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.h"

namespace pde {

constexpr int kEnumThreads = 128;
constexpr int kMaxDepth = 8;
constexpr int kMaxRow = 256;

struct EnumParams {
    const uint8_t* flags;
    const uint8_t* attrs;
    const uint32_t* rank;
    const uint32_t* term_begin;
    const int8_t* term_sign;
    const uint32_t* term_off;
    const uint8_t* pool;
    int depth;
    int prune;
    int depth_begin[kMaxDepth + 1];
    long long seg_begin[kMaxDepth + 1];  // dense slot index where segment k starts (0 = unary)
    long long n_slots;
};

// enumerator unary op index -> opcode (iteration order of UNARY_OPS, expression_operations.py:80-89)
__device__ const uint8_t kUnary[8] = {PDE_OP_FN_NEG, PDE_OP_FN_INV, PDE_OP_SQRT, PDE_OP_FN_SQUARE,
                                      PDE_OP_FN_POW32, PDE_OP_FN_POWN32, PDE_OP_EXP, PDE_OP_FN_EXPNEG};
// synthetic leaves: rho, z, PRIM(0) = rho**2 + z**2, PRIM(1) = rho/z, 1  (problems/__init__.py:73-79)
__device__ const uint8_t kLeaf[5] = {PDE_OP_VAR0, PDE_OP_VAR1, PDE_OP_PRIM0, PDE_OP_PRIM0 + 1, PDE_OP_CONST0};

struct Slot {
    int op;       // 0..7 unary, 8..12 binary
    int a, b;     // global operand indices after the add/mul swap (b = -1 for unary)
    bool keep;
};

__device__ __forceinline__ Slot decode_slot(const EnumParams& p, long long s) {
    Slot r;
    r.keep = true;
    const int d = p.depth;
    if (s < p.seg_begin[1]) {
        // unary, LBF:142-153
        const int e = (int)(s >> 3);
        r.op = (int)(s & 7);
        r.a = p.depth_begin[d - 2] + e;
        r.b = -1;
        if (p.prune) {
            const unsigned at = p.attrs[r.a];
            if (!(at & PDE_ATTR_HAS_VARS)) r.keep = false;
            if (r.op == 1 && (at & PDE_ATTR_STARTS_INV)) r.keep = false;
            if (r.op >= 2 && r.op <= 5 && (at & PDE_ATTR_IS_ONE)) r.keep = false;
        }
        return r;
    }
    int d1 = 1;
    while (d1 < d - 1 && s >= p.seg_begin[d1 + 1]) ++d1;
    const unsigned t = (unsigned)(s - p.seg_begin[d1]);      // every segment is < 2^32 slots (checked on the host)
    const int d2 = d - d1;
    const unsigned n2 = (unsigned)(p.depth_begin[d2] - p.depth_begin[d2 - 1]);
    const unsigned pair = t / 5u;
    const int bop = (int)(t - pair * 5u);
    const unsigned i1 = pair / n2;
    int a = p.depth_begin[d1 - 1] + (int)i1;
    int b = p.depth_begin[d2 - 1] + (int)(pair - i1 * n2);
    const unsigned ata = p.attrs[a], atb = p.attrs[b];
    if (p.prune && !((ata | atb) & PDE_ATTR_HAS_VARS)) r.keep = false;
    const uint32_t ra = p.rank[a], rb = p.rank[b];
    if ((bop == 0 || bop == 2) && ra > rb) { int tmp = a; a = b; b = tmp; }   // LBF:168-169
    const bool one_a = p.attrs[a] & PDE_ATTR_IS_ONE, one_b = p.attrs[b] & PDE_ATTR_IS_ONE;
    if (p.prune) {
        if (bop == 1 && ra == rb) r.keep = false;                   // a - a
        if (bop == 2 && (one_a || one_b)) r.keep = false;           // * 1
        if (bop == 3 && (one_b || ra == rb)) r.keep = false;        // / 1, a / a
        if (bop == 4 && one_b) r.keep = false;                      // 1 - 1
    }
    r.op = 8 + bop;
    r.a = a;
    r.b = b;
    return r;
}

__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    int base = 0;
    total = 0;
#pragma unroll
    for (int w = 0; w < kEnumThreads / 32; ++w) {
        const int c = s_warp[w];
        if (w < warp) base += c;
        total += c;
    }
    __syncthreads();
    return base + x - v;
}

__global__ void __launch_bounds__(kEnumThreads) enum_count_kernel(const EnumParams p, unsigned* block_sums) {
    __shared__ int s_warp[kEnumThreads / 32];
    const long long s = (long long)blockIdx.x * kEnumThreads + threadIdx.x;
    int keep = 0;
    if (s < p.n_slots) keep = decode_slot(p, s).keep ? 1 : 0;
    int total;
    block_exclusive_scan(keep, s_warp, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = (unsigned)total;
}

// single block: exclusive scan of block_sums -> block_off (int64), total
__global__ void __launch_bounds__(1024) scan_sums_kernel(const unsigned* sums, long long* off, int nblocks, long long* total) {
    __shared__ long long s_part[1024];
    const int per = (nblocks + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = min(lo + per, nblocks);
    long long acc = 0;
    for (int i = lo; i < hi; ++i) acc += sums[i];
    s_part[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long run = 0;
        for (int i = 0; i < 1024; ++i) { long long c = s_part[i]; s_part[i] = run; run += c; }
        *total = run;
    }
    __syncthreads();
    long long run = s_part[threadIdx.x];
    for (int i = lo; i < hi; ++i) { off[i] = run; run += sums[i]; }
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27; x *= 0x94D049BB133111EBULL;
    x ^= x >> 31;
    return x;
}

// 64-bit structural hash of (len, bytes); row is zero padded to a multiple of 8
__device__ __forceinline__ unsigned long long hash_row(const uint8_t* row, int len) {
    unsigned long long h = 0x9E3779B97F4A7C15ULL ^ ((unsigned long long)len * 0xD6E8FEB86659FD93ULL);
    const int nw = (len + 7) >> 3;
    const unsigned long long* w64 = reinterpret_cast<const unsigned long long*>(row);   // rows are 16-byte aligned
    for (int w = 0; w < nw; ++w) h = mix64(h ^ w64[w]);                                  // little endian
    return h;
}

struct RowWriter {
    uint8_t* row;
    int cap;
    int n;
    bool ok;
    __device__ void put(unsigned b) { if (n < cap) row[n] = (uint8_t)b; else ok = false; ++n; }
    __device__ void copy(const uint8_t* src, int len) { for (int i = 0; i < len; ++i) put(src[i]); }
};

// emit `whole` form of terms [t0, t1) of an expression; first_sign overrides the sign of the first term
__device__ __forceinline__ void emit_terms(RowWriter& w, const EnumParams& p, uint32_t t0, uint32_t t1, bool fresh, int first_sign_mul) {
    for (uint32_t t = t0; t < t1; ++t) {
        const uint32_t b0 = p.term_off[t], b1 = p.term_off[t + 1];
        w.copy(p.pool + b0, (int)(b1 - b0));
        int sg = p.term_sign[t];
        if (t == t0) sg *= first_sign_mul;
        if (fresh && t == t0) { if (sg < 0) w.put(PDE_OP_NEG); }
        else w.put(sg > 0 ? PDE_OP_ADD : PDE_OP_SUB);
    }
}

__device__ void splice(RowWriter& w, const EnumParams& p, const Slot& sl) {
    const uint32_t a0 = p.term_begin[sl.a], a1 = p.term_begin[sl.a + 1];
    if (sl.op < 8) {
        emit_terms(w, p, a0, a1, true, 1);
        w.put(kUnary[sl.op]);
        return;
    }
    const uint32_t b0 = p.term_begin[sl.b], b1 = p.term_begin[sl.b + 1];
    const int bop = sl.op - 8;
    if (bop == 0) {            // terms(a) ++ terms(b)
        emit_terms(w, p, a0, a1, true, 1);
        emit_terms(w, p, b0, b1, false, 1);
    } else if (bop == 1) {     // terms(a) ++ [-t1(b)] ++ rest(b)
        emit_terms(w, p, a0, a1, true, 1);
        emit_terms(w, p, b0, b1, false, -1);
    } else {
        // leading terms of a, then the combined last term carrying last(a)'s sign
        emit_terms(w, p, a0, a1 - 1, true, 1);
        const uint32_t tl = a1 - 1;
        w.copy(p.pool + p.term_off[tl], (int)(p.term_off[tl + 1] - p.term_off[tl]));
        if (bop == 2) {        // last(a) * first(b)
            w.copy(p.pool + p.term_off[b0], (int)(p.term_off[b0 + 1] - p.term_off[b0]));
            if (p.term_sign[b0] < 0) w.put(PDE_OP_NEG);
            w.put(PDE_OP_MUL);
        } else if (bop == 3) { // last(a) / (whole b)
            emit_terms(w, p, b0, b1, true, 1);
            w.put(PDE_OP_DIV);
        } else {               // last(a) / (1 - t1(b) +- ...)
            w.put(PDE_OP_CONST0);   // CONST(0) = 1
            w.copy(p.pool + p.term_off[b0], (int)(p.term_off[b0 + 1] - p.term_off[b0]));
            if (p.term_sign[b0] < 0) w.put(PDE_OP_NEG);
            w.put(PDE_OP_SUB);
            emit_terms(w, p, b0 + 1, b1, false, 1);
            w.put(PDE_OP_DIV);
        }
        // sign of the combined term / position in the chain
        const int sg = p.term_sign[tl];
        if (a1 - a0 == 1) { if (sg < 0) w.put(PDE_OP_NEG); }
        else w.put(sg > 0 ? PDE_OP_ADD : PDE_OP_SUB);
        if (bop == 2) emit_terms(w, p, b0 + 1, b1, false, 1);
    }
}

__global__ void __launch_bounds__(kEnumThreads)
enum_emit_kernel(const EnumParams p, const long long* block_off, long long first, long long count, int L,
                 int32_t* triple, uint8_t* code, uint8_t* len_out, unsigned long long* hash_out) {
    extern __shared__ __align__(16) uint8_t s_rows[];   // [kEnumThreads][L]
    __shared__ int s_warp[kEnumThreads / 32];
    const long long s = (long long)blockIdx.x * kEnumThreads + threadIdx.x;
    Slot sl;
    sl.keep = false;
    if (s < p.n_slots) sl = decode_slot(p, s);
    int total;
    const int local = block_exclusive_scan(sl.keep ? 1 : 0, s_warp, total);
    const long long base = block_off[blockIdx.x];
    if (total == 0 || base + total <= first || base >= first + count) return;
    {   // zero the block's tile with 16-byte stores (padding + unused rows)
        uint4* z = reinterpret_cast<uint4*>(s_rows);
        const int nz = total * L / 16;
        for (int i = threadIdx.x; i < nz; i += kEnumThreads) z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    if (sl.keep) {
        uint8_t* row = s_rows + (size_t)local * L;
        RowWriter w{row, L < 255 ? L : 255, 0, true};
        if (p.flags[sl.a] || (sl.b >= 0 && p.flags[sl.b])) w.ok = false;
        else splice(w, p, sl);
        int n = w.ok ? w.n : 0;
        if (!w.ok) for (int i = 0; i < L && i < w.n; ++i) row[i] = 0;      // overflowed rows are emitted empty
        const long long c = base + local;
        if (c >= first && c < first + count) {
            const long long o = c - first;
            triple[o * 3 + 0] = sl.op; triple[o * 3 + 1] = sl.a; triple[o * 3 + 2] = sl.b;
            len_out[o] = (uint8_t)n;
            hash_out[o] = hash_row(row, n);
        }
    }
    __syncthreads();
    // coalesced 16-byte copy of the block's tile (rows [lo, hi) of this block)
    const long long lo = max(base, first), hi = min(base + (long long)total, first + count);
    const uint4* src = reinterpret_cast<const uint4*>(s_rows + (size_t)(lo - base) * L);
    uint4* dst = reinterpret_cast<uint4*>(code + (size_t)(lo - first) * L);
    const int nvec = (int)((hi - lo) * L / 16);
    for (int i = threadIdx.x; i < nvec; i += kEnumThreads) dst[i] = src[i];
}

// ------------------------------------------------------------------ dedup
__global__ void __launch_bounds__(256) dedup_insert_kernel(const uint8_t* len, const unsigned long long* hash, long long n,
                                                           unsigned long long* keys, unsigned* vals, unsigned mask) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || len[i] == 0) return;
    unsigned long long h = hash[i];
    if (h == 0) h = 1;
    unsigned slot = (unsigned)(h >> 20) & mask;
    for (;;) {
        const unsigned long long old = atomicCAS(keys + slot, 0ULL, h);
        if (old == 0ULL || old == h) { atomicMin(vals + slot, (unsigned)i); return; }
        slot = (slot + 1) & mask;
    }
}

__global__ void __launch_bounds__(256) dedup_lookup_kernel(const uint8_t* code, const uint8_t* len, const unsigned long long* hash,
                                                           long long n, int L, const unsigned long long* keys, const unsigned* vals,
                                                           unsigned mask, uint8_t* first_occ, unsigned long long* n_unique) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int keep = 0;
    if (i < n) {
        keep = 1;
        if (len[i] != 0) {
            unsigned long long h = hash[i];
            if (h == 0) h = 1;
            unsigned slot = (unsigned)(h >> 20) & mask;
            while (keys[slot] != h) slot = (slot + 1) & mask;
            const long long f = vals[slot];
            if (f != i) {
                // confirm byte-wise: a 64-bit collision must never drop a candidate
                bool same = len[f] == len[i];
                const uint4* x = reinterpret_cast<const uint4*>(code + (size_t)i * L);
                const uint4* y = reinterpret_cast<const uint4*>(code + (size_t)f * L);
                for (int k = 0; same && k < L / 16; ++k) {
                    const uint4 u = x[k], v = y[k];
                    same = (u.x == v.x) && (u.y == v.y) && (u.z == v.z) && (u.w == v.w);
                }
                if (same) keep = 0;
            }
        }
        first_occ[i] = (uint8_t)keep;
    }
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(n_unique, (unsigned long long)__popc(b));
}

// ------------------------------------------------- synthetic trees
__global__ void __launch_bounds__(256) synth_kernel(unsigned long long seed, long long first, long long count, int depth, int L,
                                                    uint8_t* code, uint8_t* len_out, unsigned long long* hash_out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    unsigned long long st = mix64(seed ^ mix64((unsigned long long)(first + t) + 1ULL));
    auto next = [&]() { st += 0x9E3779B97F4A7C15ULL; return mix64(st); };
    uint8_t row[kMaxRow];
    int n = 0;
    // work stack: item < 0x100 = emit byte, else expand depth (item - 0x100)
    unsigned short stack[64];
    int sp = 0;
    stack[sp++] = (unsigned short)(0x100 + depth);
    while (sp > 0) {
        const unsigned it = stack[--sp];
        if (it < 0x100) { if (n < kMaxRow) row[n] = (uint8_t)it; ++n; continue; }
        const int d = (int)it - 0x100;
        if (d <= 1) { const unsigned r = (unsigned)(next() % 5ULL); if (n < kMaxRow) row[n] = kLeaf[r]; ++n; continue; }
        const unsigned r = (unsigned)(next() % 13ULL);
        if (r < 8) {
            stack[sp++] = kUnary[r];
            stack[sp++] = (unsigned short)(0x100 + d - 1);
        } else {
            const int d1 = 1 + (int)(next() % (unsigned long long)(d - 1));
            const unsigned bop = r - 8;
            if (bop == 4) {        // geom_sum: a / (1 - b)
                stack[sp++] = PDE_OP_DIV;
                stack[sp++] = PDE_OP_SUB;
                stack[sp++] = (unsigned short)(0x100 + d - d1);
                stack[sp++] = PDE_OP_CONST0;
                stack[sp++] = (unsigned short)(0x100 + d1);
            } else {
                stack[sp++] = (unsigned short)(PDE_OP_ADD + bop);
                stack[sp++] = (unsigned short)(0x100 + d - d1);
                stack[sp++] = (unsigned short)(0x100 + d1);
            }
        }
    }
    if (n > L || n > 255) n = 0;
    uint8_t* dst = code + (size_t)t * L;
    for (int i = 0; i < L; ++i) { const uint8_t v = i < n ? row[i] : 0; dst[i] = v; if (i >= n && i < kMaxRow) row[i] = 0; }
    len_out[t] = (uint8_t)n;
    {   // same hash as hash_row, byte-assembled (the local row is not 8-byte aligned)
        unsigned long long h = 0x9E3779B97F4A7C15ULL ^ ((unsigned long long)n * 0xD6E8FEB86659FD93ULL);
        const int nw = (n + 7) >> 3;
        for (int w = 0; w < nw; ++w) {
            unsigned long long v = 0;
            for (int k = 7; k >= 0; --k) v = (v << 8) | row[w * 8 + k];
            h = mix64(h ^ v);
        }
        hash_out[t] = h;
    }
}

static int fill_params(const pde_exprset* e, const int32_t* depth_begin, int depth, int prune, EnumParams& p) {
    if (!e || !depth_begin || depth < 2 || depth > kMaxDepth) { set_error("enumerate: bad argument (depth 2..%d)", kMaxDepth); return PDE_E_INVALID; }
    if (int rc = exprset_ensure_device(const_cast<pde_exprset*>(e))) return rc;
    if (depth_begin[0] != 0 || depth_begin[depth - 1] != e->n) { set_error("depth_begin must start at 0 and end at n_expr"); return PDE_E_INVALID; }
    p.flags = e->d_flags; p.attrs = e->d_attrs; p.rank = e->d_rank; p.term_begin = e->d_term_begin;
    p.term_sign = e->d_term_sign; p.term_off = e->d_term_off; p.pool = e->d_pool;
    p.depth = depth; p.prune = prune;
    for (int k = 0; k < depth; ++k) {
        p.depth_begin[k] = depth_begin[k];
        if (k > 0 && depth_begin[k] < depth_begin[k - 1]) { set_error("depth_begin must be non-decreasing"); return PDE_E_INVALID; }
    }
    long long s = 0;
    p.seg_begin[0] = 0;
    s += (long long)(depth_begin[depth - 1] - depth_begin[depth - 2]) * 8;
    for (int d1 = 1; d1 < depth; ++d1) {
        p.seg_begin[d1] = s;
        const int d2 = depth - d1;
        const long long n1 = depth_begin[d1] - depth_begin[d1 - 1], n2 = depth_begin[d2] - depth_begin[d2 - 1];
        if (n1 * n2 * 5 >= 0xffffffffLL) { set_error("candidate index space too large"); return PDE_E_OVERFLOW; }
        s += n1 * n2 * 5;
    }
    p.seg_begin[depth] = s;
    p.n_slots = s;
    if (s / kEnumThreads + 1 > 0x7fffffffLL) { set_error("candidate index space too large"); return PDE_E_OVERFLOW; }
    return PDE_OK;
}

}  // namespace pde

using namespace pde;

// scratch shared by count + emit (per process; one host thread per device)
static unsigned* g_sums = nullptr;
static long long* g_off = nullptr;
static long long* g_total = nullptr;
static long long g_cap = 0;

static int ensure_scratch(long long nblocks) {
    if (nblocks <= g_cap) return PDE_OK;
    cudaFree(g_sums); cudaFree(g_off); cudaFree(g_total);
    g_sums = nullptr; g_off = nullptr; g_total = nullptr; g_cap = 0;
    PDE_CUDA(cudaMalloc(&g_sums, sizeof(unsigned) * nblocks));
    PDE_CUDA(cudaMalloc(&g_off, sizeof(long long) * nblocks));
    PDE_CUDA(cudaMalloc(&g_total, sizeof(long long)));
    g_cap = nblocks;
    return PDE_OK;
}

static int run_count(const EnumParams& p, cudaStream_t st, long long* total_host) {
    const long long nblocks = (p.n_slots + kEnumThreads - 1) / kEnumThreads;
    if (nblocks == 0) { if (total_host) *total_host = 0; return PDE_OK; }
    int rc = ensure_scratch(nblocks);
    if (rc) return rc;
    enum_count_kernel<<<(unsigned)nblocks, kEnumThreads, 0, st>>>(p, g_sums);
    scan_sums_kernel<<<1, 1024, 0, st>>>(g_sums, g_off, (int)nblocks, g_total);
    count_launch(2);
    PDE_CUDA(cudaGetLastError());
    if (total_host) {
        PDE_CUDA(cudaMemcpyAsync(total_host, g_total, sizeof(long long), cudaMemcpyDeviceToHost, st));
        PDE_CUDA(cudaStreamSynchronize(st));
    }
    return PDE_OK;
}

extern "C" {

int pde_enumerate_count(const pde_exprset* e, const int32_t* depth_begin, int depth, int prune,
                        int64_t* n_candidates, void* stream) {
    if (!have_device()) { set_error("no CUDA device: pde_engine_b200 has no CPU fallback"); return PDE_E_NODEVICE; }
    if (!n_candidates) { set_error("null n_candidates"); return PDE_E_INVALID; }
    EnumParams p;
    int rc = fill_params(e, depth_begin, depth, prune, p);
    if (rc) return rc;
    long long total = 0;
    rc = run_count(p, (cudaStream_t)stream, &total);
    if (rc) return rc;
    *n_candidates = total;
    return PDE_OK;
}

int pde_enumerate(const pde_exprset* e, const int32_t* depth_begin, int depth, int prune,
                  int64_t first, int64_t count, int L,
                  int32_t* triple, uint8_t* code, uint8_t* len, uint64_t* hash, void* stream) {
    if (!have_device()) { set_error("no CUDA device: pde_engine_b200 has no CPU fallback"); return PDE_E_NODEVICE; }
    if (!triple || !code || !len || !hash || first < 0 || count < 0) { set_error("pde_enumerate: bad argument"); return PDE_E_INVALID; }
    if (L < 16 || L > kMaxRow || (L % 16) != 0) { set_error("L must be a multiple of 16 in [16, %d]", kMaxRow); return PDE_E_INVALID; }
    EnumParams p;
    int rc = fill_params(e, depth_begin, depth, prune, p);
    if (rc) return rc;
    if (count == 0) return PDE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    rc = run_count(p, st, nullptr);
    if (rc) return rc;
    const long long nblocks = (p.n_slots + kEnumThreads - 1) / kEnumThreads;
    if (nblocks == 0) return PDE_OK;
    const size_t smem = (size_t)kEnumThreads * L;
    PDE_CUDA(cudaFuncSetAttribute(enum_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    enum_emit_kernel<<<(unsigned)nblocks, kEnumThreads, smem, st>>>(p, g_off, first, count, L, triple, code, len,
                                                                    reinterpret_cast<unsigned long long*>(hash));
    count_launch();
    PDE_CUDA(cudaGetLastError());
    return PDE_OK;
}

int pde_dedup(const uint8_t* code, const uint8_t* len, const uint64_t* hash, int64_t n, int L,
              uint8_t* first_occurrence, int64_t* n_unique, void* stream) {
    if (!have_device()) { set_error("no CUDA device: pde_engine_b200 has no CPU fallback"); return PDE_E_NODEVICE; }
    if (!code || !len || !hash || !first_occurrence || n < 0 || (L % 16) != 0) { set_error("pde_dedup: bad argument"); return PDE_E_INVALID; }
    if (n >= 0xffffffffLL) { set_error("pde_dedup: n too large"); return PDE_E_OVERFLOW; }
    if (n == 0) { if (n_unique) *n_unique = 0; return PDE_OK; }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned cap = 1024;
    while ((long long)cap < 2 * n) cap <<= 1;
    unsigned long long* keys = nullptr;
    unsigned* vals = nullptr;
    unsigned long long* cnt = nullptr;
    PDE_CUDA(cudaMalloc(&keys, sizeof(unsigned long long) * cap));
    PDE_CUDA(cudaMalloc(&vals, sizeof(unsigned) * cap));
    PDE_CUDA(cudaMalloc(&cnt, sizeof(unsigned long long)));
    PDE_CUDA(cudaMemsetAsync(keys, 0, sizeof(unsigned long long) * cap, st));
    PDE_CUDA(cudaMemsetAsync(vals, 0xff, sizeof(unsigned) * cap, st));
    PDE_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), st));
    const unsigned blocks = (unsigned)((n + 255) / 256);
    dedup_insert_kernel<<<blocks, 256, 0, st>>>(len, reinterpret_cast<const unsigned long long*>(hash), n, keys, vals, cap - 1);
    dedup_lookup_kernel<<<blocks, 256, 0, st>>>(code, len, reinterpret_cast<const unsigned long long*>(hash), n, L, keys, vals,
                                                cap - 1, first_occurrence, cnt);
    count_launch(2);
    PDE_CUDA(cudaGetLastError());
    unsigned long long c = 0;
    PDE_CUDA(cudaMemcpyAsync(&c, cnt, sizeof(c), cudaMemcpyDeviceToHost, st));
    PDE_CUDA(cudaStreamSynchronize(st));
    if (n_unique) *n_unique = (int64_t)c;
    cudaFree(keys); cudaFree(vals); cudaFree(cnt);
    return PDE_OK;
}

int pde_synth_trees(uint64_t seed, int64_t first, int64_t count, int depth, int L,
                    uint8_t* code, uint8_t* len, uint64_t* hash, void* stream) {
    if (!have_device()) { set_error("no CUDA device: pde_engine_b200 has no CPU fallback"); return PDE_E_NODEVICE; }
    if (!code || !len || !hash || count < 0 || depth < 1 || depth > 8 || L < 16 || L > kMaxRow || (L % 16)) {
        set_error("pde_synth_trees: bad argument"); return PDE_E_INVALID;
    }
    if (count == 0) return PDE_OK;
    const unsigned blocks = (unsigned)((count + 255) / 256);
    synth_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(seed, first, count, depth, L, code, len,
                                                           reinterpret_cast<unsigned long long*>(hash));
    count_launch();
    PDE_CUDA(cudaGetLastError());
    return PDE_OK;
}

}  // extern "C"
