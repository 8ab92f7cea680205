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
// HBM-bound (write only): 69 algorithmic bytes per candidate (L = 48): code L +
// len 1 + hash 8 + triple 12.  Rows are assembled word-wise in shared memory
// from per-expression splice descriptors and written with coalesced 16-byte
// stores; measured 2.9 TB/s = 75 % of the write-only HBM rate (profiles/README.md).
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include <vector>

#include "common.h"

namespace pde {

constexpr int kEnumThreads = 160;     // a block = 160 consecutive slots of the reference's loop nest
constexpr int kMaxDepth = 8;
constexpr int kMaxRow = 256;

// Per-expression splice descriptor (built on the host next to the device mirror): every op of the
// reference's textual templates (LBF:170-195) is a concatenation of at most 8 pieces, each a contiguous
// range of the operand's WHOLE program  t1 [NEG] (tk ADD|SUB)*  or a literal byte:
//   x = offset of the whole program in wpool
//   y = whole length | first body length << 8 | offset of the last body << 16 | flags << 24
enum : unsigned { D_FIRST_NEG = 1, D_LAST_NEG = 2, D_MULTI = 4, D_BAD = 8 };

struct EnumParams {
    const uint2* desc;
    const uint8_t* wpool;      // padded by 8 bytes
    const uint8_t* attrs;
    const uint32_t* rank;
    int depth;
    int prune;
    int depth_begin[kMaxDepth + 1];
    // segment 0 = unary (160 consecutive slots per block), segment d1 >= 1 = binary (32 pairs x 5 ops per block)
    long long seg_block[kMaxDepth + 1];    // first block of the segment; seg_block[depth] = number of blocks
    long long seg_groups[kMaxDepth + 1];   // unary: slots; binary: pairs
    long long block0;                      // first block of this launch (windowed passes launch only the tiles they need)
};

// enumerator unary op index -> opcode (iteration order of UNARY_OPS, expression_operations.py:80-89)
__device__ const uint8_t kUnary[8] = {PDE_OP_FN_NEG, PDE_OP_FN_INV, PDE_OP_SQRT, PDE_OP_FN_SQUARE,
                                      PDE_OP_FN_POW32, PDE_OP_FN_POWN32, PDE_OP_EXP, PDE_OP_FN_EXPNEG};
// synthetic leaves: rho, z, PRIM(0) = rho**2 + z**2, PRIM(1) = rho/z, 1  (problems/__init__.py:73-79)
__device__ const uint8_t kLeaf[5] = {PDE_OP_VAR0, PDE_OP_VAR1, PDE_OP_PRIM0, PDE_OP_PRIM0 + 1, PDE_OP_CONST0};

struct Slot {
    int op;       // 0..7 unary, 8..12 binary
    int a, b;     // global operand indices after the add/mul swap (b = -1 for unary)
    int local;    // position of the slot in the block, in the reference's order
    bool keep;
};

// Thread -> slot.  Unary blocks: thread t = slot t (every unary op takes the same splice path).  Binary
// blocks: lane = pair, warp = op, so a warp executes ONE splice path (the v1 mapping, thread = slot, had
// five paths per warp: 9.4 of 32 threads active per instruction, profiles/README.md).
__device__ __forceinline__ Slot decode_slot(const EnumParams& p) {
    Slot r;
    r.keep = false; r.op = 0; r.a = 0; r.b = -1;
    const int d = p.depth;
    const long long blk = p.block0 + blockIdx.x;
    if (blk < p.seg_block[1]) {
        // unary, LBF:142-153
        r.local = threadIdx.x;
        const long long s = blk * kEnumThreads + threadIdx.x;
        if (s >= p.seg_groups[0]) return r;
        const int e = (int)(s >> 3);
        r.op = (int)(s & 7);
        r.a = p.depth_begin[d - 2] + e;
        r.keep = true;
        if (p.prune) {
            const unsigned at = p.attrs[r.a];
            if (!(at & PDE_ATTR_HAS_VARS)) r.keep = false;
            if (r.op == 1 && (at & PDE_ATTR_STARTS_INV)) r.keep = false;
            if (r.op >= 2 && r.op <= 5 && (at & PDE_ATTR_IS_ONE)) r.keep = false;
        }
        return r;
    }
    int d1 = 1;
    while (d1 < d - 1 && blk >= p.seg_block[d1 + 1]) ++d1;
    const int lane = threadIdx.x & 31, bop = threadIdx.x >> 5;
    r.local = lane * 5 + bop;
    const long long pair = (blk - p.seg_block[d1]) * 32 + lane;
    if (pair >= p.seg_groups[d1]) return r;
    const int d2 = d - d1;
    const unsigned n2 = (unsigned)(p.depth_begin[d2] - p.depth_begin[d2 - 1]);
    const unsigned i1 = (unsigned)pair / n2;                       // pairs < 2^32 (checked on the host)
    int a = p.depth_begin[d1 - 1] + (int)i1;
    int b = p.depth_begin[d2 - 1] + (int)((unsigned)pair - i1 * n2);
    const unsigned ata = p.attrs[a], atb = p.attrs[b];
    r.keep = true;
    if (p.prune && !((ata | atb) & PDE_ATTR_HAS_VARS)) r.keep = false;
    const uint32_t ra = p.rank[a], rb = p.rank[b];
    bool one_a = ata & PDE_ATTR_IS_ONE, one_b = atb & PDE_ATTR_IS_ONE;
    if ((bop == 0 || bop == 2) && ra > rb) {                      // LBF:168-169
        const int t = a; a = b; b = t;
        const bool tb = one_a; one_a = one_b; one_b = tb;
    }
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

// exclusive prefix of the keep flags in SLOT order (slot_local), block total in `total`
__device__ __forceinline__ int block_exclusive_scan(int keep, int slot_local, int* s_flag, int* s_warp, int& total) {
    s_flag[slot_local] = keep;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int v = s_flag[threadIdx.x];
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
    s_flag[threadIdx.x] = base + x - v;
    __syncthreads();
    return s_flag[slot_local];
}

__global__ void __launch_bounds__(kEnumThreads) enum_count_kernel(const EnumParams p, unsigned* block_sums) {
    const Slot sl = decode_slot(p);
    const int n = __syncthreads_count(sl.keep ? 1 : 0);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = (unsigned)n;
}

// Two-level exclusive scan of the per-block candidate counts (the first version scanned all ~10^5 counts
// in ONE block with strided accesses: 93 us of a 0.68 ms pass).
//   scan_tiles_kernel: tile t = 1024 consecutive counts: in-tile exclusive prefix (u32) + tile total
//   scan_super_kernel: exclusive prefix of the tile totals (one block), grand total
// A block's base = tile_off[b / 1024] + in_tile[b].
constexpr int kScanTile = 1024;

__global__ void __launch_bounds__(kScanTile) scan_tiles_kernel(const unsigned* sums, unsigned* in_tile, unsigned long long* tile_total, int nblocks) {
    __shared__ unsigned s_w[kScanTile / 32];
    const int i = blockIdx.x * kScanTile + threadIdx.x;
    const unsigned v = i < nblocks ? sums[i] : 0u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned x = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x += y;
    }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    if (warp == 0) {
        unsigned t = s_w[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, t, off);
            if (lane >= off) t += y;
        }
        s_w[lane] = t;            // inclusive prefix of the warp totals
    }
    __syncthreads();
    const unsigned base = warp ? s_w[warp - 1] : 0u;
    if (i < nblocks) in_tile[i] = base + x - v;
    if (threadIdx.x == kScanTile - 1) tile_total[blockIdx.x] = (unsigned long long)base + x;
}

__global__ void __launch_bounds__(1024) scan_super_kernel(unsigned long long* tile_total, int ntiles, long long* total) {
    // ntiles is small (blocks / 1024): serial chunks per thread, then a serial pass over 1024 partials
    __shared__ unsigned long long s_part[1024];
    const int per = (ntiles + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = min(lo + per, ntiles);
    unsigned long long acc = 0;
    for (int i = lo; i < hi; ++i) acc += tile_total[i];
    s_part[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < 1024; ++i) { const unsigned long long c = s_part[i]; s_part[i] = run; run += c; }
        *total = (long long)run;
    }
    __syncthreads();
    unsigned long long run = s_part[threadIdx.x];
    for (int i = lo; i < hi; ++i) { const unsigned long long c = tile_total[i]; tile_total[i] = run; run += c; }
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

// Row assembly in 32-bit words: pieces are appended through a 64-bit shift register and flushed a word at
// a time; source bytes come from two aligned loads + a funnel shift (the v1 writer moved single bytes with
// a bounds check each: ~6 instructions per byte).
struct RowWriter {
    uint32_t* dst;            // shared-memory row, cap words
    int cap;
    unsigned long long acc;   // pending bytes, low k
    int k, nw, n;
    __device__ __forceinline__ void flush() {
        if (nw < cap) dst[nw] = (uint32_t)acc;
        ++nw; acc >>= 32; k -= 4;
    }
    __device__ __forceinline__ void put(unsigned b) {
        acc |= (unsigned long long)b << (8 * k);
        ++k; ++n;
        if (k == 4) flush();
    }
    __device__ __forceinline__ void copy(const uint32_t* pool32, unsigned src, int len) {
        n += len;
        while (len > 0) {
            const unsigned al = src >> 2, sh = (src & 3u) * 8u;
            uint32_t v = __funnelshift_r(__ldg(pool32 + al), __ldg(pool32 + al + 1), sh);   // 4 bytes from src
            const int m = len < 4 ? len : 4;
            if (m < 4) v &= (1u << (8 * m)) - 1u;
            acc |= (unsigned long long)v << (8 * k);
            k += m;
            if (k >= 4) flush();
            src += (unsigned)m; len -= m;
        }
    }
    __device__ __forceinline__ void finish() {       // last partial word (the tile is pre-zeroed)
        if (k > 0) { if (nw < cap) dst[nw] = (uint32_t)acc; ++nw; }
    }
};

struct Desc {
    unsigned off, len, first_len, last_off, fl;
    __device__ __forceinline__ explicit Desc(const uint2 d)
        : off(d.x), len(d.y & 0xffu), first_len((d.y >> 8) & 0xffu), last_off((d.y >> 16) & 0xffu), fl(d.y >> 24) {}
    // bytes of `whole` taken by the first term in fresh mode: body + optional NEG
    __device__ __forceinline__ unsigned skip() const { return first_len + ((fl & D_FIRST_NEG) ? 1u : 0u); }
    __device__ __forceinline__ unsigned last_len() const {
        return len - last_off - ((fl & D_MULTI) ? 1u : ((fl & D_FIRST_NEG) ? 1u : 0u));
    }
};

// The reference's templates on whole programs (see the file header):
//   unary  whole(a) op
//   add    whole(a) body1(b) (ADD|SUB by sign1(b)) rest(b)          sub: sign flipped
//   mul    lead(a) last(a) body1(b) [NEG] MUL sign(last a) rest(b)
//   div    lead(a) last(a) whole(b) DIV sign(last a)
//   geom   lead(a) last(a) 1 body1(b) [NEG] SUB rest(b) DIV sign(last a)
// lead(a) = whole(a) up to its last body; rest(b) = whole(b) after its first term; sign(last a) = ADD|SUB
// for a multi-term a, NEG for a negative single term.
__device__ __forceinline__ void splice(RowWriter& w, const uint32_t* pool32, int op, const Desc& A, const Desc& B) {
    if (op < 8) {
        w.copy(pool32, A.off, (int)A.len);
        w.put(kUnary[op]);
        return;
    }
    const int bop = op - 8;
    const unsigned skip = B.skip();
    const bool bneg = B.fl & D_FIRST_NEG;
    if (bop <= 1) {
        w.copy(pool32, A.off, (int)A.len);
        w.copy(pool32, B.off, (int)B.first_len);
        w.put((bneg != (bop == 1)) ? PDE_OP_SUB : PDE_OP_ADD);
        w.copy(pool32, B.off + skip, (int)(B.len - skip));
        return;
    }
    w.copy(pool32, A.off, (int)(A.last_off + A.last_len()));          // lead(a) ++ last(a): contiguous
    if (bop == 2) {
        w.copy(pool32, B.off, (int)B.first_len);
        if (bneg) w.put(PDE_OP_NEG);
        w.put(PDE_OP_MUL);
    } else if (bop == 3) {
        w.copy(pool32, B.off, (int)B.len);
        w.put(PDE_OP_DIV);
    } else {
        w.put(PDE_OP_CONST0);   // CONST(0) = 1
        w.copy(pool32, B.off, (int)B.first_len);
        if (bneg) w.put(PDE_OP_NEG);
        w.put(PDE_OP_SUB);
        w.copy(pool32, B.off + skip, (int)(B.len - skip));
        w.put(PDE_OP_DIV);
    }
    if (A.fl & D_MULTI) w.put((A.fl & D_LAST_NEG) ? PDE_OP_SUB : PDE_OP_ADD);
    else if (A.fl & D_FIRST_NEG) w.put(PDE_OP_NEG);
    if (bop == 2) w.copy(pool32, B.off + skip, (int)(B.len - skip));
}

// ---- CSR form of the output: programs back to back in one byte pool, each padded to 16 bytes (an empty program
// takes none), offset[c] in 16-byte units.  Mean program 15.4 bytes: ~49 B per candidate instead of 149 at L = 128.
__device__ __forceinline__ int csr_row_bytes(int n) { return (n + 15) & ~15; }

// Length of what splice() writes, from the descriptors alone (the CSR count pass sizes the byte pool with it; the emit
// pass checks it against the writer).
__device__ __forceinline__ int splice_len(int op, const Desc& A, const Desc& B) {
    if (op < 8) return (int)A.len + 1;
    const int bop = op - 8;
    const int skip = (int)B.skip(), rest = (int)B.len - skip, bneg = (B.fl & D_FIRST_NEG) ? 1 : 0;
    if (bop <= 1) return (int)A.len + (int)B.first_len + 1 + rest;
    const int head = (int)(A.last_off + A.last_len());
    const int sign = (A.fl & D_MULTI) ? 1 : ((A.fl & D_FIRST_NEG) ? 1 : 0);
    if (bop == 2) return head + (int)B.first_len + bneg + 1 + sign + rest;
    if (bop == 3) return head + (int)B.len + 1 + sign;
    return head + 1 + (int)B.first_len + bneg + 1 + rest + 1 + sign;
}

__device__ __forceinline__ int slot_program_len(const EnumParams& p, const Slot& sl, int L) {
    if (!sl.keep) return 0;
    const Desc A(__ldg(p.desc + sl.a));
    const Desc B(sl.b >= 0 ? __ldg(p.desc + sl.b) : make_uint2(0u, 0u));
    if ((A.fl | B.fl) & D_BAD) return 0;
    const int n = splice_len(sl.op, A, B);
    return (n > L || n > 255) ? 0 : n;
}

__global__ void __launch_bounds__(kEnumThreads) enum_count_bytes_kernel(const EnumParams p, int L, unsigned* block_bytes) {
    __shared__ unsigned s_sum;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    const Slot sl = decode_slot(p);
    unsigned b = (unsigned)csr_row_bytes(slot_program_len(p, sl, L));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) b += __shfl_xor_sync(0xffffffffu, b, off);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&s_sum, b);
    __syncthreads();
    if (threadIdx.x == 0) block_bytes[blockIdx.x] = s_sum;
}

// Shared-memory rows are L + 16 bytes apart: with a stride of L = 128 bytes every lane's row starts in the
// same bank and each word store of the row owners was a 32-way conflict (mio_throttle + short_scoreboard
// were the top stalls, profiles/README.md); L + 16 keeps rows 16-byte aligned for the vector copy-out.
__host__ __device__ constexpr int enum_row_stride(int L) { return L + 16; }

__global__ void __launch_bounds__(kEnumThreads)
enum_emit_kernel(const EnumParams p, const unsigned* sums, const unsigned* in_tile, const unsigned long long* tile_off, long long first, long long count, int L,
                 int32_t* triple, uint8_t* code, uint8_t* len_out, unsigned long long* hash_out) {
    // A block outside the window [first, first + count) leaves before it touches shared memory: the count pass
    // already knows every block's candidate range, so an 8-way sharded pass costs each rank 1/8 of the work
    // (+ ~10 us of empty blocks), not a full pass.
    const long long blk = p.block0 + blockIdx.x;
    const long long base = (long long)(tile_off[blk / kScanTile] + in_tile[blk]);
    {
        const long long n_blk = (long long)sums[blk];
        if (n_blk == 0 || base + n_blk <= first || base >= first + count) return;
    }
    extern __shared__ __align__(16) uint8_t s_rows[];   // [kEnumThreads][L + 16] rows, then per-row hash / triple / len
    __shared__ int s_flag[kEnumThreads];
    __shared__ int s_warp[kEnumThreads / 32];
    const int Ls = enum_row_stride(L);
    unsigned long long* s_hash = reinterpret_cast<unsigned long long*>(s_rows + (size_t)kEnumThreads * Ls);
    int* s_triple = reinterpret_cast<int*>(s_hash + kEnumThreads);
    uint8_t* s_len = reinterpret_cast<uint8_t*>(s_triple + 3 * kEnumThreads);
    {   // zero the tile with 16-byte stores (padding of every row); ordered before the splice by the scan's barriers
        uint4* z = reinterpret_cast<uint4*>(s_rows);
        const int nz = kEnumThreads * Ls / 16;
        for (int i = threadIdx.x; i < nz; i += kEnumThreads) z[i] = make_uint4(0, 0, 0, 0);
    }
    const Slot sl = decode_slot(p);
    int total;
    const int local = block_exclusive_scan(sl.keep ? 1 : 0, sl.local, s_flag, s_warp, total);
    if (sl.keep) {
        uint8_t* row = s_rows + (size_t)local * Ls;
        RowWriter w{reinterpret_cast<uint32_t*>(row), L / 4, 0ULL, 0, 0, 0};
        const Desc A(__ldg(p.desc + sl.a));
        const Desc B(sl.b >= 0 ? __ldg(p.desc + sl.b) : make_uint2(0u, 0u));
        const bool bad = (A.fl | B.fl) & D_BAD;
        if (!bad) splice(w, reinterpret_cast<const uint32_t*>(p.wpool), sl.op, A, B);
        int n = w.n;
        w.finish();
        if (bad || n > L || n > 255) {                                   // not compilable / too long: emitted empty
            n = 0;
            for (int i = 0; i < L / 4 && i < w.nw; ++i) reinterpret_cast<uint32_t*>(row)[i] = 0u;
        }
        s_len[local] = (uint8_t)n;
        s_triple[3 * local + 0] = sl.op; s_triple[3 * local + 1] = sl.a; s_triple[3 * local + 2] = sl.b;
        s_hash[local] = hash_row(row, n);
    }
    __syncthreads();
    // coalesced copy-out of the block's rows [lo, hi) (the window may cut the block)
    const long long lo = max(base, first), hi = min(base + (long long)total, first + count);
    const int r0 = (int)(lo - base), nr = (int)(hi - lo);
    const long long o0 = lo - first;
    {
        // thread t copies 16-byte vectors t, t + 160, ...: (row, column) advance incrementally (no division
        // per vector: i / (L / 16) was the hottest line of the kernel)
        uint4* dst = reinterpret_cast<uint4*>(code + (size_t)o0 * L);
        const int vpr = L / 16, nvec = nr * vpr;
        const int dr = kEnumThreads / vpr, dc = kEnumThreads - dr * vpr;
        int r = (int)threadIdx.x / vpr, c = (int)threadIdx.x - r * vpr;
        const uint8_t* src = s_rows + (size_t)r0 * Ls;
        for (int i = threadIdx.x; i < nvec; i += kEnumThreads) {
            dst[i] = *reinterpret_cast<const uint4*>(src + r * Ls + c * 16);
            r += dr; c += dc;
            if (c >= vpr) { c -= vpr; ++r; }
        }
    }
    {
        unsigned long long* ho = hash_out + o0;
        uint8_t* lo8 = len_out + o0;
        int* to = triple + o0 * 3;
        const unsigned long long* hs = s_hash + r0;
        const uint8_t* ls = s_len + r0;
        const int* ts = s_triple + 3 * r0;
        for (int i = threadIdx.x; i < nr; i += kEnumThreads) { ho[i] = hs[i]; lo8[i] = ls[i]; }
        for (int i = threadIdx.x; i < 3 * nr; i += kEnumThreads) to[i] = ts[i];
    }
}

// CSR emit: the same block / slot mapping, rows assembled in shared memory as above; the copy-out packs them.
//   byte base of the block   = bytes_tile[blk / 1024] + bytes_in_tile[blk]     (scan of enum_count_bytes_kernel)
//   offset of row r          = base + exclusive scan of the padded lengths inside the block
// offset_out[c] is in 16-byte units relative to `pool_lo` (the byte base of the first block of the launch); the last
// candidate of the window also writes offset_out[count] (the end), so a row's padded size is off[c + 1] - off[c].
struct CsrParams {
    const unsigned* sums; const unsigned* in_tile; const unsigned long long* tile_off;              // candidates
    const unsigned* bytes_in_tile; const unsigned long long* bytes_tile;                            // bytes
    long long first, count;
    unsigned long long pool_lo;
    int L;
    int32_t* triple; uint8_t* pool; unsigned* offset; uint8_t* len_out; unsigned long long* hash_out;
};

__global__ void __launch_bounds__(kEnumThreads) enum_emit_csr_kernel(const EnumParams p, const CsrParams c) {
    const long long blk = p.block0 + blockIdx.x;
    const long long base = (long long)(c.tile_off[blk / kScanTile] + c.in_tile[blk]);
    {
        const long long n_blk = (long long)c.sums[blk];
        if (n_blk == 0 || base + n_blk <= c.first || base >= c.first + c.count) return;
    }
    // The lengths are known from the descriptors BEFORE anything is spliced (splice_len), so ONE scan -- candidate
    // count in the high bits, padded bytes in the low bits -- gives every kept slot its row index and its byte offset
    // in the block's packed tile; the rows are spliced straight into place (no row stride, no zero fill, no second
    // scan) and the tile leaves the block as one linear, coalesced copy.
    extern __shared__ __align__(16) uint8_t s_tile[];          // [<= kEnumThreads * L] packed rows, then per-row hash / triple / len / offset
    __shared__ int s_flag[kEnumThreads];
    __shared__ int s_warp[kEnumThreads / 32];
    const int L = c.L;
    unsigned long long* s_hash = reinterpret_cast<unsigned long long*>(s_tile + (size_t)kEnumThreads * L);
    int* s_triple = reinterpret_cast<int*>(s_hash + kEnumThreads);
    int* s_boff = s_triple + 3 * kEnumThreads;                 // [kEnumThreads + 1]
    uint8_t* s_len = reinterpret_cast<uint8_t*>(s_boff + kEnumThreads + 1);
    const Slot sl = decode_slot(p);
    int n = 0;
    Desc A(make_uint2(0u, 0u)), B(make_uint2(0u, 0u));
    if (sl.keep) {
        A = Desc(__ldg(p.desc + sl.a));
        if (sl.b >= 0) B = Desc(__ldg(p.desc + sl.b));
        if (!((A.fl | B.fl) & D_BAD)) {
            n = splice_len(sl.op, A, B);
            if (n > L || n > 255) n = 0;                        // not compilable / too long: emitted empty
        }
    }
    const int pb = csr_row_bytes(n);
    int packed_total;
    const int packed = block_exclusive_scan(sl.keep ? ((1 << 20) | pb) : 0, sl.local, s_flag, s_warp, packed_total);
    const int local = packed >> 20, boff = packed & 0xfffff, total = packed_total >> 20;
    if (sl.keep) {
        uint8_t* row = s_tile + boff;
        if (n > 0) {
            RowWriter w{reinterpret_cast<uint32_t*>(row), pb / 4, 0ULL, 0, 0, 0};
            splice(w, reinterpret_cast<const uint32_t*>(p.wpool), sl.op, A, B);
            w.finish();
            for (int i = (n + 3) >> 2; i < (pb >> 2); ++i) reinterpret_cast<uint32_t*>(row)[i] = 0u;       // pad to 16 bytes
        }
        s_len[local] = (uint8_t)n;
        s_boff[local] = boff;
        s_triple[3 * local + 0] = sl.op; s_triple[3 * local + 1] = sl.a; s_triple[3 * local + 2] = sl.b;
        s_hash[local] = hash_row(row, n);
    }
    if (threadIdx.x == 0) s_boff[total] = packed_total & 0xfffff;
    __syncthreads();
    const unsigned long long bbase = c.bytes_tile[blk / kScanTile] + c.bytes_in_tile[blk] - c.pool_lo;     // relative to the pool
    const long long lo = max(base, c.first), hi = min(base + (long long)total, c.first + c.count);
    const int r0 = (int)(lo - base), nr = (int)(hi - lo);
    const long long o0 = lo - c.first;
    {   // the window's part of the packed tile: one linear copy
        const int v0 = s_boff[r0] >> 4, v1 = s_boff[r0 + nr] >> 4;
        const uint4* src = reinterpret_cast<const uint4*>(s_tile);
        uint4* dst = reinterpret_cast<uint4*>(c.pool + bbase);
        for (int v = v0 + (int)threadIdx.x; v < v1; v += kEnumThreads) dst[v] = src[v];
    }
    {
        unsigned long long* ho = c.hash_out + o0;
        uint8_t* lo8 = c.len_out + o0;
        unsigned* oo = c.offset + o0;
        int* to = c.triple + o0 * 3;
        for (int i = threadIdx.x; i < nr; i += kEnumThreads) {
            ho[i] = s_hash[r0 + i]; lo8[i] = s_len[r0 + i];
            oo[i] = (unsigned)((bbase + (unsigned long long)s_boff[r0 + i]) >> 4);
        }
        if (threadIdx.x == 0 && hi == c.first + c.count) oo[nr] = (unsigned)((bbase + (unsigned long long)s_boff[r0 + nr]) >> 4);
        for (int i = threadIdx.x; i < 3 * nr; i += kEnumThreads) to[i] = s_triple[3 * r0 + i];
    }
}

// ------------------------------------------------------------------ dedup
// CSR rows: candidate i's program = pool + 16 * off[i], len[i] bytes (padding zero)
__global__ void __launch_bounds__(256) dedup_lookup_csr_kernel(const uint8_t* pool, const unsigned* off, const uint8_t* len, const unsigned long long* hash,
                                                               long long n, const unsigned long long* keys, const unsigned* vals,
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
                bool same = len[f] == len[i];
                const uint4* x = reinterpret_cast<const uint4*>(pool + (size_t)off[i] * 16);
                const uint4* y = reinterpret_cast<const uint4*>(pool + (size_t)off[f] * 16);
                const int nv = ((int)len[i] + 15) >> 4;
                for (int k = 0; same && k < nv; ++k) {
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
    p.desc = reinterpret_cast<const uint2*>(e->d_desc); p.wpool = e->d_wpool; p.attrs = e->d_attrs; p.rank = e->d_rank;
    p.depth = depth; p.prune = prune; p.block0 = 0;
    for (int k = 0; k < depth; ++k) {
        p.depth_begin[k] = depth_begin[k];
        if (k > 0 && depth_begin[k] < depth_begin[k - 1]) { set_error("depth_begin must be non-decreasing"); return PDE_E_INVALID; }
    }
    long long blk = 0;
    p.seg_block[0] = 0;
    p.seg_groups[0] = (long long)(depth_begin[depth - 1] - depth_begin[depth - 2]) * 8;
    blk += (p.seg_groups[0] + kEnumThreads - 1) / kEnumThreads;
    for (int d1 = 1; d1 < depth; ++d1) {
        p.seg_block[d1] = blk;
        const int d2 = depth - d1;
        const long long n1 = depth_begin[d1] - depth_begin[d1 - 1], n2 = depth_begin[d2] - depth_begin[d2 - 1];
        if (n1 * n2 >= 0xffffffffLL) { set_error("candidate index space too large"); return PDE_E_OVERFLOW; }
        p.seg_groups[d1] = n1 * n2;
        blk += (n1 * n2 + 31) / 32;
    }
    p.seg_block[depth] = blk;
    if (blk + 1 > 0x7fffffffLL) { set_error("candidate index space too large"); return PDE_E_OVERFLOW; }
    return PDE_OK;
}

}  // namespace pde

using namespace pde;

// The count pass (candidates per block + their two-level exclusive scan) is cached ON THE HANDLE: its buffers belong to
// the exprset, are valid for one (depth, prune, depth_begin) on one device, and are freed with it -- no file-scope state,
// and a windowed pde_enumerate (one window per rank or per chunk) does not repeat the pass.
static int current_device() { int d = 0; cudaGetDevice(&d); return d; }

static bool count_cached(const pde_exprset* e, const EnumParams& p) {
    if (!e->d_count_sums || e->count_depth != p.depth || e->count_prune != p.prune || e->device != current_device()) return false;
    for (int k = 0; k < p.depth; ++k) if (e->count_db[k] != p.depth_begin[k]) return false;
    return e->count_blocks == p.seg_block[p.depth];
}

static int run_count(pde_exprset* e, const EnumParams& p, cudaStream_t st, long long* total_host) {
    const long long nblocks = p.seg_block[p.depth];
    if (nblocks == 0) { if (total_host) *total_host = 0; return PDE_OK; }
    if (count_cached(e, p)) { if (total_host) *total_host = e->count_total; return PDE_OK; }
    scratch_free(e->d_count_sums, st); scratch_free(e->d_count_in_tile, st); scratch_free(e->d_count_tile, st);
    e->d_count_sums = nullptr; e->d_count_in_tile = nullptr; e->d_count_tile = nullptr;
    e->count_bytes_L = 0;                  // the CSR byte counts belong to the same (depth, prune, depth_begin)
    const int ntiles = (int)((nblocks + kScanTile - 1) / kScanTile);
    int rc;
    if ((rc = scratch_alloc(reinterpret_cast<void**>(&e->d_count_sums), sizeof(unsigned) * nblocks, st))) return rc;
    if ((rc = scratch_alloc(reinterpret_cast<void**>(&e->d_count_in_tile), sizeof(unsigned) * nblocks, st))) return rc;
    if ((rc = scratch_alloc(reinterpret_cast<void**>(&e->d_count_tile), sizeof(unsigned long long) * (ntiles + 1), st))) return rc;
    long long* d_total = nullptr;
    rc = scratch_alloc(reinterpret_cast<void**>(&d_total), sizeof(long long), st);
    if (rc) return rc;
    enum_count_kernel<<<(unsigned)nblocks, kEnumThreads, 0, st>>>(p, e->d_count_sums);
    scan_tiles_kernel<<<ntiles, kScanTile, 0, st>>>(e->d_count_sums, e->d_count_in_tile, e->d_count_tile, (int)nblocks);
    scan_super_kernel<<<1, 1024, 0, st>>>(e->d_count_tile, ntiles, d_total);
    count_launch(3);
    PDE_CUDA(cudaGetLastError());
    long long total = 0;
    e->count_tile_host.assign(ntiles, 0ULL);
    PDE_CUDA(cudaMemcpyAsync(&total, d_total, sizeof(long long), cudaMemcpyDeviceToHost, st));
    PDE_CUDA(cudaMemcpyAsync(e->count_tile_host.data(), e->d_count_tile, sizeof(unsigned long long) * ntiles, cudaMemcpyDeviceToHost, st));
    PDE_CUDA(cudaStreamSynchronize(st));
    scratch_free(d_total, st);
    e->count_depth = p.depth; e->count_prune = p.prune; e->count_blocks = nblocks; e->count_total = total;
    for (int k = 0; k < p.depth; ++k) e->count_db[k] = p.depth_begin[k];
    if (total_host) *total_host = total;
    return PDE_OK;
}

// CSR: bytes per block (programs padded to 16 bytes) for row length L, their scan, and the host copies of the
// per-block prefixes.  Cached on the handle next to the candidate counts.
static int run_count_bytes(pde_exprset* e, const EnumParams& p, int L, cudaStream_t st) {
    const long long nblocks = p.seg_block[p.depth];
    if (nblocks == 0 || (e->count_bytes_L == L && e->d_bytes_sums)) return PDE_OK;
    scratch_free(e->d_bytes_sums, st); scratch_free(e->d_bytes_in_tile, st); scratch_free(e->d_bytes_tile, st);
    e->d_bytes_sums = nullptr; e->d_bytes_in_tile = nullptr; e->d_bytes_tile = nullptr;
    const int ntiles = (int)((nblocks + kScanTile - 1) / kScanTile);
    int rc;
    if ((rc = scratch_alloc(reinterpret_cast<void**>(&e->d_bytes_sums), sizeof(unsigned) * nblocks, st))) return rc;
    if ((rc = scratch_alloc(reinterpret_cast<void**>(&e->d_bytes_in_tile), sizeof(unsigned) * nblocks, st))) return rc;
    if ((rc = scratch_alloc(reinterpret_cast<void**>(&e->d_bytes_tile), sizeof(unsigned long long) * (ntiles + 1), st))) return rc;
    long long* d_total = nullptr;
    rc = scratch_alloc(reinterpret_cast<void**>(&d_total), sizeof(long long), st);
    if (rc) return rc;
    enum_count_bytes_kernel<<<(unsigned)nblocks, kEnumThreads, 0, st>>>(p, L, e->d_bytes_sums);
    scan_tiles_kernel<<<ntiles, kScanTile, 0, st>>>(e->d_bytes_sums, e->d_bytes_in_tile, e->d_bytes_tile, (int)nblocks);
    scan_super_kernel<<<1, 1024, 0, st>>>(e->d_bytes_tile, ntiles, d_total);
    count_launch(3);
    PDE_CUDA(cudaGetLastError());
    // host prefixes per block: candidates and bytes
    std::vector<unsigned> cs(nblocks), bs(nblocks);
    PDE_CUDA(cudaMemcpyAsync(cs.data(), e->d_count_sums, sizeof(unsigned) * nblocks, cudaMemcpyDeviceToHost, st));
    PDE_CUDA(cudaMemcpyAsync(bs.data(), e->d_bytes_sums, sizeof(unsigned) * nblocks, cudaMemcpyDeviceToHost, st));
    PDE_CUDA(cudaStreamSynchronize(st));
    scratch_free(d_total, st);
    e->block_cand_host.assign(nblocks + 1, 0ULL);
    e->block_bytes_host.assign(nblocks + 1, 0ULL);
    for (long long b = 0; b < nblocks; ++b) {
        e->block_cand_host[b + 1] = e->block_cand_host[b] + cs[b];
        e->block_bytes_host[b + 1] = e->block_bytes_host[b] + bs[b];
    }
    e->count_bytes_L = L;
    return PDE_OK;
}

// blocks [b0, b1) that hold the candidates [first, first + count)
static void window_blocks(const pde_exprset* e, long long first, long long count, long long& b0, long long& b1) {
    const std::vector<unsigned long long>& cp = e->block_cand_host;
    b0 = (long long)(std::upper_bound(cp.begin(), cp.end(), (unsigned long long)first) - cp.begin()) - 1;
    b1 = (long long)(std::lower_bound(cp.begin(), cp.end(), (unsigned long long)(first + count)) - cp.begin());
    if (b0 < 0) b0 = 0;
    if (b1 > (long long)cp.size() - 1) b1 = (long long)cp.size() - 1;
    if (b1 < b0) b1 = b0;
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
    rc = run_count(const_cast<pde_exprset*>(e), p, (cudaStream_t)stream, &total);
    if (rc) return rc;
    *n_candidates = total;
    return PDE_OK;
}

int pde_enumerate(const pde_exprset* e, const int32_t* depth_begin, int depth, int prune,
                  int64_t first, int64_t count, int L,
                  int32_t* triple, uint8_t* code, uint8_t* len, uint64_t* hash, void* stream) {
    if (!have_device()) { set_error("no CUDA device: pde_engine_b200 has no CPU fallback"); return PDE_E_NODEVICE; }
    if (first < 0 || count < 0) { set_error("pde_enumerate: bad argument"); return PDE_E_INVALID; }
    if (L < 16 || L > kMaxRow || (L % 16) != 0) { set_error("L must be a multiple of 16 in [16, %d]", kMaxRow); return PDE_E_INVALID; }
    EnumParams p;
    int rc = fill_params(e, depth_begin, depth, prune, p);
    if (rc) return rc;
    if (count == 0) return PDE_OK;           // an empty window is a no-op (its buffers may be null: the reference just moves on, LBF:139-195)
    if (!triple || !code || !len || !hash) { set_error("pde_enumerate: null output"); return PDE_E_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    long long total = 0;
    rc = run_count(const_cast<pde_exprset*>(e), p, st, &total);
    if (rc) return rc;
    const long long nblocks = p.seg_block[p.depth];
    if (nblocks == 0) return PDE_OK;
    if (first + count > total) { set_error("pde_enumerate: window [%lld, %lld) exceeds the %lld candidates", (long long)first, (long long)(first + count), total); return PDE_E_INVALID; }
    const size_t smem = (size_t)kEnumThreads * (enum_row_stride(L) + 8 + 12 + 1) + 16;
    PDE_CUDA(cudaFuncSetAttribute(enum_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // launch only the 1024-block tiles whose candidate range meets the window (the host keeps the tile prefix)
    const std::vector<unsigned long long>& tp = e->count_tile_host;
    long long t0 = 0, t1 = (long long)tp.size() - 1;
    while (t0 < t1 && tp[t0 + 1] <= (unsigned long long)first) ++t0;
    while (t1 > t0 && tp[t1] >= (unsigned long long)(first + count)) --t1;
    p.block0 = t0 * kScanTile;
    const long long launch_blocks = std::min<long long>(nblocks, (t1 + 1) * kScanTile) - p.block0;
    enum_emit_kernel<<<(unsigned)launch_blocks, kEnumThreads, smem, st>>>(p, e->d_count_sums, e->d_count_in_tile, e->d_count_tile, first, count, L, triple, code, len,
                                                                    reinterpret_cast<unsigned long long*>(hash));
    count_launch();
    PDE_CUDA(cudaGetLastError());
    exprset_mark_use(e, st);
    return PDE_OK;
}

static int csr_prepare(const pde_exprset* e, const int32_t* depth_begin, int depth, int prune, int64_t first, int64_t count, int L,
                       cudaStream_t st, EnumParams& p, long long& b0, long long& b1) {
    if (!have_device()) { set_error("no CUDA device: pde_engine_b200 has no CPU fallback"); return PDE_E_NODEVICE; }
    if (first < 0 || count < 0) { set_error("pde_enumerate_csr: bad argument"); return PDE_E_INVALID; }
    if (L < 16 || L > kMaxRow || (L % 16) != 0) { set_error("L must be a multiple of 16 in [16, %d]", kMaxRow); return PDE_E_INVALID; }
    int rc = fill_params(e, depth_begin, depth, prune, p);
    if (rc) return rc;
    long long total = 0;
    pde_exprset* em = const_cast<pde_exprset*>(e);
    rc = run_count(em, p, st, &total);
    if (rc) return rc;
    if (first + count > total) { set_error("pde_enumerate_csr: window [%lld, %lld) exceeds the %lld candidates", (long long)first, (long long)(first + count), total); return PDE_E_INVALID; }
    b0 = b1 = 0;
    if (count == 0 || p.seg_block[p.depth] == 0) return PDE_OK;
    rc = run_count_bytes(em, p, L, st);
    if (rc) return rc;
    window_blocks(e, first, count, b0, b1);
    return PDE_OK;
}

int pde_enumerate_csr_size(const pde_exprset* e, const int32_t* depth_begin, int depth, int prune,
                           int64_t first, int64_t count, int L, int64_t* pool_bytes, void* stream) {
    if (!pool_bytes) { set_error("null pool_bytes"); return PDE_E_INVALID; }
    EnumParams p;
    long long b0, b1;
    int rc = csr_prepare(e, depth_begin, depth, prune, first, count, L, (cudaStream_t)stream, p, b0, b1);
    if (rc) return rc;
    *pool_bytes = (count == 0 || p.seg_block[p.depth] == 0) ? 0 : (int64_t)(e->block_bytes_host[b1] - e->block_bytes_host[b0]);
    if (*pool_bytes >= (int64_t)16 * 0xffffffffLL) { set_error("pde_enumerate_csr: the window's pool exceeds 64 GiB (32-bit offsets in 16-byte units): use smaller windows"); return PDE_E_OVERFLOW; }
    return PDE_OK;
}

int pde_enumerate_csr(const pde_exprset* e, const int32_t* depth_begin, int depth, int prune,
                      int64_t first, int64_t count, int L,
                      int32_t* triple, uint32_t* offset, uint8_t* pool, uint8_t* len, uint64_t* hash, void* stream) {
    EnumParams p;
    long long b0, b1;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = csr_prepare(e, depth_begin, depth, prune, first, count, L, st, p, b0, b1);
    if (rc) return rc;
    if (count == 0 || p.seg_block[p.depth] == 0) return PDE_OK;
    if (!triple || !offset || !len || !hash || (!pool && e->block_bytes_host[b1] > e->block_bytes_host[b0])) { set_error("pde_enumerate_csr: null output"); return PDE_E_INVALID; }
    const size_t smem = (size_t)kEnumThreads * (L + 8 + 12 + 4 + 1) + 32;      // packed tile + hash / triple / offset / len per row
    PDE_CUDA(cudaFuncSetAttribute(enum_emit_csr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CsrParams c;
    c.sums = e->d_count_sums; c.in_tile = e->d_count_in_tile; c.tile_off = e->d_count_tile;
    c.bytes_in_tile = e->d_bytes_in_tile; c.bytes_tile = e->d_bytes_tile;
    c.first = first; c.count = count; c.pool_lo = e->block_bytes_host[b0]; c.L = L;
    c.triple = triple; c.pool = pool; c.offset = offset; c.len_out = len; c.hash_out = reinterpret_cast<unsigned long long*>(hash);
    p.block0 = b0;
    enum_emit_csr_kernel<<<(unsigned)(b1 - b0), kEnumThreads, smem, st>>>(p, c);
    count_launch();
    PDE_CUDA(cudaGetLastError());
    exprset_mark_use(e, st);
    return PDE_OK;
}

int pde_dedup_csr(const uint8_t* pool, const uint32_t* offset, const uint8_t* len, const uint64_t* hash, int64_t n,
                  uint8_t* first_occurrence, int64_t* n_unique, void* stream) {
    if (!have_device()) { set_error("no CUDA device: pde_engine_b200 has no CPU fallback"); return PDE_E_NODEVICE; }
    if (n == 0) { if (n_unique) *n_unique = 0; return PDE_OK; }
    if (!pool || !offset || !len || !hash || !first_occurrence || n < 0) { set_error("pde_dedup_csr: bad argument"); return PDE_E_INVALID; }
    if (n > (1LL << 30)) { set_error("pde_dedup_csr: n too large (the 32-bit table index covers 2^30 candidates per call)"); return PDE_E_OVERFLOW; }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned cap = 1024;
    while ((long long)cap < 2 * n) cap <<= 1;
    unsigned long long* keys = nullptr;
    unsigned* vals = nullptr;
    unsigned long long* cnt = nullptr;
    int rc = scratch_alloc(reinterpret_cast<void**>(&keys), sizeof(unsigned long long) * cap, st);
    if (!rc) rc = scratch_alloc(reinterpret_cast<void**>(&vals), sizeof(unsigned) * cap, st);
    if (!rc) rc = scratch_alloc(reinterpret_cast<void**>(&cnt), sizeof(unsigned long long), st);
    if (rc) { scratch_free(keys, st); scratch_free(vals, st); scratch_free(cnt, st); return rc; }
    PDE_CUDA(cudaMemsetAsync(keys, 0, sizeof(unsigned long long) * cap, st));
    PDE_CUDA(cudaMemsetAsync(vals, 0xff, sizeof(unsigned) * cap, st));
    PDE_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), st));
    const unsigned blocks = (unsigned)((n + 255) / 256);
    dedup_insert_kernel<<<blocks, 256, 0, st>>>(len, reinterpret_cast<const unsigned long long*>(hash), n, keys, vals, cap - 1);
    dedup_lookup_csr_kernel<<<blocks, 256, 0, st>>>(pool, offset, len, reinterpret_cast<const unsigned long long*>(hash), n, keys, vals,
                                                    cap - 1, first_occurrence, cnt);
    count_launch(2);
    PDE_CUDA(cudaGetLastError());
    unsigned long long c = 0;
    PDE_CUDA(cudaMemcpyAsync(&c, cnt, sizeof(c), cudaMemcpyDeviceToHost, st));
    PDE_CUDA(cudaStreamSynchronize(st));
    scratch_free(keys, st); scratch_free(vals, st); scratch_free(cnt, st);
    if (n_unique) *n_unique = (int64_t)c;
    return PDE_OK;
}

int pde_dedup(const uint8_t* code, const uint8_t* len, const uint64_t* hash, int64_t n, int L,
              uint8_t* first_occurrence, int64_t* n_unique, void* stream) {
    if (!have_device()) { set_error("no CUDA device: pde_engine_b200 has no CPU fallback"); return PDE_E_NODEVICE; }
    if (n == 0) { if (n_unique) *n_unique = 0; return PDE_OK; }      // an empty batch is a no-op (its buffers may be null)
    if (!code || !len || !hash || !first_occurrence || n < 0 || (L % 16) != 0) { set_error("pde_dedup: bad argument"); return PDE_E_INVALID; }
    if (n > (1LL << 30)) { set_error("pde_dedup: n too large (the 32-bit table index covers 2^30 candidates per call)"); return PDE_E_OVERFLOW; }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned cap = 1024;
    while ((long long)cap < 2 * n) cap <<= 1;
    // per-call table from the library's stream-ordered pool (kept warm between calls: allocating and freeing 400 MB
    // with cudaMalloc cost 8-70 ms around 1.7 ms of kernels)
    unsigned long long* keys = nullptr;
    unsigned* vals = nullptr;
    unsigned long long* cnt = nullptr;
    int rc = scratch_alloc(reinterpret_cast<void**>(&keys), sizeof(unsigned long long) * cap, st);
    if (!rc) rc = scratch_alloc(reinterpret_cast<void**>(&vals), sizeof(unsigned) * cap, st);
    if (!rc) rc = scratch_alloc(reinterpret_cast<void**>(&cnt), sizeof(unsigned long long), st);
    if (rc) { scratch_free(keys, st); scratch_free(vals, st); scratch_free(cnt, st); return rc; }
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
    scratch_free(keys, st); scratch_free(vals, st); scratch_free(cnt, st);
    if (n_unique) *n_unique = (int64_t)c;
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
