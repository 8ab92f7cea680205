// validate.cuh -- the batched jet interpreter (stage 2 of the hot path).
//
// Replaces the per-candidate validator.validate() call of emit_to_db
// (general_method_paper_reproduction.py:1302-1316) and the validator worker
// pool (GM:1672-1824) as a numerical FILTER in front of the symbolic check.
//
// Mapping (BASELINE.json north_star): WARPS own candidates, each LANE owns
// collocation points (32-point stripes; two points per lane for the small Kerr
// jets).  A warp
//   1. stages a candidate's postfix bytecode in shared memory,
//   2. translates it once into leaf-fused micro-ops (lane 0; evaluation order
//      chosen per binary node so that two spill slots always suffice),
//   3. together with the 3 other warps of its group, for every point stripe
//      interprets the micro-ops on a register-resident top-of-stack jet T
//      (+ operand jet U); deeper stack entries spill to a per-lane column in
//      shared memory (conflict free),
//   4. applies the problem's residual operator to the finished jet,
//   5. reduces votes / maxima over the lanes with warp shuffles.
// Design history and the ncu evidence behind every choice: profiles/README.md.
#pragma once
#include <stdint.h>
#include "jet.cuh"
#include "residual_ff_gen.cuh"
#include "../../include/pde_b200.h"

#ifndef PDE_UNIVARIATE
#define PDE_UNIVARIATE 1       // single-axis bodies for sub-expressions that depend on one coordinate
#endif

namespace pde {

constexpr int kMaxL = 256;

// per-launch tables (warp-uniform reads -> constant cache).  `static`: every translation unit that instantiates the
// kernel (pde_b200.cu: the built-in residuals; program.cu: run-time residual programs) owns and uploads its own copy.
static __constant__ double c_const[PDE_N_CONST];
static __constant__ double c_rconst[PDE_N_CONST];   // reciprocals (division by a constant leaf)
static __constant__ double c_pow[PDE_N_POW];
// Taylor-ratio rows of x**k (U_POW): f_{j+1} = f_j * row[j] / x_0, row[j] = (k - j)/(j + 1)
constexpr int kNRows = PDE_N_POW;
static __constant__ double c_frow[kNRows][4];
// x**n with n = 0, 1, 2, ...: c_pow_int[slot] = n (else -1) and the binomial coefficients C(n, j), j = 0..4, so that
// the Taylor coefficients C(n, j) x^(n-j) need no division and stay finite at x = 0
static __constant__ int c_pow_int[kNRows];
static __constant__ double c_fbin[kNRows][5];
// round-off majorants (oracle/majorant.py): expansion radius t0 and theta_n / W = 2 eps n! / (t0^n tau), n = 1..4
static __constant__ float c_t0;
static __constant__ float c_constf[PDE_N_CONST];    // |c_const| and |c_rconst| as float (clamped like maj_abs), c_pow as float
static __constant__ float c_rconstf[PDE_N_CONST];
static __constant__ float c_powf[kNRows];
static __constant__ double c_theta[4];
static __constant__ double c_one = 1.0;                       // constant-bank operand: no register, no per-dispatch move
static __constant__ double c_sign[2] = {1.0, -1.0};
static __constant__ double c_rfact[4] = {1.0, 1.0 / 2, 1.0 / 3, 1.0 / 4};   // 1/(j + 1)

// Micro-ops.  Every arithmetic body exists exactly ONCE in the kernel so the interpreter's
// code stays inside the instruction cache: an earlier version that inlined the bodies per
// call site ran at a 75 % i-cache hit rate (profiles/r1_v1_*).
// Word layout (u32): bits 0-7 kind | 8-15 arg (constant slot, exponent slot, PRIM index or
// coordinate) | 16-23 flags (bit 0: spill T first; FN kinds: bits 1-2 function, 3-7 exponent slot).
// The kinds are DENSE and the interpreter is one `switch` compiled with --jump-table-density:
// a micro-op costs one indexed branch (LDC + BRX) instead of the 5-level compare tree of v5
// (33 % of all warp stall samples were dispatch, profiles/README.md).  Binary bodies have one
// entry per operand source (_S = pop the spill stack, _P = primitive table): the entry fetches
// the operand jet U and falls into the shared body.
enum UKind : uint8_t {
    U_END = 0,
    U_SETV0, U_SETV1, U_SETC, U_SETP,         // T = leaf (first leaf of a sub-tree); flag bit 0: S[sp++] = T first
    U_FNV, U_FNC,                             // T = F(coordinate), F(constant): F's Taylor coefficients ARE the jet
    U_ADD_S, U_ADD_P,                         // T = T + U
    U_SUB_S, U_SUB_P,                         // T = T - U
    U_RSUB_S,                                 // T = U - T
    U_MUL_S, U_MUL_P,                         // T = T * U
    U_DIV_S, U_DIV_P,                         // T = T / U  (in place on the numerator)
    U_RDIV_S, U_RDIV_P,                       // T = U / T  (in place on U, copied back)
    U_RDIV_V0, U_RDIV_V1, U_RDIV_C,           // T = coordinate / T, constant / T (1 / T, inv(T)): the same quotient body with U = the leaf's jet
    U_ADDC, U_SUBC, U_RSUBC, U_MULC, U_MULRC, // sparse leaf fast paths (arg = const slot; MULRC: reciprocal)
    U_ADDV0, U_ADDV1, U_SUBV0, U_SUBV1, U_MULV0, U_MULV1, U_DIVV0, U_DIVV1,
    U_NEG, U_ABS, U_SQRT, U_SQUARE,
    U_EXP,                                    // exp(+-T), arg = sign } one shared composition body
    U_POW,                                    // arg = exponent slot  } (jetv_compose_ps), in place
    U_NKINDS
};
constexpr unsigned F_SPILL = 1u << 16;
// U_SQRT, U_SQUARE, U_EXP, U_POW only (translate pass 3): T depends on ONE coordinate (or none), F_AXIS1 says
// which; the body then runs on that axis' 5 coefficients instead of all 15
constexpr unsigned F_UNI = 1u << 17;
constexpr unsigned F_AXIS1 = 1u << 18;
enum UFn : unsigned { FN_INV = 0, FN_EXP = 1, FN_EXPN = 2, FN_POW = 3 };

struct ValidateParams {
    const uint8_t* code;
    const unsigned* row_off;   // CSR rows (pde_validate_csr): program c starts at code + 16 * row_off[c]; null: code + c * L
    const uint8_t* len;
    long long n;
    int L;
    const double* pts;    // [2][P]
    const double* tab;    // [cols][P]
    const double* prim;   // [n_prim][P/32][16][32]: per 32-point stripe, coefficient-major, lanes contiguous
    int n_prim;           // PRIM(p) with p >= n_prim is malformed input
    int P;
    int P_eval;           // points evaluated (the first P_eval of the grid; a multiple of 128 * NP or == P)
    // item i of this launch is candidate index[i] (null: i), i < *n_index (null: n) -- both on the device.  First pass:
    // the evaluation order (candidate_order, common.h); confirmation pass (is_confirm): the proposed rejections
    const int* index;
    const unsigned long long* n_index;
    int is_confirm;
    int* confirm;         // [n, 2] (n_finite, n_votes) of the confirmation pass, or null
    unsigned long long* chunk_counter;   // zeroed by the launcher: chunks are dealt dynamically (null: round-robin)
    int ns;               // spill slots per lane
    float t0;             // expansion radius of the round-off majorants
    double tau;
    int min_finite;
    double vote_frac;
    int n_ref;
    // reduce-mode outputs
    double* ratio_max;
    double* resid_max;
    double* scale_at;
    int* n_finite;
    int* n_votes;
    double* ref_rs;
    unsigned* survivor_bits;
    // dump-mode outputs
    double* jets;
    double* resid;
    double* scale;        // S: sum of |monomial| of the residual (the scale parity is quoted against)
    double* scale_maj;    // S~: the decision scale (isotropic majorant at the round-off-inflated partials)
    float* maj;           // [n, 3, P]: (V, D, W) of the finished jet
};

__device__ __forceinline__ bool op_is_prim(unsigned b) { return b >= PDE_OP_PRIM0 && b < PDE_OP_PRIM0 + PDE_N_PRIM; }
__device__ __forceinline__ bool op_is_leaf(unsigned b) {
    return b == PDE_OP_VAR0 || b == PDE_OP_VAR1 || op_is_prim(b) || b >= PDE_OP_CONST0;
}
__device__ __forceinline__ bool op_is_binary(unsigned b) { return b >= PDE_OP_ADD && b <= PDE_OP_DIV; }
__device__ __forceinline__ bool op_is_unary(unsigned b) {
    return (b >= PDE_OP_NEG && b <= PDE_OP_EXP) || (b >= PDE_OP_FN_NEG && b <= PDE_OP_FN_EXPNEG) ||
           (b >= PDE_OP_POW0 && b < PDE_OP_POW0 + PDE_N_POW);
}

__host__ __device__ constexpr int kUcodeMax(int L) { return 2 * L + 6; }

// Postfix bytecode -> micro-ops.  Returns 0 ok, 1 malformed, 2 spill overflow.
//
// Machine model: the most recent unfinished value is the register jet T, older ones are spilled in
// stack order, leaves never occupy a jet (they are fused into the op that consumes them, or set into
// T by the first op of a sub-tree -- which spills T if it holds a live value).
// Evaluation ORDER is chosen per binary node (Sethi-Ullman): when both operands are sub-trees the one
// that needs more spill slots is evaluated first, so need(node) = min(max(nl, nr + 1), max(nr, nl + 1))
// instead of the postfix order's max(nl, nr + 1).  On the 143 461 real depth-4 force-free uniques the
// postfix order needs up to 7 slots (12 473 candidates need more than 2); this order needs at most 2 for
// every one of them.  Ties evaluate a denominator first: T / S then runs in place on the numerator.
//   pass 1: sub-tree start and spill need of every position (stack of root positions in `stk`);
//   pass 2: emission with an explicit frame stack (position, phase) in `stk`.
// start[], need[] and stk[] are caller-provided byte arrays of L, L and 2L bytes.
static __device__ __noinline__ int translate(const uint8_t* code, int len, uint32_t* uc, uint8_t* start, uint8_t* need, uint8_t* stk, int ns_max, int n_prim) {
    // ---- pass 1 ----
    int sp = 0;
    for (int i = 0; i < len; ++i) {
        const unsigned b = code[i];
        if (op_is_leaf(b)) {
            if (op_is_prim(b) && (int)(b - PDE_OP_PRIM0) >= n_prim) return 1;      // no table row behind it
            start[i] = (uint8_t)i; need[i] = 0; stk[sp++] = (uint8_t)i;
        } else if (op_is_unary(b)) {
            if (sp < 1) return 1;
            const int c = stk[sp - 1];
            start[i] = start[c]; need[i] = need[c] & 0x7f; stk[sp - 1] = (uint8_t)i;
        } else if (op_is_binary(b)) {
            if (sp < 2) return 1;
            const int r = stk[sp - 1], l = stk[sp - 2];
            sp -= 2;
            start[i] = start[l];
            const bool lf_l = op_is_leaf(code[l]), lf_r = op_is_leaf(code[r]);
            const int nl = need[l] & 0x7f, nr = need[r] & 0x7f;
            if (!lf_l && !lf_r) {
                const int left_first = nl > nr + 1 ? nl : nr + 1, right_first = nr > nl + 1 ? nr : nl + 1;
                const bool rf = right_first < left_first || (right_first == left_first && b == PDE_OP_DIV);
                need[i] = (uint8_t)((rf ? right_first : left_first) | (rf ? 0x80 : 0));
            } else {
                need[i] = (uint8_t)(lf_l ? (lf_r ? 0 : nr) : nl);
            }
            stk[sp++] = (uint8_t)i;
        } else {
            return 1;
        }
    }
    if (sp != 1) return 1;
    const int root = stk[0];
    if ((need[root] & 0x7f) > ns_max) return 2;

    // ---- pass 2 ----
    int nu = 0, ns = 0;
    bool t_live = false;
    auto emit = [&](unsigned kind, unsigned arg, unsigned flags = 0) { uc[nu++] = kind | (arg << 8) | flags; };
    // T is about to be overwritten by a leaf: a live value moves to the spill stack (a flag of the SET/FN op)
    auto spill_flag = [&]() -> unsigned {
        if (!t_live) { t_live = true; return 0; }
        ++ns;
        return F_SPILL;
    };
    auto set_leaf = [&](unsigned leaf) {
        const unsigned flag = spill_flag();
        if (leaf >= PDE_OP_CONST0) emit(U_SETC, leaf - PDE_OP_CONST0, flag);
        else if (op_is_prim(leaf)) emit(U_SETP, leaf - PDE_OP_PRIM0, flag);
        else emit(leaf == PDE_OP_VAR0 ? U_SETV0 : U_SETV1, 0, flag);
    };
    // T = T op leaf (leaf on the right); o: 0 add 1 sub 2 mul 3 div
    auto bin_leaf_right = [&](unsigned o, unsigned leaf) {
        if (leaf >= PDE_OP_CONST0) {
            emit(o == 0 ? U_ADDC : o == 1 ? U_SUBC : o == 2 ? U_MULC : U_MULRC, leaf - PDE_OP_CONST0);
        } else if (leaf == PDE_OP_VAR0 || leaf == PDE_OP_VAR1) {
            emit((o == 0 ? U_ADDV0 : o == 1 ? U_SUBV0 : o == 2 ? U_MULV0 : U_DIVV0) + (leaf - PDE_OP_VAR0), 0);
        } else {
            emit(o == 0 ? U_ADD_P : o == 1 ? U_SUB_P : o == 2 ? U_MUL_P : U_DIV_P, leaf - PDE_OP_PRIM0);
        }
    };
    // the unary op at position i applied to T, or fused with its leaf operand
    auto unary = [&](unsigned b, int leaf_operand /* -1: operand is T */) {
        unsigned fn = 4, slot = 0, kind = U_NEG, arg = 0;     // fn 4 = not one of the scalar-function kinds
        switch (b) {
            case PDE_OP_NEG: case PDE_OP_FN_NEG: kind = U_NEG; break;
            case PDE_OP_ABS: kind = U_ABS; break;
            case PDE_OP_SQRT: kind = U_SQRT; break;
            case PDE_OP_FN_SQUARE: kind = U_SQUARE; break;
            case PDE_OP_EXP: kind = U_EXP; arg = 0; fn = FN_EXP; break;
            case PDE_OP_FN_EXPNEG: kind = U_EXP; arg = 1; fn = FN_EXPN; break;
            case PDE_OP_FN_INV: kind = U_RDIV_C; arg = 0; fn = FN_INV; break;          // inv(T) = CONST(0) / T, CONST(0) = 1
            case PDE_OP_FN_POW32: kind = U_POW; arg = 0; fn = FN_POW; slot = 0; break;
            case PDE_OP_FN_POWN32: kind = U_POW; arg = 1; fn = FN_POW; slot = 1; break;
            default: {
                slot = b - PDE_OP_POW0;
                const double k = c_pow[slot];
                if (k == 2.0) kind = U_SQUARE;
                else if (k == 0.5) kind = U_SQRT;
                else if (k == -1.0) { kind = U_RDIV_C; arg = 0; fn = FN_INV; }
                else { kind = U_POW; arg = slot; fn = FN_POW; }
            }
        }
        if (leaf_operand >= 0) {
            const unsigned leaf = (unsigned)leaf_operand;
            if (fn < 4 && slot < 32 && !op_is_prim(leaf)) {
                // F(coordinate) / F(constant): the jet is F's own Taylor expansion, no jet arithmetic
                const unsigned fl = spill_flag() | (fn << 17) | (slot << 19);
                if (leaf >= PDE_OP_CONST0) emit(U_FNC, leaf - PDE_OP_CONST0, fl);
                else emit(U_FNV, leaf - PDE_OP_VAR0, fl);
                return;
            }
            set_leaf(leaf);
        }
        emit(kind, arg);
    };
    if (op_is_leaf(code[root])) {
        set_leaf(code[root]);
    } else {
        int fs = 0;     // frames: stk[2 f] = position, stk[2 f + 1] = phase
        stk[0] = (uint8_t)root; stk[1] = 0; fs = 1;
        while (fs > 0) {
            const int i = stk[2 * fs - 2], ph = stk[2 * fs - 1];
            const unsigned b = code[i];
            if (op_is_unary(b)) {
                const int c = i - 1;
                if (ph == 0 && !op_is_leaf(code[c])) {
                    stk[2 * fs - 1] = 1;
                    stk[2 * fs] = (uint8_t)c; stk[2 * fs + 1] = 0; ++fs;
                } else {
                    --fs;
                    unary(b, ph == 0 ? (int)code[c] : -1);
                }
                continue;
            }
            const int r = i - 1, l = start[r] - 1;
            const unsigned bl = code[l], br = code[r], o = b - PDE_OP_ADD;
            const bool lf_l = op_is_leaf(bl), lf_r = op_is_leaf(br), rf = (need[i] & 0x80) != 0;
            if (ph == 0) {
                if (lf_l && lf_r) {
                    --fs;
                    set_leaf(bl);
                    bin_leaf_right(o, br);
                } else {
                    const int first = lf_r ? l : lf_l ? r : (rf ? r : l);
                    stk[2 * fs - 1] = 1;
                    stk[2 * fs] = (uint8_t)first; stk[2 * fs + 1] = 0; ++fs;
                }
            } else if (ph == 1) {
                if (lf_r) {                                      // T op leaf
                    --fs;
                    bin_leaf_right(o, br);
                } else if (lf_l) {                               // leaf op T
                    --fs;
                    if (o == 0 || o == 2) bin_leaf_right(o, bl);                        // commutative
                    else if (o == 1) {                                                  // leaf - T
                        if (bl >= PDE_OP_CONST0) emit(U_RSUBC, bl - PDE_OP_CONST0);
                        else { emit(U_NEG, 0); bin_leaf_right(0, bl); }
                    } else if (op_is_prim(bl)) emit(U_RDIV_P, bl - PDE_OP_PRIM0);        // PRIM / T
                    else if (bl == PDE_OP_VAR0 || bl == PDE_OP_VAR1) emit(U_RDIV_V0 + (bl - PDE_OP_VAR0), 0);   // x / T by the quotient recurrence
                    else emit(U_RDIV_C, bl - PDE_OP_CONST0);                            // c / T by the quotient recurrence
                } else {
                    const int second = rf ? l : r;
                    if (ns >= ns_max) return 2;                  // the second sub-tree's first leaf will spill T
                    stk[2 * fs - 1] = 2;
                    stk[2 * fs] = (uint8_t)second; stk[2 * fs + 1] = 0; ++fs;
                }
            } else {
                --fs;
                // left first: S = a, T = b;  right first: S = b, T = a
                if (rf) emit(o == 0 ? U_ADD_S : o == 1 ? U_SUB_S : o == 2 ? U_MUL_S : U_DIV_S, 0);
                else emit(o == 0 ? U_ADD_S : o == 1 ? U_RSUB_S : o == 2 ? U_MUL_S : U_RDIV_S, 0);
                --ns;
            }
        }
    }
    emit(U_END, 0);
#if PDE_UNIVARIATE
    // ---- pass 3: which coordinates does every jet depend on?  (bit 0: coordinate 0, bit 1: coordinate 1) ----
    // The stream is straight-line code, so one walk with a mask for T and a mask stack for the spilled jets
    // (in `stk`, free again) is exact.  A heavy body whose result depends on at most one coordinate is replaced by
    // its single-axis version: 15 instead of 70 multiply-adds for an order-4 product.  Constants ride on axis 0.
    // PRIM rows are treated as bivariate.  The light bodies are correct as they are (zeros stay zeros).
    unsigned mt = 0;
    int msp = 0;
    for (int k = 0; k + 1 < nu; ++k) {
        const unsigned w = uc[k], kind = w & 0xffu, arg = (w >> 8) & 0xffu;
        switch (kind) {
            case U_SETV0: case U_SETV1: case U_SETC: case U_SETP: case U_FNV: case U_FNC:
                if (w & F_SPILL) stk[msp++] = (uint8_t)mt;
                mt = kind == U_SETV0 ? 1u : kind == U_SETV1 ? 2u : kind == U_SETP ? 3u : kind == U_FNV ? (1u << arg) : 0u;
                break;
            case U_ADD_S: case U_SUB_S: case U_RSUB_S: case U_MUL_S: case U_DIV_S: case U_RDIV_S: mt |= stk[--msp]; break;
            case U_ADD_P: case U_SUB_P: case U_MUL_P: case U_DIV_P: case U_RDIV_P: mt = 3u; break;
            case U_ADDV0: case U_SUBV0: case U_MULV0: case U_DIVV0: case U_RDIV_V0: mt |= 1u; break;
            case U_ADDV1: case U_SUBV1: case U_MULV1: case U_DIVV1: case U_RDIV_V1: mt |= 2u; break;
            case U_SQRT: case U_SQUARE: case U_EXP: case U_POW:
                if (mt != 3u) uc[k] = w | F_UNI | (mt == 2u ? F_AXIS1 : 0u);
                break;
            default: break;
        }
    }
#endif
    return 0;
}

template <int N>
struct PointCtx {
    double x0, x1;
    float ax0, ax1;       // |x0|, |x1| as float (maj_abs): the V of a coordinate leaf
    const double* prim;   // this point's slot in PRIM(0)'s stripe block: coefficient g at [g * 32]
};

__device__ __forceinline__ double opaque_zero() {
    double z;
    asm volatile("mov.f64 %0, 0d0000000000000000;" : "=d"(z));
    return z;
}
template <int N>
__device__ __forceinline__ void jet_fill(Jet<N>& t, double v) {
#pragma unroll
    for (int g = 0; g < Jet<N>::NC; ++g) t.c[g] = v;
}

// 1/x without letting the compiler hoist it out of the interpreter loop (only DIVV needs it)
__device__ __forceinline__ double lazy_rcp(double x) {
    double r;
    asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

// Shared-memory accesses of the interpreter go through 32-bit shared-window addresses held in
// registers: with ordinary pointers ptxas re-derives the window base (S2R SR_CgaCtaId + LEA + IMAD)
// and the thread's column (S2R SR_TID) in front of EVERY micro-op instead of keeping two registers.
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned keep_in_register(unsigned v) {
    asm volatile("" : "+r"(v));
    return v;
}
__device__ __forceinline__ unsigned lds_u32(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
template <int OFF>
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ void sts_f64(unsigned addr, double v) {
    asm volatile("st.shared.f64 [%0+%1], %2;" ::"r"(addr), "n"(OFF), "d"(v) : "memory");
}
template <int OFF>
__device__ __forceinline__ void lds_f32x2(unsigned addr, float& a, float& b) {
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(a), "=f"(b) : "r"(addr), "n"(OFF));
}
template <int OFF>
__device__ __forceinline__ void sts_f32x2(unsigned addr, float a, float b) {
    asm volatile("st.shared.v2.f32 [%0+%1], {%2, %3};" ::"r"(addr), "n"(OFF), "f"(a), "f"(b) : "memory");
}
// A spill slot holds NC * NP jet coefficients + 2 NP majorant elements; element k lives at byte offset k * TPB * 8
// (an immediate of the LDS/STS): coefficient g of point h at k = g * NP + h, the majorants of point h at
// NC * NP + 2 h (V, D) and NC * NP + 2 h + 1 (W).
struct Maj;
template <int N, int NP>
__host__ __device__ constexpr int spill_slot_elems() { return (Jet<N>::NC + 2) * NP; }
template <int N, int NP, int TPB, bool MAJ, int K = 0, class M>
__device__ __forceinline__ void spill_store(unsigned addr, const Jet<N> (&T)[NP], const M (&mj)[NP]) {
    if constexpr (K < Jet<N>::NC * NP) {
        sts_f64<K * TPB * 8>(addr, T[K % NP].c[K / NP]);
        spill_store<N, NP, TPB, MAJ, K + 1>(addr, T, mj);
    } else if constexpr (MAJ && K < spill_slot_elems<N, NP>()) {
        constexpr int h = (K - Jet<N>::NC * NP) / 2;
        if constexpr (((K - Jet<N>::NC * NP) & 1) == 0) sts_f32x2<K * TPB * 8>(addr, mj[h].V, mj[h].D);
        else sts_f32x2<K * TPB * 8>(addr, mj[h].W, 0.0f);
        spill_store<N, NP, TPB, MAJ, K + 1>(addr, T, mj);
    }
}
template <int N, int NP, int TPB, bool MAJ, int K = 0, class M>
__device__ __forceinline__ void spill_load(unsigned addr, Jet<N> (&U)[NP], M (&mj)[NP]) {
    if constexpr (K < Jet<N>::NC * NP) {
        U[K % NP].c[K / NP] = lds_f64<K * TPB * 8>(addr);
        spill_load<N, NP, TPB, MAJ, K + 1>(addr, U, mj);
    } else if constexpr (MAJ && K < spill_slot_elems<N, NP>()) {
        constexpr int h = (K - Jet<N>::NC * NP) / 2;
        float pad;
        if constexpr (((K - Jet<N>::NC * NP) & 1) == 0) lds_f32x2<K * TPB * 8>(addr, mj[h].V, mj[h].D);
        else lds_f32x2<K * TPB * 8>(addr, mj[h].W, pad);
        spill_load<N, NP, TPB, MAJ, K + 1>(addr, U, mj);
    }
}

// Taylor coefficients f_j = F^(j)(x)/j! of the scalar functions (UFn) at x
template <int N>
__device__ __forceinline__ void scalar_taylor(unsigned fn, unsigned slot, double x, double (&f)[N + 1]) {
    if (fn == FN_INV) {                  // (-1)^j / x^(j+1)
        const double r = fast_rcp(x);
        f[0] = r;
#pragma unroll
        for (int j = 0; j < N; ++j) f[j + 1] = -f[j] * r;
    } else if (fn == FN_POW) {
        const int n = c_pow_int[slot];
        if (n >= 0) {
            // x**n, n = 0, 1, 2, ...: f_j = C(n, j) x^(n-j) by products only -- exact structure, finite at x = 0
            // (the ratio recurrence below divides by x: `(2*z - 1)**3` at the reference point z = 1/2 was NaN)
            int ex = n > N ? n - N : 0;
            double pw = 1.0, base = x;
#pragma unroll 1
            for (int e = ex; e; e >>= 1) { if (e & 1) pw *= base; base *= base; }
#pragma unroll
            for (int j = N; j >= 0; --j) {
                const int want = n > j ? n - j : 0;
                if (want > ex) { pw *= x; ex = want; }
                f[j] = c_fbin[slot][j] * pw;
            }
        } else {                         // f_{j+1} = f_j (k - j)/(j + 1) / x
            const double r = fast_rcp(x);
            f[0] = pow0(x, c_pow[slot]);
#pragma unroll
            for (int j = 0; j < N; ++j) f[j + 1] = f[j] * (r * c_frow[slot][j]);
        }
    } else {                             // exp(+-x): (+-1)^j exp(+-x) / j!
        const double sg = c_sign[fn == FN_EXPN];
        f[0] = fast_exp(sg * x);
#pragma unroll
        for (int j = 0; j < N; ++j) f[j + 1] = f[j] * (sg * c_rfact[j]);
    }
}

// ---------------------------------------------------------------------------------
// Round-off majorants (the calculus and its proof sketch: oracle/majorant.py).  Next to every jet the
// interpreter carries three float32 numbers:  V >= |c_0| (the value, summed WITHOUT cancellation),
// D >= sum_{n>=1} [C]_n t0^n (the non-constant part of a one-variable majorant series of the jet, evaluated at the
// radius t0) and W, the same for the accumulated rounding error in units of 2^-52, so that
// |computed c_g - exact c_g| <= 2^-52 W / t0^|g|.  The residual's decision scale is evaluated at partials inflated by
// that bound (Residual<>::eval): a point votes "non-zero" only if no float64 evaluation order of an exact solution
// could have produced its |R|.  float32 because magnitudes need no precision and the FP64 pipe is the bottleneck:
// the rules run on the FMA / MUFU pipes (approximate reciprocal, log2, exp2: a 1e-4 safety factor covers them);
// only quotients and compositions convert the ACTUAL value of their operand (one F2F each: the pole distance
// |d_0| - D needs it).  Overflow gives inf / NaN (the point does not vote).
// ---------------------------------------------------------------------------------
struct Maj { float V, D, W; };
__device__ __forceinline__ float f_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float f_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float f_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float f_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
constexpr float kMajSlop = 1.0001f;      // MUFU approximations (2^-22 absolute in lg2 / ex2 arguments, 1 ulp in rcp)
// |x| of a jet's actual value as float, clamped from below so that products of magnitudes cannot flush to zero
__device__ __forceinline__ float maj_abs(double x) { return fmaxf(__double2float_rn(fabs(x)), 1e-18f); }
__device__ __forceinline__ Maj maj_leaf(float v, float d, float w) { Maj m; m.V = v; m.D = d; m.W = w; return m; }
// a +- b
__device__ __forceinline__ void maj_add(Maj& a, const Maj& b) {
    a.W = (a.W + b.W) + ((a.V + a.D) + (b.V + b.D));
    a.V += b.V;
    a.D += b.D;
}
// Every rule that rounds clamps its V and W from below: float32 products of small magnitudes would flush to zero
// and a zero W means "exact" (a coordinate under -x / |x|).  A clamped majorant is merely larger than necessary.
constexpr float kMajFloor = 1e-30f;
__device__ __forceinline__ float maj_floor(float x) { return x < kMajFloor ? kMajFloor : x; }    // NaN (outside a radius) stays NaN
// a * b
__device__ __forceinline__ void maj_mul(Maj& a, const Maj& b) {
    const float Ma = a.V + a.D, Mb = b.V + b.D;
    a.W = maj_floor(fmaf(a.W, Mb, fmaf(Ma, b.W, 16.0f * Ma * Mb)));
    a.D = fmaf(a.V, b.D, a.D * Mb);
    a.V = maj_floor(a.V * b.V);
}
// q = a / d; d0 = |actual value of d|
__device__ __forceinline__ Maj maj_div(const Maj& a, const Maj& d, float d0) {
    const float den = d0 - d.D;                               // distance to the pole, in majorant terms
    const float r = den > 0.0f ? f_rcp(den) * kMajSlop : __int_as_float(0x7f800000);
    Maj q;
    q.V = maj_floor(a.V * f_rcp(d0) * kMajSlop);
    q.D = fmaf(q.V, d.D, a.D) * r;
    const float Mq = q.V + q.D;
    q.W = maj_floor((fmaf(Mq, d.W, a.W) + 16.0f * fmaf(d.D, Mq, a.V + a.D)) * r);
    return q;
}
// F(x): G >= sum |F_j| D^j, G1 >= dG/dD  ->  V = G, D = G1 D, W = G1 W + cF G
__device__ __forceinline__ void maj_apply(Maj& x, float G, float G1, float cF) {
    G = maj_floor(G);
    x.W = fmaf(G1, x.W, cF * G);
    x.D = G1 * x.D;
    x.V = G;
}
// x ** n, n = 0, 1, 2, ... (all binomial coefficients positive: no actual value needed)
__device__ __forceinline__ void maj_ipow(Maj& x, float n) {
    const float M = x.V + x.D;
    const float G = f_ex2(n * f_lg2(M)) * kMajSlop;
    maj_apply(x, G, n * G * f_rcp(M) * kMajSlop, 16.0f * n);
}
__device__ __forceinline__ void maj_square(Maj& x) {
    const float M = x.V + x.D;
    maj_apply(x, M * M, 2.0f * M, 32.0f);
}
// x ** k for any other constant k; a = |actual value|.  |binom(k, j)| <= binom(|k| + j - 1, j): G = a^k (1 - D/a)^-|k|
__device__ __forceinline__ void maj_pow(Maj& x, float k, float a) {
    const float ak = fabsf(k), r = a - x.D;                   // r <= 0: outside the radius -> NaN, the point does not vote
    const float G = f_ex2(fmaf(k, f_lg2(a), -ak * f_lg2(r * f_rcp(a)))) * kMajSlop;
    maj_apply(x, G, G * ak * f_rcp(r) * kMajSlop, 64.0f);
}
__device__ __forceinline__ void maj_sqrt(Maj& x, float a) {   // k = 1/2: G = a / sqrt(a - D)
    const float r = a - x.D;
    const float G = a * f_rsqrt(r) * kMajSlop;
    maj_apply(x, G, 0.5f * G * f_rcp(r) * kMajSlop, 64.0f);
}
// exp(sg * x); x0 = the actual (signed) value
__device__ __forceinline__ void maj_exp(Maj& x, float sg, float x0) {
    const float G = f_ex2(fmaf(sg, x0, x.D) * 1.4426950408889634f) * kMajSlop;
    maj_apply(x, G, G, 64.0f);
}
// the scalar-function kinds of U_FNV / U_FNC (fn, slot) applied to a leaf (x0 = its actual value)
__device__ __forceinline__ void maj_fn(unsigned fn, unsigned slot, float x0, Maj& x) {
    const float a = fmaxf(fabsf(x0), 1e-18f);
    if (fn == FN_INV) {
        x = maj_div(maj_leaf(1.0f, 0.0f, 1.0f), x, a);
    } else if (fn == FN_POW) {
        if (c_pow_int[slot] >= 0) maj_ipow(x, c_powf[slot]); else maj_pow(x, c_powf[slot], a);
    } else {
        maj_exp(x, fn == FN_EXPN ? -1.0f : 1.0f, x0);
    }
}

// Interpret the micro-ops for NP points per lane at once: results in T[0..NP), majorants of the result in MT.
// uc: shared address of the micro-op words; sp_addr: shared address of this thread's spill column
// (layout [slot][element][thread], conflict free), moved up and down by one slot.
// MAJ = false: the majorants are not carried (their arithmetic has no side effects and is dead-code eliminated; the
// spill / table traffic is compiled out explicitly) -- the fast first pass of pde_validate.
template <int N, int NP, int TPB, bool MAJ>
__device__ __forceinline__ void run_program(unsigned uc, unsigned sp_addr,
                                            size_t prim_stride, const PointCtx<N> (&cx)[NP], Jet<N> (&T)[NP], Maj (&MT)[NP]) {
    constexpr int NC = Jet<N>::NC;
    constexpr unsigned kSlotBytes = spill_slot_elems<N, NP>() * TPB * 8;
    Jet<N> U[NP];
    Maj MU[NP];
    double f[NP][N + 1];
    const float t0 = c_t0;
#define PDE_EACH for (int h = 0; h < NP; ++h)
#define PDE_SPILL_IF_FLAGGED                                                                 \
    if (ins_cur & F_SPILL) {                                                                 \
        spill_store<N, NP, TPB, MAJ>(sp_addr, T, MT);                                             \
        sp_addr += kSlotBytes;                                                               \
    }
#define PDE_FETCH_S                                                                          \
    {                                                                                        \
        sp_addr -= kSlotBytes;                                                               \
        spill_load<N, NP, TPB, MAJ>(sp_addr, U, MU);                                              \
    }
// PRIM(p): jet rows 0..NC-1 and the (D, W) pair in row 15 of the stripe block; V from the value itself
#define PDE_LOAD_P(J, M)                                                                     \
    {                                                                                        \
        _Pragma("unroll") PDE_EACH {                                                         \
            const double* from = cx[h].prim + (size_t)arg * prim_stride;                     \
            _Pragma("unroll") for (int g = 0; g < NC; ++g) J[h].c[g] = __ldg(from + g * 32); \
            if (MAJ) {                                                                       \
                const float2 mj = __ldg(reinterpret_cast<const float2*>(from + 15 * 32));    \
                M[h] = maj_leaf(maj_abs(J[h].c[0]), mj.x, mj.y);                             \
            }                                                                                \
        }                                                                                    \
    }
#define PDE_FETCH_P PDE_LOAD_P(U, MU)
// the operand jet of U_RDIV_V*: a coordinate
#define PDE_U_VAR(X, AX, GSLOT)                                                              \
    {                                                                                        \
        const double z = opaque_zero();                                                      \
        _Pragma("unroll") PDE_EACH { jet_fill(U[h], z); U[h].c[0] = cx[h].X; U[h].c[GSLOT] = c_one; MU[h] = maj_leaf(cx[h].AX, t0, 0.0f); } \
    }
#pragma unroll 1
    for (;;) {
        // no software prefetch of the next word: measured equal (135.9 / 136.0 / 136.6 ms for depth 0 / 1 / 2),
        // and every prefetched word costs a move per dispatch
        const unsigned ins_cur = lds_u32(uc); uc += 4;
        const unsigned kind = ins_cur & 0xffu, arg = (ins_cur >> 8) & 0xffu;
        switch (kind) {
            case U_END: return;
            // The SET bodies go through an opaque zero: a case that only assigns constants becomes an
            // EMPTY block, the indexed branch then jumps straight to the loop header and the header's
            // phi copies (30 register moves) land in front of the branch -- executed by EVERY micro-op.
            case U_SETV0: {
                PDE_SPILL_IF_FLAGGED
                const double z = opaque_zero();
#pragma unroll
                PDE_EACH { jet_fill(T[h], z); T[h].c[0] = cx[h].x0; T[h].c[1] = c_one; MT[h] = maj_leaf(cx[h].ax0, t0, 0.0f); }
            } break;
            case U_SETV1: {
                PDE_SPILL_IF_FLAGGED
                const double z = opaque_zero();
#pragma unroll
                PDE_EACH { jet_fill(T[h], z); T[h].c[0] = cx[h].x1; T[h].c[2] = c_one; MT[h] = maj_leaf(cx[h].ax1, t0, 0.0f); }
            } break;
            case U_SETC: {
                PDE_SPILL_IF_FLAGGED
                const double z = opaque_zero();
#pragma unroll
                PDE_EACH { jet_fill(T[h], z); T[h].c[0] = c_const[arg]; MT[h] = maj_leaf(c_constf[arg], 0.0f, c_constf[arg]); }
            } break;
            case U_SETP: {
                PDE_SPILL_IF_FLAGGED
                PDE_LOAD_P(T, MT)
            } break;
            // F(coordinate): the jet of F(x_k + dx_k) is F's Taylor expansion along dx_k
            case U_FNV: {
                PDE_SPILL_IF_FLAGGED
                const double z = opaque_zero();
#pragma unroll
                PDE_EACH {
                    const double x = arg ? cx[h].x1 : cx[h].x0;
                    scalar_taylor<N>((ins_cur >> 17) & 3u, ins_cur >> 19, x, f[h]);
                    MT[h] = maj_leaf(arg ? cx[h].ax1 : cx[h].ax0, t0, 0.0f);
                    maj_fn((ins_cur >> 17) & 3u, ins_cur >> 19, __double2float_rn(x), MT[h]);
                    jet_fill(T[h], z);
                    T[h].c[0] = f[h][0];
                    if (arg) {
#pragma unroll
                        for (int k = 1; k <= N; ++k) T[h].c[jidx(0, k)] = f[h][k];
                    } else {
#pragma unroll
                        for (int k = 1; k <= N; ++k) T[h].c[jidx(k, 0)] = f[h][k];
                    }
                }
            } break;
            case U_FNC: {
                PDE_SPILL_IF_FLAGGED
                const double z = opaque_zero();
#pragma unroll
                PDE_EACH {
                    scalar_taylor<N>((ins_cur >> 17) & 3u, ins_cur >> 19, c_const[arg], f[h]);
                    MT[h] = maj_leaf(c_constf[arg], 0.0f, c_constf[arg]);
                    maj_fn((ins_cur >> 17) & 3u, ins_cur >> 19, __double2float_rn(c_const[arg]), MT[h]);
                    jet_fill(T[h], z);
                    T[h].c[0] = f[h][0];
                }
            } break;
            case U_ADD_S: PDE_FETCH_S goto l_add;
            case U_ADD_P: PDE_FETCH_P
            l_add:
#pragma unroll
                PDE_EACH { maj_add(MT[h], MU[h]); jet_add(T[h], U[h]); }
                break;
            case U_SUB_S: PDE_FETCH_S goto l_sub;
            case U_SUB_P: PDE_FETCH_P
            l_sub:
#pragma unroll
                PDE_EACH { maj_add(MT[h], MU[h]); jet_sub(T[h], U[h]); }
                break;
            case U_RSUB_S: PDE_FETCH_S
#pragma unroll
                PDE_EACH { maj_add(MT[h], MU[h]); jet_rsub(T[h], U[h]); }
                break;
            case U_MUL_S: PDE_FETCH_S goto l_mul;
            case U_MUL_P: PDE_FETCH_P
            l_mul:
#pragma unroll
                PDE_EACH maj_mul(MT[h], MU[h]);
                jetv_mul<N, NP>(T, U);
                break;
            case U_DIV_S: PDE_FETCH_S goto l_div;
            case U_DIV_P: PDE_FETCH_P
            l_div:
#pragma unroll
                PDE_EACH MT[h] = maj_div(MT[h], MU[h], maj_abs(U[h].c[0]));
                jetv_div<N, NP>(T, U);
                break;
            // U / T: the division runs in place on the numerator U; the copy back is opaque to the
            // register allocator (jet_copy, jet.cuh) and costs 30 moves against 76 FP64 instructions
            case U_RDIV_S: PDE_FETCH_S goto l_rdiv;
            case U_RDIV_V0: PDE_U_VAR(x0, ax0, 1) goto l_rdiv;
            case U_RDIV_V1: PDE_U_VAR(x1, ax1, 2) goto l_rdiv;
            case U_RDIV_C: {
                const double z = opaque_zero();
#pragma unroll
                PDE_EACH { jet_fill(U[h], z); U[h].c[0] = c_const[arg]; MU[h] = maj_leaf(c_constf[arg], 0.0f, c_constf[arg]); }
            } goto l_rdiv;
            case U_RDIV_P: PDE_FETCH_P
            l_rdiv:
#pragma unroll
                PDE_EACH MT[h] = maj_div(MU[h], MT[h], maj_abs(T[h].c[0]));
                jetv_div<N, NP>(U, T);
#pragma unroll
                PDE_EACH jet_copy(T[h], U[h]);
                break;
            case U_ADDC:
#pragma unroll
                PDE_EACH { maj_add(MT[h], maj_leaf(c_constf[arg], 0.0f, c_constf[arg])); T[h].c[0] += c_const[arg]; }
                break;
            case U_SUBC:
#pragma unroll
                PDE_EACH { maj_add(MT[h], maj_leaf(c_constf[arg], 0.0f, c_constf[arg])); T[h].c[0] -= c_const[arg]; }
                break;
            case U_RSUBC:
#pragma unroll
                PDE_EACH { maj_add(MT[h], maj_leaf(c_constf[arg], 0.0f, c_constf[arg])); jet_neg(T[h]); T[h].c[0] += c_const[arg]; }
                break;
            // T * c and T / c (multiplication by the reciprocal): V' = |c| V, D' = |c| D, W' = |c| (W + 17 M) in both rules
            case U_MULC:
#pragma unroll
                PDE_EACH { const float c = c_constf[arg]; MT[h].W = maj_floor(c * fmaf(17.0f, MT[h].V + MT[h].D, MT[h].W)); MT[h].V = maj_floor(MT[h].V * c); MT[h].D *= c; jet_scale(T[h], c_const[arg]); }
                break;
            case U_MULRC:
#pragma unroll
                PDE_EACH { const float c = c_rconstf[arg]; MT[h].W = maj_floor(c * fmaf(17.0f, MT[h].V + MT[h].D, MT[h].W)); MT[h].V = maj_floor(MT[h].V * c); MT[h].D *= c; jet_scale(T[h], c_rconst[arg]); }
                break;
            case U_ADDV0:
#pragma unroll
                PDE_EACH { maj_add(MT[h], maj_leaf(cx[h].ax0, t0, 0.0f)); T[h].c[0] += cx[h].x0; T[h].c[1] += c_one; }
                break;
            case U_ADDV1:
#pragma unroll
                PDE_EACH { maj_add(MT[h], maj_leaf(cx[h].ax1, t0, 0.0f)); T[h].c[0] += cx[h].x1; T[h].c[2] += c_one; }
                break;
            case U_SUBV0:
#pragma unroll
                PDE_EACH { maj_add(MT[h], maj_leaf(cx[h].ax0, t0, 0.0f)); T[h].c[0] -= cx[h].x0; T[h].c[1] -= c_one; }
                break;
            case U_SUBV1:
#pragma unroll
                PDE_EACH { maj_add(MT[h], maj_leaf(cx[h].ax1, t0, 0.0f)); T[h].c[0] -= cx[h].x1; T[h].c[2] -= c_one; }
                break;
            case U_MULV0:
#pragma unroll
                PDE_EACH { maj_mul(MT[h], maj_leaf(cx[h].ax0, t0, 0.0f)); jet_mul_var(T[h], 0, cx[h].x0); }
                break;
            case U_MULV1:
#pragma unroll
                PDE_EACH { maj_mul(MT[h], maj_leaf(cx[h].ax1, t0, 0.0f)); jet_mul_var(T[h], 1, cx[h].x1); }
                break;
            case U_DIVV0:
#pragma unroll
                PDE_EACH { MT[h] = maj_div(MT[h], maj_leaf(cx[h].ax0, t0, 0.0f), cx[h].ax0); jet_div_var(T[h], 0, lazy_rcp(cx[h].x0)); }
                break;
            case U_DIVV1:
#pragma unroll
                PDE_EACH { MT[h] = maj_div(MT[h], maj_leaf(cx[h].ax1, t0, 0.0f), cx[h].ax1); jet_div_var(T[h], 1, lazy_rcp(cx[h].x1)); }
                break;
            case U_NEG:
#pragma unroll
                PDE_EACH jet_neg(T[h]);
                break;
            case U_ABS:
#pragma unroll
                PDE_EACH jet_abs(T[h]);
                break;
#define PDE_BY_AXIS(FN, ...)                                                 \
    if (ins_cur & F_UNI) {                                                      \
        if (ins_cur & F_AXIS1) FN<N, NP, 1>(__VA_ARGS__); else FN<N, NP, 0>(__VA_ARGS__); \
    } else {                                                                    \
        FN<N, NP, -1>(__VA_ARGS__);                                             \
    }
            case U_SQRT:
#pragma unroll
                PDE_EACH maj_sqrt(MT[h], maj_abs(T[h].c[0]));
                PDE_BY_AXIS(jetv_sqrt, T)
                break;
            case U_SQUARE:
#pragma unroll
                PDE_EACH maj_square(MT[h]);
                PDE_BY_AXIS(jetv_square, T)
                break;
            // scalar functions: Taylor coefficients of F at T_0 by one ratio recurrence, then the
            // shared in-place composition body
            case U_EXP:
#pragma unroll
                PDE_EACH {
                    scalar_taylor<N>(FN_EXP + arg, 0, T[h].c[0], f[h]);
                    maj_exp(MT[h], arg ? -1.0f : 1.0f, __double2float_rn(T[h].c[0]));
                }
                goto l_compose;
            case U_POW:
#pragma unroll
                PDE_EACH {
                    scalar_taylor<N>(FN_POW, arg, T[h].c[0], f[h]);
                    if (c_pow_int[arg] >= 0) maj_ipow(MT[h], c_powf[arg]); else maj_pow(MT[h], c_powf[arg], maj_abs(T[h].c[0]));
                }
            l_compose:
                PDE_BY_AXIS(jetv_compose_ps, T, U, f)
                break;
            default: __builtin_unreachable();
        }
    }
#undef PDE_BY_AXIS
#undef PDE_SPILL_IF_FLAGGED
#undef PDE_FETCH_S
#undef PDE_FETCH_P
#undef PDE_LOAD_P
#undef PDE_U_VAR
#undef PDE_EACH
}

// A load the compiler may not sink below the interpreter loop (it otherwise moves the table
// fetch next to its use and the residual stalls on the full L2 latency).
__device__ __forceinline__ double ldg_early(const double* p) {
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// Residual operators: R, the decision scale S~ and (SHARP: parity / tooling mode) the plain scale S from the finished
// jet.  S = sum of |monomial| bounds the round-off of evaluating R from the partials; S~ >= S also covers the round-off
// INSIDE the partials: it is the residual's majorant at partials inflated by theta_n = W c_theta[n-1] (oracle/majorant.py
// "Decision scale"), so |R| <= tau S~ for every float64 evaluation of an exact solution.
template <int PROBLEM> struct Residual;

// W of a finished jet as the double the thetas are built from (0 = exact: a coordinate under -x / |x|; every
// rounding rule clamps at kMajFloor, so an underflow cannot pose as exactness)
__device__ __forceinline__ double maj_final(float W) { return (double)W; }

template <> struct Residual<PDE_PROBLEM_FORCE_FREE> {
    static constexpr int N = 4;
    static constexpr int COLS = 1;
    // FFV:305-347; entries expanded by tools/gen_residual.py
    struct Coef { double c[1]; };
    __device__ static __forceinline__ void fetch(const double* tab, int P, int pt, Coef& k) { k.c[0] = ldg_early(tab + pt); }
    template <bool SHARP, int STRIDE>
    __device__ static __forceinline__ void eval(const Jet<4>& u, const Coef& k, double Wd, unsigned, double& R, double& St, double& S) {
        const double (&c)[1] = k.c;
        double d[15];
#pragma unroll
        for (int n = 0; n <= 4; ++n) {
#pragma unroll
            for (int j = 0; j <= n; ++j) d[jidx(n - j, j)] = u.c[jidx(n - j, j)] * (factorial(n - j) * factorial(j));
        }
        const double w = c[0];   // 1/rho
        double p[4], a[4];
        ff_residual_entries(d, w, p, a);
        // det M, FFV:347 -- the contraction is spelled out so that the run-time program of the same residual
        // (residual_programs.py, generated from the same schedule) is bit-identical
        R = fma(p[0], p[3], -(p[1] * p[2]));
        if (SHARP) S = fma(a[0], a[3], a[1] * a[2]);
        // isotropic majorant (oracle/majorant.py: iso_tables): every partial of order n replaced by
        // m_n = sum_{|g| = n} |d_g| + theta_n; one polynomial in m_1..m_4 and |w| instead of 48 monomials
        const double m1 = fma(Wd, c_theta[0], fabs(d[1]) + fabs(d[2]));
        const double m2 = fma(Wd, c_theta[1], (fabs(d[3]) + fabs(d[4])) + fabs(d[5]));
        const double m3 = fma(Wd, c_theta[2], (fabs(d[6]) + fabs(d[7])) + (fabs(d[8]) + fabs(d[9])));
        const double m4 = fma(Wd, c_theta[3], ((fabs(d[10]) + fabs(d[11])) + (fabs(d[12]) + fabs(d[13]))) + fabs(d[14]));
        const double aw = fabs(w), m1w = m1 * aw;
        const double b0 = m1 * fma(4.0, m3, aw * fma(2.0, m2, m1w));                                         // LT_A
        const double b1 = 8.0 * (m1 * m1) * m2;                                                               // LT_B
        const double b2 = m1 * fma(m2, fma(8.0, m3, 2.0 * (m2 * aw)),
                                   m1 * fma(8.0, m4, aw * fma(4.0, m3, aw * fma(4.0, m2, 2.0 * m1w))));       // L2T_A
        const double b3 = 8.0 * (m1 * m1) * fma(3.0 * m2, m2, 2.0 * (m1 * m3));                               // L2T_B
        St = fma(b0, b3, b1 * b2);
    }
};

template <> struct Residual<PDE_PROBLEM_KERR> {
    static constexpr int N = 2;
    static constexpr int COLS = 4;
    // KV:77-91 expanded: R = c1_r u_r + c1 u_rr + c2_x u_x + c2 u_xx
    struct Coef { double c[4]; };
    __device__ static __forceinline__ void fetch(const double* tab, int P, int pt, Coef& k) {
        k.c[0] = ldg_early(tab + pt); k.c[1] = ldg_early(tab + P + pt);
        k.c[2] = ldg_early(tab + 2 * (size_t)P + pt); k.c[3] = ldg_early(tab + 3 * (size_t)P + pt);
    }
    template <bool SHARP, int STRIDE>
    __device__ static __forceinline__ void eval(const Jet<2>& u, const Coef& k, double Wd, unsigned, double& R, double& St, double& S) {
        const double c1 = k.c[0], c1r = k.c[1], c2 = k.c[2], c2x = k.c[3];
        // products and sums by intrinsics: never contracted into FMAs, so the run-time program of the same residual
        // (residual_programs.py) is bit-identical
        const double t0 = __dmul_rn(c1r, u.c[1]);
        const double t1 = __dmul_rn(c1, 2.0 * u.c[3]);
        const double t2 = __dmul_rn(c2x, u.c[2]);
        const double t3 = __dmul_rn(c2, 2.0 * u.c[5]);
        R = __dadd_rn(__dadd_rn(t0, t1), __dadd_rn(t2, t3));
        const double s = __dadd_rn(__dadd_rn(fabs(t0), fabs(t1)), __dadd_rn(fabs(t2), fabs(t3)));
        if (SHARP) S = s;
        const double th1 = Wd * c_theta[0], th2 = Wd * c_theta[1];
        St = fma(fabs(c1r) + fabs(c2x), th1, fma(fabs(c1) + fabs(c2), th2, s));
    }
};

// Not a PDE: the value of u itself (order-2 jets are the cheapest instantiated interpreter).  Used by
// pde_fingerprint to bucket candidates by the FUNCTION they denote (GM:1256-1286).
constexpr int kProblemValue = 2;
template <> struct Residual<kProblemValue> {
    static constexpr int N = 2;
    static constexpr int COLS = 1;
    struct Coef {};
    __device__ static __forceinline__ void fetch(const double*, int, int, Coef&) {}
    template <bool SHARP, int STRIDE>
    __device__ static __forceinline__ void eval(const Jet<2>& u, const Coef&, double, unsigned, double& R, double& St, double& S) {
        R = u.c[0];
        St = S = fabs(u.c[1]) + fabs(u.c[2]);
    }
};

// ---------------------------------------------------------------------------------
// Run-time residual programs (BASELINE north_star item 3; the plugin seam is ProblemSpec.validator, PI:34-63, and a
// plugin's own `_lhs(u)`, KV:77-91): the residual operator of a problem arrives at run time as a straight-line scalar
// program over the finished partial derivatives of u -- pde_compile_residual_program, include/pde_b200.h; produced from
// the plugin's SymPy formula by pde_engine_b200/residual_compiler.py -- so a new plugin needs no CUDA and no rebuild.
//
// Machine: a scalar FILE F[0..n_file) of float64 per lane + one accumulator in a register.
//   F[0 .. NC)                 the partial derivatives d_g (jet index order, factorials applied)
//   F[NC .. NC + n_cols)       this point's row of the coefficient table (functions of the point only)
//   F[.. + n_consts)           the program's constants
//   above                      temporaries (written by MUL / STA before they are read: checked on the host)
// The file lives in the lane's SPILL COLUMN of shared memory ([element][thread], conflict free): when the residual
// runs the interpreter's spill stack is empty, so the column is free; the launcher sizes it (spill slots) for n_file.
// Instruction word: op | a << 4 | b << 12 | dst << 20 | neg << 28, fetched from the constant bank (warp-uniform).
// The program runs TWICE per point: on the values (R) and on magnitudes (|d_g| + theta_|g|, |c|, |k|, signs dropped)
// -- the residual's majorant at the round-off-inflated partials, i.e. the decision scale S~ of oracle/majorant.py in
// its non-isotropic form; with theta = 0 it is S = sum of |monomial|, the scale parity is quoted against.
// ---------------------------------------------------------------------------------
constexpr int kResMaxWords = PDE_R_MAX_WORDS;
constexpr int kResMaxConsts = PDE_R_MAX_CONSTS;
static __constant__ uint32_t c_res_words[kResMaxWords];
static __constant__ double c_res_consts[kResMaxConsts];
static __constant__ int c_res_dims[2];            // n_cols, n_consts

template <int STRIDE>
__device__ __forceinline__ double file_ld(unsigned base, unsigned k) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(base + k * (unsigned)STRIDE));
    return v;
}
template <int STRIDE>
__device__ __forceinline__ void file_st(unsigned base, unsigned k, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(base + k * (unsigned)STRIDE), "d"(v) : "memory");
}

// MAG: the magnitude pass (every sign dropped).  One copy per instantiation (__noinline__): the program is the same
// for the whole launch, the loop is warp-uniform.
template <bool MAG, int STRIDE>
__device__ __noinline__ double res_run(unsigned base) {
    double acc = 0.0;
#pragma unroll 1
    for (int pc = 0; pc < kResMaxWords; ++pc) {
        const unsigned w = c_res_words[pc];
        const unsigned op = w & 15u, a = (w >> 4) & 255u, b = (w >> 12) & 255u, d = (w >> 20) & 255u;
        const bool neg = !MAG && ((w >> 28) & 1u);
        if (op == PDE_R_OUT || op == PDE_R_END) break;
        if (op == PDE_R_STA) { file_st<STRIDE>(base, d, acc); continue; }
        double x = file_ld<STRIDE>(base, a);
        if (neg) x = -x;
        if (op == PDE_R_LDA) { acc = x; continue; }
        if (op == PDE_R_ADDA) { acc = __dadd_rn(acc, x); continue; }
        const double y = file_ld<STRIDE>(base, b);
        if (op == PDE_R_MUL) file_st<STRIDE>(base, d, __dmul_rn(x, y));
        else if (op == PDE_R_ACC0) acc = __dmul_rn(x, y);
        else acc = __fma_rn(x, y, acc);               // PDE_R_ACC
    }
    return acc;
}

template <int N_>
struct ResidualProg {
    static constexpr int N = N_;
    static constexpr int NC = Jet<N_>::NC;
    static constexpr int COLS = 0;                    // nothing prefetched: the row is read when the residual runs
    struct Coef { const double* row; int P; };
    __device__ static __forceinline__ void fetch(const double* tab, int P, int pt, Coef& k) { k.row = tab + pt; k.P = P; }
    template <bool SHARP, int STRIDE>
    __device__ static __forceinline__ void eval(const Jet<N_>& u, const Coef& k, double Wd, unsigned F, double& R, double& St, double& S) {
        const int nc = c_res_dims[0], nk = c_res_dims[1];
        double d[NC];
#pragma unroll
        for (int n = 0; n <= N; ++n) {
#pragma unroll
            for (int j = 0; j <= n; ++j) d[jidx(n - j, j)] = u.c[jidx(n - j, j)] * (factorial(n - j) * factorial(j));
        }
        // ---- values ----
#pragma unroll
        for (int g = 0; g < NC; ++g) file_st<STRIDE>(F, g, d[g]);
        for (int c = 0; c < nc; ++c) file_st<STRIDE>(F, NC + c, __ldg(k.row + (size_t)c * k.P));
        for (int c = 0; c < nk; ++c) file_st<STRIDE>(F, NC + nc + c, c_res_consts[c]);
        R = res_run<false, STRIDE>(F);
        // ---- magnitudes at the inflated partials: theta_n = W * 2 eps n! / (t0^n tau), theta_0 = theta_1 t0 ----
        double th[N + 1];
        th[0] = Wd * c_theta[0] * (double)c_t0;
#pragma unroll
        for (int n = 1; n <= N; ++n) th[n] = Wd * c_theta[n - 1];
#pragma unroll
        for (int n = 0; n <= N; ++n) {
#pragma unroll
            for (int j = 0; j <= n; ++j) file_st<STRIDE>(F, jidx(n - j, j), fabs(d[jidx(n - j, j)]) + th[n]);
        }
        for (int c = 0; c < nc + nk; ++c) file_st<STRIDE>(F, NC + c, fabs(file_ld<STRIDE>(F, NC + c)));
        St = res_run<true, STRIDE>(F);
        if (SHARP) {
            S = St;
            if (Wd != 0.0) {
#pragma unroll
                for (int g = 0; g < NC; ++g) file_st<STRIDE>(F, g, fabs(d[g]));
                S = res_run<true, STRIDE>(F);
            }
        }
    }
};
constexpr int kProblemProgram2 = 16 + 2, kProblemProgram4 = 16 + 4;     // template tags of ResidualProg<2>, <4>
template <> struct Residual<kProblemProgram2> : ResidualProg<2> {};
template <> struct Residual<kProblemProgram4> : ResidualProg<4> {};

// ---------------------------------------------------------------------------------
// Work decomposition.  Warps own candidates, lanes own collocation points.  A CTA is G groups
// of 4 warps -- one warp of a group on each of the SM's four schedulers (warp id mod 4).  A group
// takes a chunk of 4 consecutive candidates: each of its warps stages + translates one of them,
// then the 4 warps sweep the chunk's candidates one after the other (warp wg takes the 32-point
// stripes wg, wg+4, ...), so a candidate's micro-op stream is fetched by all four schedulers at
// the same time.  Different groups are at different candidates, hence the warps that SHARE a
// scheduler are in different phases (one dispatching or fetching an operand while another streams
// DFMAs): v5/v6 swept one candidate with all 16 warps in lockstep and the FP64 pipe idled whenever
// the whole SM was in a dispatch phase (66 % pipe-active, profiles/README.md).  Groups never
// synchronise with each other (named barriers, 128 threads); chunks are dealt round-robin.
// ---------------------------------------------------------------------------------
struct WarpPartial {
    double best_ratio, best_S, max_R;
    int n_fin, n_vote;
};

template <int N, int NP>
__host__ __device__ constexpr size_t cta_smem_bytes(int L, int ns, int W) {
    // per warp: code[L] | start[L] | need[L] | frames[2L] | ucode[2L+6] u32 ; then partials[W][4] ; status[W] ;
    // then spill[ns][(NC + 2) NP][32 W] 8-byte elements   (16-byte aligned pieces)
    return ((size_t)((L + 15) / 16 * 16) * 5 + (size_t)(kUcodeMax(L) * 4 + 15) / 16 * 16) * W +
           (size_t)W * 4 * sizeof(WarpPartial) + 16 * W +
           (size_t)ns * spill_slot_elems<N, NP>() * 32 * W * 8;
}

__device__ __forceinline__ void group_barrier(int g) {
    asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");
}

// MAJ = false: first pass, all P points, no round-off majorants (S~ = the residual's isotropic majorant at the
// computed partials): fast, but its rejections are only PROPOSALS.  MAJ = true with p.index: confirmation pass over the
// proposed rejections on the first P_eval points with the majorants carried; a rejection stands only if this pass
// votes it too, everything else gets its survivor bit back.  MAJ = true without p.index: one-pass mode / DUMP.
template <int PROBLEM, bool DUMP, int W, int NP, int MINB, bool MAJ>
__global__ void __launch_bounds__(W * 32, MINB)
validate_kernel(const ValidateParams p) {
    using Res = Residual<PROBLEM>;
    constexpr int N = Res::N;
    constexpr int NC = Jet<N>::NC;
    constexpr int TPB = W * 32;
    constexpr int G = W / 4;
    static_assert(W % 4 == 0, "a CTA is made of 4-warp groups");
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = warp >> 2, wg = warp & 3;
    const int Lp = (p.L + 15) / 16 * 16;
    const int ucb = (kUcodeMax(p.L) * 4 + 15) / 16 * 16;
    const size_t per_cand = (size_t)Lp * 5 + ucb;
    unsigned char* my = smem + per_cand * warp;
    uint8_t* s_code = my;
    uint32_t* s_uc_mine = reinterpret_cast<uint32_t*>(my + 5 * Lp);
    WarpPartial* s_part = reinterpret_cast<WarpPartial*>(smem + per_cand * W) + grp * 16;   // [cand slot][warp of group]
    int* s_status = reinterpret_cast<int*>(smem + per_cand * W + (size_t)W * 4 * sizeof(WarpPartial)) + grp * 4;
    double* s_spill = reinterpret_cast<double*>(smem + per_cand * W + (size_t)W * 4 * sizeof(WarpPartial) + 16 * W) + threadIdx.x;

    const unsigned spill_addr = keep_in_register(smem_addr(s_spill));
    const bool indexed = MAJ && p.is_confirm;              // confirmation pass: only the verdict is written
    const bool mapped = p.index != nullptr;
    const long long n_items = p.n_index ? (long long)*p.n_index : p.n;
    const long long n_chunks = (n_items + 3) / 4;
    // Chunks are dealt DYNAMICALLY (one atomicAdd per chunk by the group's first lane): candidates differ in cost by an
    // order of magnitude (2 to 60 micro-ops), and with a static round-robin the last groups of a small batch -- one
    // rank's shard of the depth-4 uniques is 6 chunks per group -- ran long after the others had finished.  Which
    // group evaluates a candidate never changes its outputs.
    int* s_next = reinterpret_cast<int*>(s_status) + 4 * (G - grp) + grp;          // [G] ints behind the status words
    for (long long chunk = (long long)blockIdx.x * G + grp;; chunk += (long long)gridDim.x * G) {
        if (p.chunk_counter) {
            if (wg == 0 && lane == 0) *s_next = (int)atomicAdd(p.chunk_counter, 1ULL);
            group_barrier(grp);
            chunk = *s_next;
        }
        if (chunk >= n_chunks) break;
        const long long cand0 = chunk * 4;
        // ---- phase 1: every warp of the group stages + translates its own candidate ----
        {
            const long long item = cand0 + wg;
            const long long cand = (mapped && item < n_items) ? (long long)p.index[item] : item;
            int status = -1;
            if (item < n_items) {
                int len = p.len[cand];
                if (len > p.L) len = p.L + 1;          // a length beyond the row is malformed input (never read past the row)
                const uint32_t* src = reinterpret_cast<const uint32_t*>(p.row_off ? p.code + (size_t)p.row_off[cand] * 16 : p.code + (size_t)cand * p.L);
                uint32_t* dst = reinterpret_cast<uint32_t*>(s_code);
                for (int i = lane; i * 4 < len && i * 4 < p.L; i += 32) dst[i] = __ldg(src + i);
                __syncwarp();
                if (lane == 0) {
                    status = (len == 0) ? -1 : (len > p.L) ? 1 : translate(s_code, len, s_uc_mine, my + Lp, my + 2 * Lp, my + 3 * Lp, p.ns, p.n_prim);
                    s_status[wg] = status;
                }
            } else if (lane == 0) {
                s_status[wg] = -9;   // no candidate in this slot
            }
        }
        group_barrier(grp);
        // ---- phase 2: the group's 4 warps sweep the chunk's candidates together ----
        for (int c = 0; c < 4; ++c) {
            const int status = s_status[c];
            if (status != 0) continue;
            const long long cand = mapped ? (long long)p.index[cand0 + c] : cand0 + c;
            const unsigned uc = keep_in_register(smem_addr(smem + per_cand * (grp * 4 + c) + 5 * Lp));
            int n_fin = 0, n_vote = 0;
            double best_ratio = 0.0, best_S = 0.0, max_R = 0.0;
            const size_t prim_stride = (size_t)p.P * 16;
#pragma unroll 1
            for (int stripe = wg * 32 * NP; stripe < p.P_eval; stripe += 128 * NP) {
                PointCtx<N> cx[NP];
                if (NP == 2) {
                    // one 128-bit load per coordinate: two consecutive points per lane
                    const double2 xa = __ldg(reinterpret_cast<const double2*>(p.pts + stripe) + lane);
                    const double2 xb = __ldg(reinterpret_cast<const double2*>(p.pts + p.P + stripe) + lane);
                    cx[0].x0 = xa.x; cx[0].x1 = xb.x; cx[NP - 1].x0 = xa.y; cx[NP - 1].x1 = xb.y;
                } else {
                    cx[0].x0 = __ldg(p.pts + stripe + lane);
                    cx[0].x1 = __ldg(p.pts + p.P + stripe + lane);
                }
                if (MAJ) {
#pragma unroll
                    for (int h = 0; h < NP; ++h) { cx[h].ax0 = maj_abs(cx[h].x0); cx[h].ax1 = maj_abs(cx[h].x1); }
                }
                typename Res::Coef coef[NP];
                int pt[NP];
#pragma unroll
                for (int h = 0; h < NP; ++h) {
                    pt[h] = stripe + NP * lane + h;
                    cx[h].prim = p.prim + (size_t)(pt[h] >> 5) * (16 * 32) + (pt[h] & 31);   // coalesced, immediate offsets
                    Res::fetch(p.tab, p.P, pt[h], coef[h]);      // issued early: latency hides behind the program
                }
                Jet<N> T[NP];
                Maj MT[NP];
                run_program<N, NP, TPB, MAJ>(uc, spill_addr, prim_stride, cx, T, MT);
#pragma unroll
                for (int h = 0; h < NP; ++h) {
                    double R, S, St;
                    Res::template eval<DUMP, TPB * 8>(T[h], coef[h], MAJ ? maj_final(MT[h].W) : 0.0, spill_addr, R, St, S);
                    if (DUMP) {
                        if (p.jets) {
#pragma unroll
                            for (int g = 0; g < NC; ++g) p.jets[((size_t)cand * NC + g) * p.P + pt[h]] = T[h].c[g];
                        }
                        if (p.resid) p.resid[(size_t)cand * p.P + pt[h]] = R;
                        if (p.scale) p.scale[(size_t)cand * p.P + pt[h]] = S;
                        if (p.scale_maj) p.scale_maj[(size_t)cand * p.P + pt[h]] = St;
                        if (p.maj) {
                            p.maj[((size_t)cand * 3 + 0) * p.P + pt[h]] = MT[h].V;
                            p.maj[((size_t)cand * 3 + 1) * p.P + pt[h]] = MT[h].D;
                            p.maj[((size_t)cand * 3 + 2) * p.P + pt[h]] = MT[h].W;
                        }
                    } else {
                        const double aR = fabs(R);
                        const bool fin = (aR <= 1.79769313486231570e308) && (St <= 1.79769313486231570e308) && (St > 0.0);
                        if (fin) {
                            ++n_fin;
                            const double ratio = aR * fast_rcp(St);
                            n_vote += (aR > p.tau * St) ? 1 : 0;
                            if (ratio > best_ratio) { best_ratio = ratio; best_S = St; }
                            max_R = fmax(max_R, aR);
                        }
                        if (p.ref_rs && pt[h] < p.n_ref) {
                            p.ref_rs[((size_t)cand * p.n_ref + pt[h]) * 2 + 0] = R;
                            p.ref_rs[((size_t)cand * p.n_ref + pt[h]) * 2 + 1] = St;
                        }
                    }
                }
            }
            if (!DUMP) {
                // ---- warp-shuffle reduction over the lanes, one row per (candidate, warp) ----
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    n_fin += __shfl_xor_sync(0xffffffffu, n_fin, off);
                    n_vote += __shfl_xor_sync(0xffffffffu, n_vote, off);
                    const double r2 = __shfl_xor_sync(0xffffffffu, best_ratio, off);
                    const double s2 = __shfl_xor_sync(0xffffffffu, best_S, off);
                    const double m2 = __shfl_xor_sync(0xffffffffu, max_R, off);
                    if (r2 > best_ratio) { best_ratio = r2; best_S = s2; }
                    max_R = fmax(max_R, m2);
                }
                if (lane == 0) {
                    WarpPartial wp;
                    wp.best_ratio = best_ratio; wp.best_S = best_S; wp.max_R = max_R; wp.n_fin = n_fin; wp.n_vote = n_vote;
                    s_part[c * 4 + wg] = wp;
                }
            }
        }
        group_barrier(grp);
        // ---- phase 3: every warp writes the row of the candidate it translated ----
        if (!DUMP && lane == 0) {
            const int c = wg;
            const long long item = cand0 + c;
            const int status = s_status[c];
            if (item < n_items) {
                const long long cand = mapped ? (long long)p.index[item] : item;
                WarpPartial t = s_part[c * 4];
                if (status == 0) {
                    for (int w = 1; w < 4; ++w) {
                        const WarpPartial o = s_part[c * 4 + w];
                        t.n_fin += o.n_fin; t.n_vote += o.n_vote;
                        if (o.best_ratio > t.best_ratio) { t.best_ratio = o.best_ratio; t.best_S = o.best_S; }
                        t.max_R = fmax(t.max_R, o.max_R);
                    }
                }
                const bool reject = status == 0 && (t.n_fin >= p.min_finite) && (t.n_vote > 0) && ((double)t.n_vote >= p.vote_frac * (double)t.n_fin);
                if (indexed) {
                    // confirmation pass: only the verdict (and its evidence) is written; the per-candidate maxima stay those
                    // of the first pass over the whole grid
                    if (p.confirm) { p.confirm[2 * cand] = status == 0 ? t.n_fin : -1 - status; p.confirm[2 * cand + 1] = status == 0 ? t.n_vote : 0; }
                } else if (status != 0) {
                    p.ratio_max[cand] = 0.0; p.resid_max[cand] = 0.0; p.scale_at[cand] = 0.0;
                    p.n_finite[cand] = (status < 0) ? -1 : -1 - status;   // -1 empty, -2 malformed, -3 spill overflow
                    p.n_votes[cand] = 0;
                    if (p.ref_rs) for (int k = 0; k < 2 * p.n_ref; ++k) p.ref_rs[(size_t)cand * 2 * p.n_ref + k] = __longlong_as_double(0x7ff8000000000000LL);
                } else {
                    p.ratio_max[cand] = t.best_ratio;
                    p.resid_max[cand] = t.max_R;
                    p.scale_at[cand] = t.best_S;
                    p.n_finite[cand] = t.n_fin;
                    p.n_votes[cand] = t.n_vote;
                }
                if (!reject) atomicOr(p.survivor_bits + (cand >> 5), 1u << (cand & 31));
            }
        }
        // no barrier here: a warp only overwrites its OWN status/ucode slot in the next phase 1, and the
        // partial rows of its candidate are rewritten only after the next phase-1 barrier
    }
}

}  // namespace pde
