// validate.cuh -- the batched jet interpreter (stage 2 of the hot path).
//
// Replaces the per-candidate validator.validate() call of emit_to_db
// (general_method_paper_reproduction.py:1302-1316) and the validator worker
// pool (GM:1672-1824) as a numerical FILTER in front of the symbolic check.
//
// Mapping (BASELINE.json north_star): one WARP owns a candidate, each LANE owns
// collocation points (two per 64-point stripe, fetched with one 128-bit load
// per coordinate).  Per candidate the warp
//   1. stages the postfix bytecode in shared memory,
//   2. translates it once into leaf-fused micro-ops (lane 0),
//   3. for every point stripe interprets the micro-ops on a register-resident
//      top-of-stack jet T (+ operand jet U); deeper stack entries spill to a
//      per-lane column in shared memory (conflict free),
//   4. applies the problem's residual operator to the finished jet,
//   5. reduces votes / maxima over the lanes with warp shuffles.
#pragma once
#include <stdint.h>
#include "jet.cuh"
#include "residual_ff_gen.cuh"
#include "../../include/pde_b200.h"

namespace pde {

#ifndef PDE_FF_MINBLOCKS
#define PDE_FF_MINBLOCKS 4
#endif
constexpr int kWarpsPerBlock = 4;
constexpr int kMaxL = 256;

// per-launch tables (warp-uniform reads -> constant cache)
__constant__ double c_const[PDE_N_CONST];
__constant__ double c_rconst[PDE_N_CONST];   // reciprocals (division by a constant leaf)
__constant__ double c_pow[PDE_N_POW];

// Micro-ops.  Every arithmetic body exists exactly ONCE in the kernel (operands are
// brought into the operand jet U by separate micro-ops) so the interpreter's code
// stays inside the 32 KB instruction cache: an earlier version that inlined the
// bodies per call site ran at a 75 % i-cache hit rate (profiles/r1_v1_*).
enum UKind : uint8_t {
    U_END = 0,
    U_SPILL,                          // S[sp++] = T
    U_MOVTU,                          // T = U
    U_SETU_C, U_SETU_V0, U_SETU_V1,   // U = const / coordinate jet (one shared body)
    U_LOADU_P, U_LOADU_S,             // U = primitive jet table / S[--sp]
    U_ADD, U_SUB, U_RSUB, U_MUL, U_DIV, U_RDIV,   // T = T op U;  RSUB: U - T, RDIV: U / T
    U_ADDC, U_SUBC, U_MULC,           // sparse leaf fast paths (arg = const slot; MULC bit7 = reciprocal)
    U_ADDV0, U_ADDV1, U_SUBV0, U_SUBV1, U_MULV0, U_MULV1, U_DIVV0, U_DIVV1,
    U_NEG, U_ABS, U_SQRT, U_EXP, U_SQUARE, U_POW
};

struct ValidateParams {
    const uint8_t* code;
    const uint8_t* len;
    long long n;
    int L;
    const double* pts;    // [2][P]
    const double* tab;    // [cols][P]
    const double* prim;   // [n_prim][NC][P]
    int P;
    int ns;               // spill slots per lane
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
    double* scale;
};

__device__ __forceinline__ bool op_is_leaf(unsigned b) {
    return b == PDE_OP_VAR0 || b == PDE_OP_VAR1 || (b >= PDE_OP_PRIM0 && b < PDE_OP_PRIM0 + PDE_N_PRIM) || b >= PDE_OP_CONST0;
}
__device__ __forceinline__ bool op_is_binary(unsigned b) { return b >= PDE_OP_ADD && b <= PDE_OP_DIV; }
__device__ __forceinline__ bool op_is_unary(unsigned b) {
    return (b >= PDE_OP_NEG && b <= PDE_OP_EXP) || (b >= PDE_OP_FN_NEG && b <= PDE_OP_FN_EXPNEG) ||
           (b >= PDE_OP_POW0 && b < PDE_OP_POW0 + PDE_N_POW);
}

constexpr uint8_t V_JET_T = 0x03;  // virtual-stack markers (unused opcode values)
constexpr uint8_t V_JET_S = 0x04;

__host__ __device__ constexpr int kUcodeMax(int L) { return 3 * L + 4; }

// Postfix bytecode -> micro-ops.  Returns 0 ok, 1 malformed, 2 spill overflow.
// Invariant: the top-most jet of the virtual stack is always T; older jets are
// spilled in stack order, leaves never occupy a jet.
__device__ __noinline__ int translate(const uint8_t* code, int len, uint16_t* uc, uint8_t* vst, int ns_max) {
    int sp = 0, nu = 0, ns = 0, tpos = -1;
    auto emit = [&](unsigned kind, unsigned arg) { uc[nu++] = (uint16_t)((kind << 8) | arg); };
    auto set_u = [&](unsigned leaf) {   // U = jet of a leaf
        if (leaf >= PDE_OP_CONST0) emit(U_SETU_C, leaf - PDE_OP_CONST0);
        else if (leaf == PDE_OP_VAR0) emit(U_SETU_V0, 0);
        else if (leaf == PDE_OP_VAR1) emit(U_SETU_V1, 0);
        else emit(U_LOADU_P, leaf - PDE_OP_PRIM0);
    };
    auto spill_t = [&]() -> bool {
        if (tpos < 0) return true;
        if (ns >= ns_max) return false;
        emit(U_SPILL, 0); vst[tpos] = V_JET_S; ++ns;
        return true;
    };
    auto emit_inv = [&]() { emit(U_SETU_C, 0); emit(U_RDIV, 0); };   // 1 / T  (CONST(0) = 1)
    // T = T op leaf (leaf on the right)
    auto bin_leaf_right = [&](unsigned o, unsigned leaf) {
        if (leaf >= PDE_OP_CONST0) {
            const unsigned k = leaf - PDE_OP_CONST0;
            emit(o == 0 ? U_ADDC : o == 1 ? U_SUBC : U_MULC, o == 3 ? (k | 0x80u) : k);
        } else if (leaf == PDE_OP_VAR0 || leaf == PDE_OP_VAR1) {
            const unsigned v = leaf - PDE_OP_VAR0;
            emit((o == 0 ? U_ADDV0 : o == 1 ? U_SUBV0 : o == 2 ? U_MULV0 : U_DIVV0) + v, 0);
        } else {
            set_u(leaf);
            emit(o == 0 ? U_ADD : o == 1 ? U_SUB : o == 2 ? U_MUL : U_DIV, 0);
        }
    };
    for (int pc = 0; pc < len; ++pc) {
        const unsigned b = code[pc];
        if (op_is_leaf(b)) {
            vst[sp++] = (uint8_t)b;
        } else if (op_is_unary(b)) {
            if (sp < 1) return 1;
            const unsigned top = vst[sp - 1];
            if (top != V_JET_T) {
                if (top == V_JET_S) return 1;
                if (!spill_t()) return 2;
                set_u(top); emit(U_MOVTU, 0);
                vst[sp - 1] = V_JET_T; tpos = sp - 1;
            }
            switch (b) {
                case PDE_OP_NEG: case PDE_OP_FN_NEG: emit(U_NEG, 0); break;
                case PDE_OP_ABS: emit(U_ABS, 0); break;
                case PDE_OP_SQRT: emit(U_SQRT, 0); break;
                case PDE_OP_EXP: emit(U_EXP, 0); break;
                case PDE_OP_FN_INV: emit_inv(); break;
                case PDE_OP_FN_SQUARE: emit(U_SQUARE, 0); break;
                case PDE_OP_FN_POW32: emit(U_POW, 0); break;
                case PDE_OP_FN_POWN32: emit(U_POW, 1); break;
                case PDE_OP_FN_EXPNEG: emit(U_NEG, 0); emit(U_EXP, 0); break;
                default: {
                    const unsigned slot = b - PDE_OP_POW0;
                    const double k = c_pow[slot];
                    if (k == 2.0) emit(U_SQUARE, 0);
                    else if (k == 0.5) emit(U_SQRT, 0);
                    else if (k == -1.0) emit_inv();
                    else emit(U_POW, slot);
                }
            }
        } else if (op_is_binary(b)) {
            if (sp < 2) return 1;
            const unsigned bb = vst[sp - 1], aa = vst[sp - 2];
            sp -= 2;
            const unsigned o = b - PDE_OP_ADD;  // 0 add 1 sub 2 mul 3 div
            if (aa == V_JET_S && bb == V_JET_T) {
                emit(U_LOADU_S, 0); --ns;
                emit(o == 0 ? U_ADD : o == 1 ? U_RSUB : o == 2 ? U_MUL : U_RDIV, 0);     // U op T
            } else if (aa == V_JET_T && bb != V_JET_S) {
                bin_leaf_right(o, bb);
            } else if (bb == V_JET_T && aa != V_JET_S) {
                if (o == 0 || o == 2) bin_leaf_right(o, aa);                  // commutative
                else if (o == 1) { emit(U_NEG, 0); bin_leaf_right(0, aa); }   // leaf - T = -T + leaf
                else { set_u(aa); emit(U_RDIV, 0); }      // leaf / T
            } else if (aa != V_JET_S && bb != V_JET_S && aa != V_JET_T && bb != V_JET_T) {
                if (!spill_t()) return 2;
                set_u(aa); emit(U_MOVTU, 0);
                bin_leaf_right(o, bb);
            } else {
                return 1;
            }
            vst[sp] = V_JET_T; tpos = sp; ++sp;
        } else {
            return 1;
        }
    }
    if (sp != 1) return 1;
    if (vst[0] != V_JET_T) { set_u(vst[0]); emit(U_MOVTU, 0); }
    emit(U_END, 0);
    return 0;
}

template <int N>
struct PointCtx {
    double x0, x1;
    int pt;
    int P;
    const double* prim;
};

// Interpret the micro-ops for one point: result in T.
template <int N>
__device__ __forceinline__ void run_program(const uint16_t* __restrict__ uc, double* __restrict__ spill,
                                            const PointCtx<N>& cx, Jet<N>& T) {
    constexpr int NC = Jet<N>::NC;
    Jet<N> U;
    int sp = 0;  // spill depth
    int pc = 0;
    unsigned ins = uc[0];
#pragma unroll 1
    for (;;) {
        const unsigned kind = ins >> 8, arg = ins & 0xff;
        ins = uc[++pc];            // prefetch the next micro-op behind this one's body
        switch (kind) {
            case U_END: return;
            case U_SPILL: {
                double* dst = spill + (size_t)sp * NC * 32;
#pragma unroll
                for (int g = 0; g < NC; ++g) dst[g * 32] = T.c[g];
                ++sp;
            } break;
            case U_SETU_C: case U_SETU_V0: case U_SETU_V1: {
                const double v = kind == U_SETU_C ? c_const[arg] : kind == U_SETU_V0 ? cx.x0 : cx.x1;
                jet_set_const(U, v);
                U.c[1] = kind == U_SETU_V0 ? 1.0 : 0.0;
                U.c[2] = kind == U_SETU_V1 ? 1.0 : 0.0;
            } break;
            case U_LOADU_P: {
                const double* src = cx.prim + (size_t)arg * NC * cx.P + cx.pt;
#pragma unroll
                for (int g = 0; g < NC; ++g) U.c[g] = __ldg(src + (size_t)g * cx.P);
            } break;
            case U_LOADU_S: {
                --sp;
                const double* src = spill + (size_t)sp * NC * 32;
#pragma unroll
                for (int g = 0; g < NC; ++g) U.c[g] = src[g * 32];
            } break;
            case U_ADD: jet_add(T, U); break;
            case U_SUB: jet_sub(T, U); break;
            case U_RSUB: jet_rsub(T, U); break;
            case U_MUL: jet_mul(T, U); break;
            case U_DIV: jet_div(T, U); break;
            case U_ADDC: T.c[0] += c_const[arg]; break;
            case U_SUBC: T.c[0] -= c_const[arg]; break;
            case U_MULC: jet_scale(T, (arg & 0x80u) ? c_rconst[arg & 0x7fu] : c_const[arg]); break;
            case U_ADDV0: T.c[0] += cx.x0; T.c[1] += 1.0; break;
            case U_ADDV1: T.c[0] += cx.x1; T.c[2] += 1.0; break;
            case U_SUBV0: T.c[0] -= cx.x0; T.c[1] -= 1.0; break;
            case U_SUBV1: T.c[0] -= cx.x1; T.c[2] -= 1.0; break;
            case U_MULV0: jet_mul_var(T, 0, cx.x0); break;
            case U_MULV1: jet_mul_var(T, 1, cx.x1); break;
            case U_DIVV0: jet_div_var(T, 0, cx.x0); break;
            case U_DIVV1: jet_div_var(T, 1, cx.x1); break;
            case U_NEG: jet_neg(T); break;
            case U_ABS: jet_abs(T); break;
            case U_SQRT: jet_sqrt(T); break;
            case U_SQUARE: jet_square(T); break;
            // out-of-place bodies compute into U and copy back (jet_copy is opaque to the
            // register allocator, see jet.cuh)
            case U_RDIV: jet_div(U, T); jet_copy(T, U); break;
            case U_EXP: jet_exp(U, T); jet_copy(T, U); break;
            case U_POW: jet_pow(U, T, c_pow[arg]); jet_copy(T, U); break;
            case U_MOVTU: jet_copy(T, U); break;
            default: return;
        }
    }
}

// Residual operators: R and its round-off scale S from the finished jet.
template <int PROBLEM> struct Residual;

template <> struct Residual<PDE_PROBLEM_FORCE_FREE> {
    static constexpr int N = 4;
    static constexpr int COLS = 1;
    // FFV:305-347; entries expanded by tools/gen_residual.py
    __device__ static __forceinline__ void eval(const Jet<4>& u, const double* tab, int P, int pt, double& R, double& S) {
        double d[15];
#pragma unroll
        for (int n = 0; n <= 4; ++n) {
#pragma unroll
            for (int j = 0; j <= n; ++j) d[jidx(n - j, j)] = u.c[jidx(n - j, j)] * (factorial(n - j) * factorial(j));
        }
        const double w = __ldg(tab + pt);   // 1/rho
        double p[4], a[4];
        ff_residual_entries(d, w, p, a);
        R = p[0] * p[3] - p[1] * p[2];       // det M, FFV:347
        S = a[0] * a[3] + a[1] * a[2];
    }
};

template <> struct Residual<PDE_PROBLEM_KERR> {
    static constexpr int N = 2;
    static constexpr int COLS = 4;
    // KV:77-91 expanded: R = c1_r u_r + c1 u_rr + c2_x u_x + c2 u_xx
    __device__ static __forceinline__ void eval(const Jet<2>& u, const double* tab, int P, int pt, double& R, double& S) {
        const double c1 = __ldg(tab + pt), c1r = __ldg(tab + P + pt);
        const double c2 = __ldg(tab + 2 * (size_t)P + pt), c2x = __ldg(tab + 3 * (size_t)P + pt);
        const double t0 = c1r * u.c[1];
        const double t1 = c1 * (2.0 * u.c[3]);
        const double t2 = c2x * u.c[2];
        const double t3 = c2 * (2.0 * u.c[5]);
        R = (t0 + t1) + (t2 + t3);
        S = (fabs(t0) + fabs(t1)) + (fabs(t2) + fabs(t3));
    }
};

template <int N>
__host__ __device__ constexpr size_t warp_smem_bytes(int L, int ns) {
    // code[L] | vstack[L] | ucode[3L+4] u16 | spill[ns][NC][32] f64   (16-byte aligned pieces)
    return (size_t)((L + 15) / 16 * 16) * 2 + (size_t)(kUcodeMax(L) * 2 + 15) / 16 * 16 + (size_t)ns * Jet<N>::NC * 32 * 8;
}

template <int PROBLEM, bool DUMP>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, PROBLEM == PDE_PROBLEM_FORCE_FREE ? PDE_FF_MINBLOCKS : 5)
validate_kernel(const ValidateParams p) {
    using Res = Residual<PROBLEM>;
    constexpr int N = Res::N;
    constexpr int NC = Jet<N>::NC;
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t wbytes = warp_smem_bytes<N>(p.L, p.ns);
    unsigned char* base = smem + wbytes * warp;
    const int Lp = (p.L + 15) / 16 * 16;
    uint8_t* s_code = base;
    uint8_t* s_vst = base + Lp;
    uint16_t* s_uc = reinterpret_cast<uint16_t*>(base + 2 * Lp);
    double* s_spill = reinterpret_cast<double*>(base + 2 * Lp + (kUcodeMax(p.L) * 2 + 15) / 16 * 16) + lane;

    const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
    for (long long cand = (long long)blockIdx.x * kWarpsPerBlock + warp; cand < p.n; cand += nwarps) {
        const int len = p.len[cand];
        // ---- stage bytecode (coalesced 4-byte loads) ----
        {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(p.code + (size_t)cand * p.L);
            uint32_t* dst = reinterpret_cast<uint32_t*>(s_code);
            for (int i = lane; i * 4 < len; i += 32) dst[i] = __ldg(src + i);
        }
        __syncwarp();
        int status = 0;
        if (lane == 0) status = (len == 0) ? -1 : translate(s_code, len, s_uc, s_vst, p.ns);
        status = __shfl_sync(0xffffffffu, status, 0);
        if (status != 0) {
            if (!DUMP && lane == 0) {
                p.ratio_max[cand] = 0.0; p.resid_max[cand] = 0.0; p.scale_at[cand] = 0.0;
                p.n_finite[cand] = (status < 0) ? -1 : -1 - status;   // -1 empty, -2 malformed, -3 spill overflow
                p.n_votes[cand] = 0;
                if (p.ref_rs) for (int k = 0; k < 2 * p.n_ref; ++k) p.ref_rs[(size_t)cand * 2 * p.n_ref + k] = __longlong_as_double(0x7ff8000000000000LL);
                atomicOr(p.survivor_bits + (cand >> 5), 1u << (cand & 31));
            }
            __syncwarp();
            continue;
        }
        int n_fin = 0, n_vote = 0;
        double best_ratio = 0.0, best_S = 0.0, max_R = 0.0;
#pragma unroll 1
        for (int stripe = 0; stripe < p.P; stripe += 64) {
            // one 128-bit load per coordinate: two consecutive points per lane
            const double2 xa = __ldg(reinterpret_cast<const double2*>(p.pts + stripe) + lane);
            const double2 xb = __ldg(reinterpret_cast<const double2*>(p.pts + p.P + stripe) + lane);
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                PointCtx<N> cx;
                cx.x0 = h ? xa.y : xa.x;
                cx.x1 = h ? xb.y : xb.x;
                cx.pt = stripe + 2 * lane + h;
                cx.P = p.P;
                cx.prim = p.prim;
                Jet<N> T;
                run_program<N>(s_uc, s_spill, cx, T);
                double R, S;
                Res::eval(T, p.tab, p.P, cx.pt, R, S);
                if (DUMP) {
                    if (p.jets) {
#pragma unroll
                        for (int g = 0; g < NC; ++g) p.jets[((size_t)cand * NC + g) * p.P + cx.pt] = T.c[g];
                    }
                    if (p.resid) p.resid[(size_t)cand * p.P + cx.pt] = R;
                    if (p.scale) p.scale[(size_t)cand * p.P + cx.pt] = S;
                } else {
                    const double aR = fabs(R);
                    const bool fin = (aR <= 1.79769313486231570e308) && (S <= 1.79769313486231570e308) && (S > 0.0);
                    if (fin) {
                        ++n_fin;
                        const double ratio = aR / S;
                        n_vote += (aR > p.tau * S) ? 1 : 0;
                        if (ratio > best_ratio) { best_ratio = ratio; best_S = S; }
                        max_R = fmax(max_R, aR);
                    }
                    if (p.ref_rs && cx.pt < p.n_ref) {
                        p.ref_rs[((size_t)cand * p.n_ref + cx.pt) * 2 + 0] = R;
                        p.ref_rs[((size_t)cand * p.n_ref + cx.pt) * 2 + 1] = S;
                    }
                }
            }
        }
        if (!DUMP) {
            // ---- warp-shuffle reduction over the lanes ----
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
                p.ratio_max[cand] = best_ratio;
                p.resid_max[cand] = max_R;
                p.scale_at[cand] = best_S;
                p.n_finite[cand] = n_fin;
                p.n_votes[cand] = n_vote;
                const bool reject = (n_fin >= p.min_finite) && (n_vote > 0) && ((double)n_vote >= p.vote_frac * (double)n_fin);
                if (!reject) atomicOr(p.survivor_bits + (cand >> 5), 1u << (cand & 31));
            }
        }
        __syncwarp();
    }
}

}  // namespace pde
