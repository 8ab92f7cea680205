// jet.cuh -- register-resident FP64 Taylor-mode jets, two variables, order N.
//
// Generalises the reference's 2nd-order forward-mode rules
// (problems/force_free/validator.py:70-180: leaf rules 78-83, Add 102-110,
// Mul/Leibniz 113-129, Pow 132-141, sqrt/exp 151-163) to order N.
//
// A jet holds NORMALISED Taylor coefficients c[idx(i,j)] = d0^i d1^j f/(i! j!),
// idx(i,j) = n(n+1)/2 + j, n = i+j, so a product is a plain truncated
// convolution.  Every loop below has compile-time bounds and is fully
// unrolled: all indices are static and the jets live in registers.
//
// Recurrences (D = radial Euler operator, D c_g = |g| c_g):
//   mul   c_g = sum_{b<=g} a_b b_{g-b}                         (in place, descending |g|)
//   div   q_g = (a_g - sum_{b!=0} d_b q_{g-b}) / d_0           (in place on the numerator)
//   sqrt  s_g = (a_g - sum_{0<b<g} s_b s_{g-b}) / (2 s_0)       (in place)
//   exp   |g| e_g = sum_{b!=0} |b| a_b e_{g-b}
//   pow   |g| a_0 p_g = sum_{b!=0} (k|b| - |g-b|) a_b p_{g-b}
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace pde {

__host__ __device__ constexpr int jidx(int i, int j) { return (i + j) * (i + j + 1) / 2 + j; }

// Branch-free reciprocal / reciprocal square root: MUFU seed (about 2^-23) + two Newton steps.
// The IEEE-exact library sequences carry a slow-path CALL/branch per use; here 0, inf and
// subnormal inputs simply come out non-finite (NaN instead of inf), which the validator
// treats identically (the point is not counted).
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
// y = 1/sqrt(x); sqrt(x) = x * y refined once more by the caller
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    y = y * fma(-hx * y, y, 1.5);
    y = y * fma(-hx * y, y, 1.5);
    return y;
}
__device__ __forceinline__ double fast_sqrt(double x, double& half_rsqrt) {
    const double y = fast_rsqrt(x);
    double s = x * y;
    half_rsqrt = 0.5 * y;
    s = fma(fma(-s, s, x), half_rsqrt, s);
    return s;
}

// exp(x), branch free, constants as constant-bank operands (the library exp materialises its
// eleven 64-bit coefficients with two UMOVs each and carries a slow-path branch):
// n = rint(x log2 e) by the 2^52+2^51 trick, r = x - n ln2 (two-term Cody-Waite), Taylor to degree 13
// on |r| <= ln2/2 (remainder 4e-18 relative), scaled by 2^n in two halves so that overflow gives inf
// and underflow is gradual.  NaN propagates; |x| > 750 is clamped first (exp is 0 / inf there).
static __constant__ double c_expc[16] = {
    1.0, 1.0, 1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880,
    1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0,
    6.93147180369123816490e-01, 1.90821492927058770002e-10};     // ln2 hi / lo
__device__ __forceinline__ double fast_exp(double x) {
    const double xc = fmin(fmax(x, -750.0), 750.0);
    const double magic = 6755399441055744.0;
    const double nf = fma(xc, 1.4426950408889634074, magic);
    const int n = __double2loint(nf);
    const double nr = nf - magic;
    double r = fma(-nr, c_expc[14], xc);
    r = fma(-nr, c_expc[15], r);
    double p = c_expc[13];
#pragma unroll
    for (int k = 12; k >= 0; --k) p = fma(p, r, c_expc[k]);
    const int n1 = n >> 1, n2 = n - n1;
    p *= __hiloint2double((n1 + 1023) << 20, 0);
    p *= __hiloint2double((n2 + 1023) << 20, 0);
    return (x != x) ? x : p;
}

template <int N>
struct Jet {
    static constexpr int NC = (N + 1) * (N + 2) / 2;
    double c[NC];
};

// The copy is an opaque asm mov on purpose: a plain assignment lets ptxas alias the two
// jets and then re-shuffle all 30 registers at the head of the interpreter loop on EVERY
// micro-op (38 % of all executed instructions were IMAD.MOV, profiles/r1_v2_*).
template <int N>
__device__ __forceinline__ void jet_copy(Jet<N>& t, const Jet<N>& u) {
#pragma unroll
    for (int g = 0; g < Jet<N>::NC; ++g) asm("mov.f64 %0, %1;" : "=d"(t.c[g]) : "d"(u.c[g]));
}

template <int N>
__device__ __forceinline__ void jet_add(Jet<N>& t, const Jet<N>& u) {
#pragma unroll
    for (int g = 0; g < Jet<N>::NC; ++g) t.c[g] += u.c[g];
}

template <int N>  // t = t - u
__device__ __forceinline__ void jet_sub(Jet<N>& t, const Jet<N>& u) {
#pragma unroll
    for (int g = 0; g < Jet<N>::NC; ++g) t.c[g] -= u.c[g];
}

template <int N>  // t = u - t
__device__ __forceinline__ void jet_rsub(Jet<N>& t, const Jet<N>& u) {
#pragma unroll
    for (int g = 0; g < Jet<N>::NC; ++g) t.c[g] = u.c[g] - t.c[g];
}

template <int N>
__device__ __forceinline__ void jet_scale(Jet<N>& t, double s) {
#pragma unroll
    for (int g = 0; g < Jet<N>::NC; ++g) t.c[g] *= s;
}

// sign flip on the high word: an integer-pipe LOP3 per coefficient instead of a half-rate FP64 op
template <int N>
__device__ __forceinline__ void jet_neg(Jet<N>& t) {
#pragma unroll
    for (int g = 0; g < Jet<N>::NC; ++g)
        t.c[g] = __hiloint2double(__double2hiint(t.c[g]) ^ (int)0x80000000, __double2loint(t.c[g]));
}

// t = t * (x_k + dx_k): multiply by a coordinate (2 non-zero coefficients)
template <int N>
__device__ __forceinline__ void jet_mul_var(Jet<N>& t, int k, double x) {
#pragma unroll
    for (int n = N; n >= 0; --n) {
#pragma unroll
        for (int gj = 0; gj <= n; ++gj) {
            const int gi = n - gj;
            double acc = t.c[jidx(gi, gj)] * x;
            if (k == 0) {
                if (gi > 0) acc += t.c[jidx(gi - 1, gj)];
            } else {
                if (gj > 0) acc += t.c[jidx(gi, gj - 1)];
            }
            t.c[jidx(gi, gj)] = acc;
        }
    }
}

// t = t / (x_k + dx_k); r = 1 / x_k
template <int N>
__device__ __forceinline__ void jet_div_var(Jet<N>& t, int k, double r) {
#pragma unroll
    for (int n = 0; n <= N; ++n) {
#pragma unroll
        for (int gj = 0; gj <= n; ++gj) {
            const int gi = n - gj;
            double acc = t.c[jidx(gi, gj)];
            if (k == 0) {
                if (gi > 0) acc -= t.c[jidx(gi - 1, gj)];
            } else {
                if (gj > 0) acc -= t.c[jidx(gi, gj - 1)];
            }
            t.c[jidx(gi, gj)] = acc * r;
        }
    }
}

// value of b0 ** k in REAL arithmetic (NaN where SymPy's principal value is complex)
__device__ __forceinline__ double pow0(double b0, double k) {
    const double ak = fabs(k);
    const double ik = rint(ak);
    if (ik == ak && ak <= 16.0) {          // integer exponent: exact products, any sign of b0
        double r = 1.0, base = b0;
        int e = (int)ik;
#pragma unroll 1
        while (e) {
            if (e & 1) r *= base;
            base *= base;
            e >>= 1;
        }
        return k < 0 ? fast_rcp(r) : r;
    }
    const double a2 = 2.0 * ak;
    if (rint(a2) == a2 && a2 <= 32.0) {    // half-integer: sqrt(b0)^(2k)
        double unused_h;
        const double s = fast_sqrt(b0, unused_h);
        double r = 1.0, base = s;
        int e = (int)a2;
#pragma unroll 1
        while (e) {
            if (e & 1) r *= base;
            base *= base;
            e >>= 1;
        }
        return k < 0 ? fast_rcp(r) : r;
    }
    return b0 >= 0.0 ? pow(b0, k) : __longlong_as_double(0x7ff8000000000000LL);
}

// t = |t|   (not differentiable at 0: derivatives become NaN there)
template <int N>
__device__ __forceinline__ void jet_abs(Jet<N>& t) {
    const double v = t.c[0];
    const double s = v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : __longlong_as_double(0x7ff8000000000000LL));
#pragma unroll
    for (int g = 1; g < Jet<N>::NC; ++g) t.c[g] *= s;
    t.c[0] = fabs(v);
}


// Univariate jets.  A sub-expression that depends on ONE coordinate only has a jet whose mixed and other-axis
// coefficients are structurally zero: coefficient (i, j) is on axis AX = 0 iff j == 0, on AX = 1 iff i == 0
// (AX < 0: the general bivariate jet).  The bodies below take AX as a template parameter and skip every
// output and every product that involves an off-axis coefficient -- the loops are fully unrolled, so the test is a
// compile-time constant.  An order-4 product costs 15 multiply-adds instead of 70 on an axis.  The skipped
// coefficients are never read and never written: they stay the zeros the leaf put there.
template <int AX>
__host__ __device__ constexpr bool ax_on(int i, int j) { return AX < 0 || (AX == 0 ? j == 0 : i == 0); }

// ---------------------------------------------------------------------------------
// NP-point versions of the long bodies: the point index h is the INNERMOST loop, so the
// NP independent dependency chains are interleaved in program order (NP-way ILP for
// the in-order issue of one warp; DFMA dependent latency is 8.2 cycles, the pipe takes
// one warp-DFMA every 2 cycles -- tools/microbench/dfma_latency.cu).
// ---------------------------------------------------------------------------------
#define PDE_H for (int h = 0; h < NP; ++h)

// One accumulator chain per output coefficient: the 15 outputs are independent, which is all the
// ILP the in-order issue needs, and every extra instruction costs an issue slot (a second chain
// per output added 14 DADDs to the 70 multiply-adds).
template <int N, int NP, int AX = -1>
__device__ __forceinline__ void jetv_mul(Jet<N> (&t)[NP], const Jet<N> (&u)[NP]) {
#pragma unroll
    for (int n = N; n >= 0; --n) {
#pragma unroll
        for (int gj = 0; gj <= n; ++gj) {
            const int gi = n - gj;
            if (!ax_on<AX>(gi, gj)) continue;
            double acc[NP];
#pragma unroll
            PDE_H acc[h] = t[h].c[jidx(gi, gj)] * u[h].c[0];
#pragma unroll
            for (int bi = 0; bi <= gi; ++bi) {
#pragma unroll
                for (int bj = 0; bj <= gj; ++bj) {
                    if (bi == gi && bj == gj) continue;
                    if (!ax_on<AX>(bi, bj)) continue;
#pragma unroll
                    PDE_H acc[h] = fma(t[h].c[jidx(bi, bj)], u[h].c[jidx(gi - bi, gj - bj)], acc[h]);
                }
            }
#pragma unroll
            PDE_H t[h].c[jidx(gi, gj)] = acc[h];
        }
    }
}

// t = t / d (in place on the numerator)
template <int N, int NP, int AX = -1>
__device__ __forceinline__ void jetv_div(Jet<N> (&t)[NP], const Jet<N> (&d)[NP]) {
    double r0[NP];
#pragma unroll
    PDE_H r0[h] = fast_rcp(d[h].c[0]);
#pragma unroll
    for (int n = 0; n <= N; ++n) {
#pragma unroll
        for (int gj = 0; gj <= n; ++gj) {
            const int gi = n - gj;
            if (!ax_on<AX>(gi, gj)) continue;
            double acc[NP];
#pragma unroll
            PDE_H acc[h] = t[h].c[jidx(gi, gj)];
#pragma unroll
            for (int bi = 0; bi <= gi; ++bi) {
#pragma unroll
                for (int bj = 0; bj <= gj; ++bj) {
                    if (bi == 0 && bj == 0) continue;
                    if (!ax_on<AX>(bi, bj)) continue;
#pragma unroll
                    PDE_H acc[h] = fma(-d[h].c[jidx(bi, bj)], t[h].c[jidx(gi - bi, gj - bj)], acc[h]);
                }
            }
#pragma unroll
            PDE_H t[h].c[jidx(gi, gj)] = acc[h] * r0[h];
        }
    }
}

template <int N, int NP, int AX = -1>
__device__ __forceinline__ void jetv_square(Jet<N> (&t)[NP]) {
#pragma unroll
    for (int n = N; n >= 0; --n) {
#pragma unroll
        for (int gj = 0; gj <= n; ++gj) {
            const int gi = n - gj;
            if (!ax_on<AX>(gi, gj)) continue;
            double acc[NP], mid[NP];
#pragma unroll
            PDE_H { acc[h] = 0.0; mid[h] = 0.0; }
#pragma unroll
            for (int bi = 0; bi <= gi; ++bi) {
#pragma unroll
                for (int bj = 0; bj <= gj; ++bj) {
                    const int ib = jidx(bi, bj), ic = jidx(gi - bi, gj - bj);
                    if (!ax_on<AX>(bi, bj)) continue;
                    if (ib < ic) {
#pragma unroll
                        PDE_H acc[h] = fma(t[h].c[ib], t[h].c[ic], acc[h]);
                    } else if (ib == ic) {
#pragma unroll
                        PDE_H mid[h] = t[h].c[ib] * t[h].c[ib];
                    }
                }
            }
#pragma unroll
            PDE_H t[h].c[jidx(gi, gj)] = fma(2.0, acc[h], mid[h]);
        }
    }
}

template <int N, int NP, int AX = -1>
__device__ __forceinline__ void jetv_sqrt(Jet<N> (&t)[NP]) {
    double hh[NP];
#pragma unroll
    PDE_H { const double s0 = fast_sqrt(t[h].c[0], hh[h]); t[h].c[0] = s0; }
#pragma unroll
    for (int n = 1; n <= N; ++n) {
#pragma unroll
        for (int gj = 0; gj <= n; ++gj) {
            const int gi = n - gj;
            if (!ax_on<AX>(gi, gj)) continue;
            double acc[NP], mid[NP];
#pragma unroll
            PDE_H { acc[h] = 0.0; mid[h] = 0.0; }
#pragma unroll
            for (int bi = 0; bi <= gi; ++bi) {
#pragma unroll
                for (int bj = 0; bj <= gj; ++bj) {
                    const int ci = gi - bi, cj = gj - bj;
                    if ((bi == 0 && bj == 0) || (ci == 0 && cj == 0)) continue;
                    if (!ax_on<AX>(bi, bj)) continue;
                    const int ib = jidx(bi, bj), ic = jidx(ci, cj);
                    if (ib < ic) {
#pragma unroll
                        PDE_H acc[h] = fma(t[h].c[ib], t[h].c[ic], acc[h]);
                    } else if (ib == ic) {
#pragma unroll
                        PDE_H mid[h] = t[h].c[ib] * t[h].c[ib];
                    }
                }
            }
#pragma unroll
            PDE_H t[h].c[jidx(gi, gj)] = (t[h].c[jidx(gi, gj)] - fma(2.0, acc[h], mid[h])) * hh[h];
        }
    }
}

// t = F(t) for a scalar function F given by its Taylor coefficients f[k] = F^(k)(t_0)/k! at the jet's value (one body
// serves 1/x, x**k, exp(x) and exp(-x)), in Paterson-Stockmeyer form:  F = f_0 + f_1 d + d^2 (f_2 + f_3 d + f_4 d^2),
// d = t - t_0: TWO truncated jet products (d^2 with its symmetry, then d^2 * G) instead of a Horner scheme's three nested
// ones (91 multiply-adds, measured 138.3 vs 136.1 ms per 10^6 trees in round 1), and the scalar-times-jet parts share
// their scalar between consecutive multiply-adds:  23 + 8 + 35 + 14 = 80 multiply-adds for N = 4.  In place on t;
// `a` is scratch (the operand jet, dead during unary ops): it holds d^2 (degrees 2..N) and, in its degree-1 slots,
// the degree-1 part of G.   Valid for N <= 4.
template <int N, int NP, int AX = -1>
__device__ __forceinline__ void jetv_compose_ps(Jet<N> (&t)[NP], Jet<N> (&a)[NP], const double (&f)[NP][N + 1]) {
    static_assert(N <= 4, "G = f_2 + f_3 d + f_4 d^2 covers N <= 4");
    // ---- d^2, degrees 2..N: unordered pairs once, doubled, plus the square of the middle term ----
#pragma unroll
    for (int n = 2; n <= N; ++n) {
#pragma unroll
        for (int gj = 0; gj <= n; ++gj) {
            const int gi = n - gj;
            if (!ax_on<AX>(gi, gj)) continue;
            double acc[NP], sq[NP];
            int cnt = 0;
            bool has_sq = false;
#pragma unroll
            PDE_H { acc[h] = 0.0; sq[h] = 0.0; }
#pragma unroll
            for (int bi = 0; bi <= gi; ++bi) {
#pragma unroll
                for (int bj = 0; bj <= gj; ++bj) {
                    const int ci = gi - bi, cj = gj - bj;
                    if ((bi == 0 && bj == 0) || (ci == 0 && cj == 0)) continue;
                    const int ib = jidx(bi, bj), ic = jidx(ci, cj);
                    if (ib > ic) continue;
                    if (ib == ic) {
                        has_sq = true;
#pragma unroll
                        PDE_H sq[h] = t[h].c[ib] * t[h].c[ib];
                    } else if (cnt++ == 0) {
#pragma unroll
                        PDE_H acc[h] = t[h].c[ib] * t[h].c[ic];
                    } else {
#pragma unroll
                        PDE_H acc[h] = fma(t[h].c[ib], t[h].c[ic], acc[h]);
                    }
                }
            }
#pragma unroll
            PDE_H a[h].c[jidx(gi, gj)] = cnt == 0 ? sq[h] : has_sq ? fma(2.0, acc[h], sq[h]) : acc[h] + acc[h];
        }
    }
    // ---- G = f_2 + f_3 d + f_4 d^2, truncated at degree N - 2: g1 (degree 1) in a.c[1..2], g2 (degree 2) in locals ----
    double g2[NP][3];
    if (N >= 3) {
#pragma unroll
        PDE_H {
            if (ax_on<AX>(1, 0)) a[h].c[1] = f[h][3] * t[h].c[1];
            if (ax_on<AX>(0, 1)) a[h].c[2] = f[h][3] * t[h].c[2];
        }
    }
    if (N >= 4) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (!ax_on<AX>(2 - k, k)) continue;
#pragma unroll
            PDE_H g2[h][k] = fma(f[h][4], a[h].c[3 + k], f[h][3] * t[h].c[3 + k]);
        }
    }
    // ---- t_g = f_1 t_g + sum_{|b| >= 2} d2_b G_{g-b}  (descending degree; G_0 = f_2) ----
#pragma unroll
    for (int n = N; n >= 2; --n) {
#pragma unroll
        for (int gj = 0; gj <= n; ++gj) {
            const int gi = n - gj, g = jidx(gi, gj);
            if (!ax_on<AX>(gi, gj)) continue;
            double acc[NP];
#pragma unroll
            PDE_H acc[h] = fma(f[h][2], a[h].c[g], f[h][1] * t[h].c[g]);
#pragma unroll
            for (int ci = 0; ci <= gi; ++ci) {
#pragma unroll
                for (int cj = 0; cj <= gj; ++cj) {
                    const int m = ci + cj;                 // degree of the G factor
                    if (m == 0 || m > N - 2 || n - m < 2) continue;
                    const int ib = jidx(gi - ci, gj - cj);
#pragma unroll
                    PDE_H acc[h] = fma(a[h].c[ib], m == 1 ? a[h].c[jidx(ci, cj)] : g2[h][cj], acc[h]);
                }
            }
#pragma unroll
            PDE_H t[h].c[g] = acc[h];
        }
    }
#pragma unroll
    PDE_H {
        if (N >= 1 && ax_on<AX>(1, 0)) t[h].c[1] *= f[h][1];
        if (N >= 1 && ax_on<AX>(0, 1)) t[h].c[2] *= f[h][1];
        t[h].c[0] = f[h][0];
    }
}

#undef PDE_H

__host__ __device__ __forceinline__ constexpr double factorial(int n) {
    return n <= 1 ? 1.0 : n == 2 ? 2.0 : n == 3 ? 6.0 : n == 4 ? 24.0 : n == 5 ? 120.0 : 720.0;
}

}  // namespace pde
