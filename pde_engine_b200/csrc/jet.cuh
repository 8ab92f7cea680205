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
__constant__ double c_expc[16] = {
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

template <int N>
__device__ __forceinline__ void jet_set_const(Jet<N>& t, double v) {
#pragma unroll
    for (int g = 0; g < Jet<N>::NC; ++g) t.c[g] = 0.0;
    t.c[0] = v;
}

template <int N>
__device__ __forceinline__ void jet_set_var(Jet<N>& t, int k, double x) {
    jet_set_const(t, x);
    if (N >= 1) {
        t.c[1] = (k == 0) ? 1.0 : 0.0;
        t.c[2] = (k == 0) ? 0.0 : 1.0;
    }
}

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

// t = t * u   (in place: descending total degree; c_g only reads t_b with b <= g)
template <int N>
__device__ __forceinline__ void jet_mul(Jet<N>& t, const Jet<N>& u) {
#pragma unroll
    for (int n = N; n >= 0; --n) {
#pragma unroll
        for (int gj = 0; gj <= n; ++gj) {
            const int gi = n - gj;
            // two accumulators: halves the dependent-DFMA chain of the long sums
            double acc = t.c[jidx(gi, gj)] * u.c[0], acc1 = 0.0;
            int cnt = 0;
#pragma unroll
            for (int bi = 0; bi <= gi; ++bi) {
#pragma unroll
                for (int bj = 0; bj <= gj; ++bj) {
                    if (bi == gi && bj == gj) continue;
                    if ((cnt++ & 1) == 0) acc1 = fma(t.c[jidx(bi, bj)], u.c[jidx(gi - bi, gj - bj)], acc1);
                    else acc = fma(t.c[jidx(bi, bj)], u.c[jidx(gi - bi, gj - bj)], acc);
                }
            }
            t.c[jidx(gi, gj)] = cnt > 0 ? acc + acc1 : acc;
        }
    }
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

// t = t * t
template <int N>
__device__ __forceinline__ void jet_square(Jet<N>& t) {
#pragma unroll
    for (int n = N; n >= 0; --n) {
#pragma unroll
        for (int gj = 0; gj <= n; ++gj) {
            const int gi = n - gj;
            // pairs (b, g-b): count each unordered pair once, doubled
            double acc = 0.0;
            double mid = 0.0;
#pragma unroll
            for (int bi = 0; bi <= gi; ++bi) {
#pragma unroll
                for (int bj = 0; bj <= gj; ++bj) {
                    const int ci = gi - bi, cj = gj - bj;
                    const int ib = jidx(bi, bj), ic = jidx(ci, cj);
                    if (ib < ic) acc = fma(t.c[ib], t.c[ic], acc);
                    else if (ib == ic) mid = t.c[ib] * t.c[ib];
                }
            }
            t.c[jidx(gi, gj)] = fma(2.0, acc, mid);
        }
    }
}

// t = t / d   (in place on the numerator, ascending degree)
template <int N>
__device__ __forceinline__ void jet_div(Jet<N>& t, const Jet<N>& d) {
    const double r0 = 1.0 / d.c[0];
#pragma unroll
    for (int n = 0; n <= N; ++n) {
#pragma unroll
        for (int gj = 0; gj <= n; ++gj) {
            const int gi = n - gj;
            double acc = t.c[jidx(gi, gj)], acc1 = 0.0;
            int cnt = 0;
#pragma unroll
            for (int bi = 0; bi <= gi; ++bi) {
#pragma unroll
                for (int bj = 0; bj <= gj; ++bj) {
                    if (bi == 0 && bj == 0) continue;
                    if ((cnt++ & 1) == 0) acc = fma(-d.c[jidx(bi, bj)], t.c[jidx(gi - bi, gj - bj)], acc);
                    else acc1 = fma(-d.c[jidx(bi, bj)], t.c[jidx(gi - bi, gj - bj)], acc1);
                }
            }
            t.c[jidx(gi, gj)] = (cnt > 1 ? acc + acc1 : acc) * r0;
        }
    }
}

// t = num / t  for a scalar numerator: r_g = -r_0/num * ... ; computed into o
template <int N>
__device__ __forceinline__ void jet_inv(Jet<N>& o, const Jet<N>& t) {
    const double r0 = 1.0 / t.c[0];
    o.c[0] = r0;
#pragma unroll
    for (int n = 1; n <= N; ++n) {
#pragma unroll
        for (int gj = 0; gj <= n; ++gj) {
            const int gi = n - gj;
            double acc = 0.0;
#pragma unroll
            for (int bi = 0; bi <= gi; ++bi) {
#pragma unroll
                for (int bj = 0; bj <= gj; ++bj) {
                    if (bi == 0 && bj == 0) continue;
                    acc = fma(t.c[jidx(bi, bj)], o.c[jidx(gi - bi, gj - bj)], acc);
                }
            }
            o.c[jidx(gi, gj)] = -r0 * acc;
        }
    }
}

// t = sqrt(t)  (in place, ascending degree)
template <int N>
__device__ __forceinline__ void jet_sqrt(Jet<N>& t) {
    const double s0 = sqrt(t.c[0]);   // NaN for negative values: SymPy goes complex there
    const double h = 0.5 / s0;
    t.c[0] = s0;
#pragma unroll
    for (int n = 1; n <= N; ++n) {
#pragma unroll
        for (int gj = 0; gj <= n; ++gj) {
            const int gi = n - gj;
            double acc = 0.0, mid = 0.0;
#pragma unroll
            for (int bi = 0; bi <= gi; ++bi) {
#pragma unroll
                for (int bj = 0; bj <= gj; ++bj) {
                    const int ci = gi - bi, cj = gj - bj;
                    if ((bi == 0 && bj == 0) || (ci == 0 && cj == 0)) continue;
                    const int ib = jidx(bi, bj), ic = jidx(ci, cj);
                    if (ib < ic) acc = fma(t.c[ib], t.c[ic], acc);
                    else if (ib == ic) mid = t.c[ib] * t.c[ib];
                }
            }
            t.c[jidx(gi, gj)] = (t.c[jidx(gi, gj)] - fma(2.0, acc, mid)) * h;
        }
    }
}

// o = exp(t); t is clobbered (pre-scaled by |b|)
template <int N>
__device__ __forceinline__ void jet_exp(Jet<N>& o, Jet<N>& t) {
    o.c[0] = exp(t.c[0]);
#pragma unroll
    for (int n = 2; n <= N; ++n) {
#pragma unroll
        for (int j = 0; j <= n; ++j) t.c[jidx(n - j, j)] *= (double)n;
    }
#pragma unroll
    for (int n = 1; n <= N; ++n) {
#pragma unroll
        for (int gj = 0; gj <= n; ++gj) {
            const int gi = n - gj;
            double acc = 0.0, acc1 = 0.0;
            int cnt = 0;
#pragma unroll
            for (int bi = 0; bi <= gi; ++bi) {
#pragma unroll
                for (int bj = 0; bj <= gj; ++bj) {
                    if (bi == 0 && bj == 0) continue;
                    if ((cnt++ & 1) == 0) acc = fma(t.c[jidx(bi, bj)], o.c[jidx(gi - bi, gj - bj)], acc);
                    else acc1 = fma(t.c[jidx(bi, bj)], o.c[jidx(gi - bi, gj - bj)], acc1);
                }
            }
            o.c[jidx(gi, gj)] = (cnt > 1 ? acc + acc1 : acc) * (1.0 / (double)n);
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

// o = t ** k  (constant real exponent)
template <int N>
__device__ __forceinline__ void jet_pow(Jet<N>& o, const Jet<N>& t, double k) {
    o.c[0] = pow0(t.c[0], k);
    const double rb0 = 1.0 / t.c[0];
    const double k1 = k + 1.0;
#pragma unroll
    for (int n = 1; n <= N; ++n) {
#pragma unroll
        for (int gj = 0; gj <= n; ++gj) {
            const int gi = n - gj;
            double tot = 0.0;
            // group by m = |b|: coefficient ((k+1) m - n)
#pragma unroll
            for (int m = 1; m <= n; ++m) {
                double sm = 0.0;
#pragma unroll
                for (int bj = 0; bj <= m; ++bj) {
                    const int bi = m - bj;
                    if (bi > gi || bj > gj) continue;
                    sm = fma(t.c[jidx(bi, bj)], o.c[jidx(gi - bi, gj - bj)], sm);
                }
                tot = fma(k1 * (double)m - (double)n, sm, tot);
            }
            o.c[jidx(gi, gj)] = tot * (rb0 * (1.0 / (double)n));
        }
    }
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

template <int N, int AX>
__device__ __forceinline__ void jet_copy_ax(Jet<N>& t, const Jet<N>& u) {
#pragma unroll
    for (int n = 0; n <= N; ++n) {
#pragma unroll
        for (int j = 0; j <= n; ++j) {
            if (!ax_on<AX>(n - j, j)) continue;
            asm("mov.f64 %0, %1;" : "=d"(t.c[jidx(n - j, j)]) : "d"(u.c[jidx(n - j, j)]));
        }
    }
}

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

// o = exp(t); t is clobbered
template <int N, int NP>
__device__ __forceinline__ void jetv_exp(Jet<N> (&o)[NP], Jet<N> (&t)[NP], bool negate) {
    const double sg = negate ? -1.0 : 1.0;      // exp(-t): the sign rides on the pre-scaling
#pragma unroll
    PDE_H o[h].c[0] = exp(sg * t[h].c[0]);
#pragma unroll
    for (int n = 1; n <= N; ++n) {
#pragma unroll
        for (int j = 0; j <= n; ++j) {
#pragma unroll
            PDE_H t[h].c[jidx(n - j, j)] *= sg * (double)n;
        }
    }
#pragma unroll
    for (int n = 1; n <= N; ++n) {
#pragma unroll
        for (int gj = 0; gj <= n; ++gj) {
            const int gi = n - gj;
            double acc[NP], acc1[NP];
#pragma unroll
            PDE_H { acc[h] = 0.0; acc1[h] = 0.0; }
            int cnt = 0;
#pragma unroll
            for (int bi = 0; bi <= gi; ++bi) {
#pragma unroll
                for (int bj = 0; bj <= gj; ++bj) {
                    if (bi == 0 && bj == 0) continue;
                    if ((cnt++ & 1) == 0) {
#pragma unroll
                        PDE_H acc[h] = fma(t[h].c[jidx(bi, bj)], o[h].c[jidx(gi - bi, gj - bj)], acc[h]);
                    } else {
#pragma unroll
                        PDE_H acc1[h] = fma(t[h].c[jidx(bi, bj)], o[h].c[jidx(gi - bi, gj - bj)], acc1[h]);
                    }
                }
            }
#pragma unroll
            PDE_H o[h].c[jidx(gi, gj)] = (cnt > 1 ? acc[h] + acc1[h] : acc[h]) * (1.0 / (double)n);
        }
    }
}

// o = t ** k
template <int N, int NP>
__device__ __forceinline__ void jetv_pow(Jet<N> (&o)[NP], const Jet<N> (&t)[NP], double k) {
    double rb0[NP];
#pragma unroll
    PDE_H { o[h].c[0] = pow0(t[h].c[0], k); rb0[h] = fast_rcp(t[h].c[0]); }
    const double k1 = k + 1.0;
#pragma unroll
    for (int n = 1; n <= N; ++n) {
#pragma unroll
        for (int gj = 0; gj <= n; ++gj) {
            const int gi = n - gj;
            double tot[NP];
#pragma unroll
            PDE_H tot[h] = 0.0;
#pragma unroll
            for (int m = 1; m <= n; ++m) {
                double sm[NP];
#pragma unroll
                PDE_H sm[h] = 0.0;
#pragma unroll
                for (int bj = 0; bj <= m; ++bj) {
                    const int bi = m - bj;
                    if (bi > gi || bj > gj) continue;
#pragma unroll
                    PDE_H sm[h] = fma(t[h].c[jidx(bi, bj)], o[h].c[jidx(gi - bi, gj - bj)], sm[h]);
                }
                const double cm = k1 * (double)m - (double)n;
#pragma unroll
                PDE_H tot[h] = fma(cm, sm[h], tot[h]);
            }
#pragma unroll
            PDE_H o[h].c[jidx(gi, gj)] = tot[h] * (rb0[h] * (1.0 / (double)n));
        }
    }
}

// t = F(t) for a scalar function F given by its Taylor coefficients f[k] = F^(k)(t_0)/k! at the
// jet's value: Horner on delta = t - t_0, truncated at total degree N,
//     A <- f_N;   A <- f_k + delta * A  (k = N-1 .. 1, order N-k);   t <- f_0 + delta * A.
// Every level is computed in place in DESCENDING degree (level m only reads levels < m of the
// previous A), and the last level overwrites t itself (t_g reads t_b only for b <= g), so the
// result lands in t's own registers: no out-of-place body, no copy-back.  One body serves
// 1/x, x**k, exp(x) and exp(-x); `a` is scratch (the operand jet, dead during unary ops).
// N = 4: 2 + 9 + 25 + 55 = 91 multiply-adds.
// one Horner level: K > 0: a <- f_K + delta * a (order N-K, in place);  K == 0: t <- f_0 + delta * a.
// The constant term of `a` is never stored: at level K it is f_{K+1}.
template <int N, int NP, int K>
__device__ __forceinline__ void jetv_compose_level(Jet<N> (&t)[NP], Jet<N> (&a)[NP], const double (&f)[NP][N + 1]) {
#pragma unroll
    for (int m = N - K; m >= 1; --m) {
#pragma unroll
        for (int gj = 0; gj <= m; ++gj) {
            const int gi = m - gj;
            double acc[NP];
            int cnt = 0;
#pragma unroll
            for (int bi = 0; bi <= gi; ++bi) {
#pragma unroll
                for (int bj = 0; bj <= gj; ++bj) {
                    if (bi == 0 && bj == 0) continue;
                    const int ib = jidx(bi, bj), ic = jidx(gi - bi, gj - bj);
                    if (cnt == 0) {
#pragma unroll
                        PDE_H acc[h] = t[h].c[ib] * (ic == 0 ? f[h][K + 1] : a[h].c[ic]);
                    } else {
#pragma unroll
                        PDE_H acc[h] = fma(t[h].c[ib], ic == 0 ? f[h][K + 1] : a[h].c[ic], acc[h]);
                    }
                    ++cnt;
                }
            }
#pragma unroll
            PDE_H {
                if (K > 0) a[h].c[jidx(gi, gj)] = acc[h];
                else t[h].c[jidx(gi, gj)] = acc[h];
            }
        }
    }
    if constexpr (K > 0) jetv_compose_level<N, NP, K - 1>(t, a, f);
    else {
#pragma unroll
        PDE_H t[h].c[0] = f[h][0];
    }
}

template <int N, int NP>
__device__ __forceinline__ void jetv_compose(Jet<N> (&t)[NP], Jet<N> (&a)[NP], const double (&f)[NP][N + 1]) {
    jetv_compose_level<N, NP, N - 1>(t, a, f);
}
// Paterson-Stockmeyer form of the same composition:  F = f_0 + f_1 d + d^2 (f_2 + f_3 d + f_4 d^2),  d = t - t_0:
// TWO truncated jet products (d^2 with its symmetry, then d^2 * G) instead of Horner's three nested ones, and the
// scalar-times-jet parts share their scalar between consecutive multiply-adds:  23 + 8 + 35 + 14 = 80 multiply-adds
// for N = 4 (Horner 91), about 50 of them with three different register pairs (Horner 61).  In place on t;
// `a` holds d^2 (degrees 2..N) and, in its degree-1 slots, the degree-1 part of G.   Valid for N <= 4.
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
