// pde_b200.cu -- C-ABI glue: errors, sessions, residual programs, stage-2 launchers.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <atomic>
#include <cmath>
#include <mutex>
#include <vector>

#include "common.h"
#include "validate.cuh"
#include "launch.cuh"

namespace pde {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(int err, const char* what) {
    set_error("CUDA error %d (%s) in %s", err, cudaGetErrorString((cudaError_t)err), what);
    return PDE_E_CUDA;
}

void count_launch(int n) { g_launches += n; }

// ---- library-owned, stream-ordered scratch pool (one per device) ----
static std::mutex g_pool_mu;
static cudaMemPool_t g_pools[64] = {nullptr};

int scratch_alloc(void** ptr, size_t bytes, void* stream) {
    int dev = 0;
    PDE_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("device index %d out of range", dev); return PDE_E_INVALID; }
    cudaMemPool_t pool;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (!g_pools[dev]) {
            cudaMemPoolProps props{};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            PDE_CUDA(cudaMemPoolCreate(&g_pools[dev], &props));
            unsigned long long keep = ~0ULL;                    // never give memory back between calls
            PDE_CUDA(cudaMemPoolSetAttribute(g_pools[dev], cudaMemPoolAttrReleaseThreshold, &keep));
        }
        pool = g_pools[dev];
    }
    PDE_CUDA(cudaMallocFromPoolAsync(ptr, bytes ? bytes : 1, pool, (cudaStream_t)stream));
    return PDE_OK;
}

void scratch_free(void* ptr, void* stream) {
    if (ptr) cudaFreeAsync(ptr, (cudaStream_t)stream);
}

void exprset_mark_use(const pde_exprset* e, void* stream) {
    pde_exprset* em = const_cast<pde_exprset*>(e);
    if (!em->used_event) {
        cudaEvent_t ev = nullptr;
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return; }
        em->used_event = ev;
    }
    cudaEventRecord((cudaEvent_t)em->used_event, (cudaStream_t)stream);
}

bool have_device() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return false; }
    return n > 0;
}

// ---- FP64 pipe microbenchmark: 8 independent DFMA chains per thread ----
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// The same with THREE different register pairs per DFMA (acc = fma(y, z, acc)), the operand pattern of a
// jet convolution: the register file delivers two new 64-bit operands per DFMA slot, so this stream issues
// every 3 cycles per scheduler instead of 2 (tools/microbench/dfma_operands.cu) -- the ceiling that applies
// to sums of products of two different jets, unless the operand-reuse cache serves one of the factors.
__global__ void __launch_bounds__(256) fp64_peak3_kernel(double* out, int iters, double a, double b) {
    double x[8], y[8], z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x + i; y[i] = a + 1e-9 * (threadIdx.x + i); z[i] = a - 1e-9 * (3 * threadIdx.x + i); }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fma(y[(i + k) & 7], z[(i + 2 * k + 1) & 7], x[i]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + y[i] + z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + b;
}

// ---- function fingerprints ----
// One warp per candidate: every value u(x_k) is rounded to (52 - drop_bits) mantissa bits (round half up on the
// magnitude; -0 -> +0), salted with its point index and summed -- an order-independent 64-bit key, so
// candidates that denote the same function get the same key unless a value sits within round-off of a
// rounding boundary (probability ~ 2^-drop_bits-ish per value; a split bucket only costs a redundant
// CPU simplify, it never merges different functions).  Non-finite values hash as one NaN pattern.
__device__ __forceinline__ unsigned long long fp_mix64(unsigned long long x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27; x *= 0x94D049BB133111EBULL;
    x ^= x >> 31;
    return x;
}

__global__ void __launch_bounds__(256)
fingerprint_kernel(const double* __restrict__ values, long long n, int P, int drop_bits,
                   unsigned long long* __restrict__ key, int* __restrict__ n_finite) {
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long cand = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); cand < n; cand += warps) {
        unsigned long long h = 0;
        int nf = 0;
        for (int k = lane; k < P; k += 32) {
            const unsigned long long b = (unsigned long long)__double_as_longlong(values[(size_t)cand * P + k]);
            const unsigned long long mag = b & 0x7fffffffffffffffULL;
            unsigned long long q = 0x7ff8000000000000ULL;
            if (mag < 0x7ff0000000000000ULL) {
                ++nf;
                q = ((mag + (1ULL << (drop_bits - 1))) >> drop_bits) << drop_bits;
                if (q) q |= b & 0x8000000000000000ULL;
            }
            h += fp_mix64(q ^ (0x9E3779B97F4A7C15ULL * (unsigned long long)(k + 1)));
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            h += __shfl_xor_sync(0xffffffffu, h, off);
            nf += __shfl_xor_sync(0xffffffffu, nf, off);
        }
        if (lane == 0) {
            unsigned long long k64 = fp_mix64(h);
            key[cand] = nf ? (k64 ? k64 : 1ULL) : 0ULL;      // 0 = no finite value: unknown, leave to the CPU
            n_finite[cand] = nf;
        }
    }
}

}  // namespace pde

using namespace pde;

extern "C" {

int pde_abi_version(void) { return PDE_B200_ABI_VERSION; }
const char* pde_last_error(void) { return g_err; }
int64_t pde_launch_count(void) { return (int64_t)g_launches.load(); }

int pde_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ---------------------------------------------------------------- sessions
int pde_session_create(const char* var0, const char* var1, const char* const* const_names,
                       const double* const_vals, int n_named, pde_session** out) {
    if (!var0 || !var1 || !out || n_named < 0) { set_error("pde_session_create: bad argument"); return PDE_E_INVALID; }
    pde_session* s = new pde_session();
    s->var[0] = var0;
    s->var[1] = var1;
    bool has_e = false;
    for (int i = 0; i < n_named; ++i) {
        s->named.push_back(const_names[i]);
        s->named_vals.push_back(const_vals[i]);
        if (s->named.back() == "E") has_e = true;
    }
    if (!has_e) { s->named.push_back("E"); s->named_vals.push_back(2.718281828459045); }
    s->const_keys = {"1"};
    s->const_vals = {1.0};
    s->pow_keys = {"3/2", "-3/2", "2"};
    s->pow_vals = {1.5, -1.5, 2.0};
    *out = s;
    return PDE_OK;
}

void pde_session_free(pde_session* s) { delete s; }

int pde_session_tables(const pde_session* s, double* const_vals, int* n_const, double* pow_vals, int* n_pow) {
    if (!s) { set_error("null session"); return PDE_E_INVALID; }
    if (const_vals) {
        memset(const_vals, 0, sizeof(double) * PDE_N_CONST);
        memcpy(const_vals, s->const_vals.data(), sizeof(double) * s->const_vals.size());
    }
    if (pow_vals) {
        memset(pow_vals, 0, sizeof(double) * PDE_N_POW);
        memcpy(pow_vals, s->pow_vals.data(), sizeof(double) * s->pow_vals.size());
    }
    if (n_const) *n_const = (int)s->const_vals.size();
    if (n_pow) *n_pow = (int)s->pow_vals.size();
    return PDE_OK;
}

const char* pde_session_const_key(const pde_session* s, int k) {
    if (!s || k < 0 || k >= (int)s->const_keys.size()) return nullptr;
    return s->const_keys[k].c_str();
}
const char* pde_session_pow_key(const pde_session* s, int k) {
    if (!s || k < 0 || k >= (int)s->pow_keys.size()) return nullptr;
    return s->pow_keys[k].c_str();
}

// ------------------------------------------------------- residual programs
int pde_compile_residual(int problem_id, const double* consts, int n_consts, pde_program** out) {
    if (!out) { set_error("null out"); return PDE_E_INVALID; }
    pde_program* p = new pde_program();
    p->problem = problem_id;
    if (problem_id == PDE_PROBLEM_FORCE_FREE) {
        p->order = 4; p->cols = 1;
    } else if (problem_id == PDE_PROBLEM_KERR) {
        if (n_consts < 2 || !consts) { delete p; set_error("Kerr residual needs consts (M, a)"); return PDE_E_INVALID; }
        p->order = 2; p->cols = 4;
        p->consts[0] = consts[0]; p->consts[1] = consts[1];
    } else {
        delete p;
        set_error("unknown problem id %d", problem_id);
        return PDE_E_INVALID;
    }
    p->n_coef = (p->order + 1) * (p->order + 2) / 2;
    *out = p;
    return PDE_OK;
}

int pde_compile_residual_program(int jet_order, int n_cols, const double* consts, int n_consts,
                                 const uint32_t* words, int n_words, pde_program** out) {
    if (!out || !words || n_words < 1 || (n_consts > 0 && !consts)) { set_error("pde_compile_residual_program: bad argument"); return PDE_E_INVALID; }
    if (jet_order != 2 && jet_order != 4) { set_error("jet_order must be 2 or 4 (the instantiated jet algebras)"); return PDE_E_INVALID; }
    if (n_cols < 0 || n_cols > PDE_R_MAX_COLS || n_consts < 0 || n_consts > PDE_R_MAX_CONSTS || n_words > PDE_R_MAX_WORDS) {
        set_error("residual program exceeds the limits (cols %d/%d, consts %d/%d, words %d/%d)", n_cols, PDE_R_MAX_COLS, n_consts, PDE_R_MAX_CONSTS, n_words, PDE_R_MAX_WORDS);
        return PDE_E_OVERFLOW;
    }
    const int nc = (jet_order + 1) * (jet_order + 2) / 2, first_tmp = nc + n_cols + n_consts;
    // abstract run: operands are inputs or temporaries written earlier; dst is a temporary; the program ends with OUT
    std::vector<char> defined(256, 0);
    for (int k = 0; k < first_tmp; ++k) defined[k] = 1;
    int n_file = first_tmp;
    bool acc_set = false, ended = false;
    for (int i = 0; i < n_words; ++i) {
        const uint32_t w = words[i];
        const unsigned op = w & 15u, a = (w >> 4) & 255u, b = (w >> 12) & 255u, d = (w >> 20) & 255u;
        if (ended) { set_error("residual program: word %d follows OUT", i); return PDE_E_INVALID; }
        auto need = [&](unsigned k) { return k < 256 && defined[k]; };
        auto dst_ok = [&](unsigned k) { return (int)k >= first_tmp && k < PDE_R_MAX_FILE; };
        bool ok = true;
        switch (op) {
            case PDE_R_MUL: ok = need(a) && need(b) && dst_ok(d); if (ok) { defined[d] = 1; if ((int)d + 1 > n_file) n_file = d + 1; } break;
            case PDE_R_ACC0: ok = need(a) && need(b); acc_set = true; break;
            case PDE_R_ACC: ok = need(a) && need(b) && acc_set; break;
            case PDE_R_LDA: ok = need(a); acc_set = true; break;
            case PDE_R_ADDA: ok = need(a) && acc_set; break;
            case PDE_R_STA: ok = acc_set && dst_ok(d); if (ok) { defined[d] = 1; if ((int)d + 1 > n_file) n_file = d + 1; } break;
            case PDE_R_OUT: ok = acc_set; ended = true; break;
            default: ok = false;
        }
        if (!ok) { set_error("residual program: word %d (op %u a %u b %u dst %u) reads an undefined entry, writes an input, or has no accumulator", i, op, a, b, d); return PDE_E_INVALID; }
    }
    if (!ended) { set_error("residual program does not end with OUT"); return PDE_E_INVALID; }
    pde_program* p = new pde_program();
    p->problem = PDE_PROBLEM_PROGRAM;
    p->order = jet_order; p->n_coef = nc; p->cols = n_cols; p->n_file = n_file;
    p->words.assign(words, words + n_words);
    p->prog_consts.assign(consts, consts + n_consts);
    *out = p;
    return PDE_OK;
}

void pde_program_free(pde_program* p) { delete p; }

int pde_program_info(const pde_program* p, int* jet_order, int* n_coef, int* n_point_cols) {
    if (!p) { set_error("null program"); return PDE_E_INVALID; }
    if (jet_order) *jet_order = p->order;
    if (n_coef) *n_coef = p->n_coef;
    if (n_point_cols) *n_point_cols = p->cols;
    return PDE_OK;
}

int pde_program_point_table(const pde_program* p, const double* pts, int P, double* tab) {
    if (!p || !pts || !tab || P <= 0) { set_error("pde_program_point_table: bad argument"); return PDE_E_INVALID; }
    if (p->problem == PDE_PROBLEM_PROGRAM) {
        set_error("pde_program_point_table: the table of a run-time residual program is computed by its front end (residual_compiler.py)");
        return PDE_E_INVALID;
    }
    if (p->problem == PDE_PROBLEM_FORCE_FREE) {
        for (int i = 0; i < P; ++i) tab[i] = 1.0 / pts[i];     // w = 1/rho  (FFV:319: u_rho/rho)
    } else {
        // KV:69-91: Delta = r^2 - 2Mr + a^2, G = 1 - 2Mr/(r^2 + a^2 x^2)
        const double M = p->consts[0], a = p->consts[1];
        for (int i = 0; i < P; ++i) {
            const double r = pts[i], x = pts[P + i];
            const double Sg = r * r + a * a * x * x;
            const double G = 1.0 - 2.0 * M * r / Sg;
            const double G_r = -2.0 * M / Sg + 4.0 * M * r * r / (Sg * Sg);
            const double G_x = 4.0 * M * r * a * a * x / (Sg * Sg);
            const double Delta = r * r - 2.0 * M * r + a * a;
            const double om = 1.0 - x * x;
            tab[i] = G / om;
            tab[P + i] = G_r / om;
            tab[2 * (size_t)P + i] = G / Delta;
            tab[3 * (size_t)P + i] = G_x / Delta;
        }
    }
    return PDE_OK;
}

}  // extern "C"

// ------------------------------------------------------------- stage 2 (launch plumbing: launch.cuh)
// Kernel configuration: a CTA is W/4 groups of 4 warps (validate.cuh).  One CTA per SM:
//   W = 20 (640 threads, 94 registers, no local-memory spills): five warps per scheduler; fits up to
//          2 spill slots per lane (154 KB); W = 24 (80 registers, 128 B of register spills) measured equal;
//   W = 16 (512 threads, 104 registers): up to 3 spill slots;
//   W = 4  (four 128-thread CTAs per SM): small grids / deep spill stacks, and the DUMP (tooling) mode.
// Two points per lane in separate registers (NP = 2) was measured SLOWER on B200 again in v9 (194 registers,
// 8 warps/SM: 193 ms; 168 registers, 12 warps/SM: 167 ms; vs 138 ms for NP = 1 with 20 warps), DESIGN.md 4.1.
// PDE_B200_VARIANT=4|16 forces a smaller configuration.
#ifndef PDE_NP_BIG
#define PDE_NP_BIG 1
#endif
#ifndef PDE_NP_KERR
#define PDE_NP_KERR 2      // points per lane for the order-2 (6-coefficient) Kerr jets: 46.7 vs 37.9 G evals/s (depth-3 uniques)
#endif
#ifndef PDE_W_BIG
#define PDE_W_BIG 20
#endif
#ifndef PDE_W_KERR
#define PDE_W_KERR PDE_W_BIG   // Kerr, NP = 2, depth <= 3 uniques x 8: W = 20 (96 regs, 100 B spilled) 46.4 G evals/s; 16 (121 regs) 40.1; 24 (80 regs) 45.2
#endif
static int g_variant = -1;
template <int PROBLEM, bool DUMP, bool MAJ>
static int launch_validate(const ValidateParams& vp, cudaStream_t st) {
    if (g_variant < 0) { const char* e = getenv("PDE_B200_VARIANT"); g_variant = e ? atoi(e) : 0; }
    if constexpr (DUMP) {
        return launch_validate_cfg<PROBLEM, DUMP, 4, 1, 4, MAJ>(vp, st, nullptr);
    } else {
        bool fits = false;
        if (g_variant != 4 && g_variant != 16 && vp.P_eval >= 128) {
            constexpr int NPB = PROBLEM == PDE_PROBLEM_KERR ? PDE_NP_KERR : PDE_NP_BIG;
            constexpr int WB = PROBLEM == PDE_PROBLEM_KERR ? PDE_W_KERR : PDE_W_BIG;
            if (vp.P_eval % (128 * NPB) == 0 || vp.P_eval == vp.P) {
                int rc = launch_validate_cfg<PROBLEM, DUMP, WB, NPB, 1, MAJ>(vp, st, &fits);
                if (rc || fits) return rc;
            }
        }
        if (g_variant != 4 && PDE_W_BIG != 16 && vp.P_eval >= 128) {
            int rc = launch_validate_cfg<PROBLEM, DUMP, 16, 1, 1, MAJ>(vp, st, &fits);
            if (rc || fits) return rc;
        }
        return launch_validate_cfg<PROBLEM, DUMP, 4, 1, 4, MAJ>(vp, st, nullptr);
    }
}

template <int PROBLEM>
struct BuiltinLauncher {
    template <bool MAJ> static int launch(const ValidateParams& vp, cudaStream_t st) { return launch_validate<PROBLEM, false, MAJ>(vp, st); }
};

static int check_common(const pde_session* s, const pde_program* p, const void* code, const void* len,
                        int64_t n, int L, const void* pts, const void* tab, int P, int ns) {
    if (!have_device()) { set_error("no CUDA device: pde_engine_b200 has no CPU fallback"); return PDE_E_NODEVICE; }
    if (!s || !p || !pts || !tab || (n > 0 && (!code || !len))) { set_error("null argument"); return PDE_E_INVALID; }
    if (n < 0 || L < 4 || L > kMaxL || (L % 4) != 0) { set_error("L must be a multiple of 4 in [4, %d]", kMaxL); return PDE_E_INVALID; }
    if (P < 64 || (P % 64) != 0) { set_error("P must be a positive multiple of 64"); return PDE_E_INVALID; }
    if (ns < 1 || ns > 8) { set_error("spill_slots must be in 1..8"); return PDE_E_INVALID; }
    return PDE_OK;
}

static int check_tau_t0(double tau, double t0) {
    if (!(tau > 0.0 && tau <= 1.0)) { set_error("tau must be in (0, 1]"); return PDE_E_INVALID; }
    if (!(t0 >= 1e-6 && t0 <= 1.0)) { set_error("t0 (majorant radius) must be in [1e-6, 1]"); return PDE_E_INVALID; }
    return PDE_OK;
}

extern "C" {

static int validate_impl(const pde_session* s, const pde_program* p, const uint8_t* code, const uint32_t* row_off, const uint8_t* len,
                         int64_t n, int L, const double* pts, const double* tab, const double* prim, int n_prim, int P,
                         double tau, int min_finite, double vote_frac, double t0, int confirm_points, int n_ref, int spill_slots,
                         const pde_validate_out* out, void* stream);

int pde_validate(const pde_session* s, const pde_program* p, const uint8_t* code, const uint8_t* len,
                 int64_t n, int L, const double* pts, const double* tab, const double* prim, int n_prim, int P,
                 double tau, int min_finite, double vote_frac, double t0, int confirm_points, int n_ref, int spill_slots,
                 const pde_validate_out* out, void* stream) {
    return validate_impl(s, p, code, nullptr, len, n, L, pts, tab, prim, n_prim, P, tau, min_finite, vote_frac, t0, confirm_points,
                         n_ref, spill_slots, out, stream);
}

int pde_validate_csr(const pde_session* s, const pde_program* p, const uint8_t* pool, const uint32_t* row_off, const uint8_t* len,
                     int64_t n, int L, const double* pts, const double* tab, const double* prim, int n_prim, int P,
                     double tau, int min_finite, double vote_frac, double t0, int confirm_points, int n_ref, int spill_slots,
                     const pde_validate_out* out, void* stream) {
    if (n > 0 && !row_off) { set_error("pde_validate_csr: null row offsets"); return PDE_E_INVALID; }
    return validate_impl(s, p, pool, row_off, len, n, L, pts, tab, prim, n_prim, P, tau, min_finite, vote_frac, t0, confirm_points,
                         n_ref, spill_slots, out, stream);
}

static int validate_impl(const pde_session* s, const pde_program* p, const uint8_t* code, const uint32_t* row_off, const uint8_t* len,
                         int64_t n, int L, const double* pts, const double* tab, const double* prim, int n_prim, int P,
                         double tau, int min_finite, double vote_frac, double t0, int confirm_points, int n_ref, int spill_slots,
                         const pde_validate_out* out, void* stream) {
    int rc = check_common(s, p, code, len, n, L, pts, tab, P, spill_slots);
    if (rc) return rc;
    rc = check_tau_t0(tau, t0);
    if (rc) return rc;
    if (n == 0) return PDE_OK;              // an empty batch is a no-op (its buffers may be null)
    if (!out || !out->ratio_max || !out->resid_max || !out->scale_at || !out->n_finite || !out->n_votes || !out->survivor_bits) {
        set_error("pde_validate: null output"); return PDE_E_INVALID;
    }
    if (n_ref < 0 || n_ref > 4) { set_error("n_ref must be in 0..4"); return PDE_E_INVALID; }
    if (n >= 0x7fffffffLL) { set_error("pde_validate: n too large for one call"); return PDE_E_OVERFLOW; }
    if (confirm_points < 0 || confirm_points > P || (confirm_points % 128) != 0) {
        set_error("confirm_points must be 0 (one pass with majorants) or a multiple of 128 <= P"); return PDE_E_INVALID;
    }
    if (confirm_points > 0 && !out->scratch) { set_error("pde_validate: the two-pass mode needs out->scratch [n + 2] int32"); return PDE_E_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const bool is_program = p->problem == PDE_PROBLEM_PROGRAM;      // program.cu owns (and uploads) its own tables
    if (!is_program) {
        rc = upload_tables(s, tau, t0, st);
        if (rc) return rc;
    }
    PDE_CUDA(cudaMemsetAsync(out->survivor_bits, 0, sizeof(uint32_t) * (size_t)((n + 31) / 32), st));
    ValidateParams vp{};
    vp.code = code; vp.row_off = row_off; vp.len = len; vp.n = n; vp.L = L; vp.pts = pts; vp.tab = tab; vp.prim = prim; vp.n_prim = (prim && n_prim > 0) ? (n_prim < PDE_N_PRIM ? n_prim : PDE_N_PRIM) : 0; vp.P = P;
    vp.P_eval = P;
    vp.ns = spill_slots; vp.t0 = (float)t0; vp.tau = tau; vp.min_finite = min_finite; vp.vote_frac = vote_frac; vp.n_ref = out->ref_rs ? n_ref : 0;
    vp.ratio_max = out->ratio_max; vp.resid_max = out->resid_max; vp.scale_at = out->scale_at;
    vp.n_finite = out->n_finite; vp.n_votes = out->n_votes; vp.ref_rs = out->ref_rs; vp.survivor_bits = out->survivor_bits;
    if (is_program) return program_validate(s, p, vp, out, confirm_points, tau, t0, st);
    if (p->problem == PDE_PROBLEM_FORCE_FREE) return run_two_pass<BuiltinLauncher<PDE_PROBLEM_FORCE_FREE>>(vp, out, confirm_points, st);
    return run_two_pass<BuiltinLauncher<PDE_PROBLEM_KERR>>(vp, out, confirm_points, st);
}

int pde_eval_points(const pde_session* s, const pde_program* p, const uint8_t* code, const uint8_t* len,
                    int64_t n, int L, const double* pts, const double* tab, const double* prim, int n_prim, int P,
                    double tau, double t0, int spill_slots, double* jets, double* resid, double* scale, double* scale_maj,
                    float* maj, void* stream) {
    int rc = check_common(s, p, code, len, n, L, pts, tab, P, spill_slots);
    if (rc) return rc;
    rc = check_tau_t0(tau, t0);
    if (rc) return rc;
    if (n == 0) return PDE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (p->problem != PDE_PROBLEM_PROGRAM) {
        rc = upload_tables(s, tau, t0, st);
        if (rc) return rc;
    }
    ValidateParams vp{};
    vp.code = code; vp.len = len; vp.n = n; vp.L = L; vp.pts = pts; vp.tab = tab; vp.prim = prim; vp.n_prim = (prim && n_prim > 0) ? (n_prim < PDE_N_PRIM ? n_prim : PDE_N_PRIM) : 0; vp.P = P;
    vp.P_eval = P; vp.ns = spill_slots; vp.t0 = (float)t0; vp.tau = tau; vp.jets = jets; vp.resid = resid; vp.scale = scale; vp.scale_maj = scale_maj; vp.maj = maj;
    if (p->problem == PDE_PROBLEM_PROGRAM) return program_eval_points(s, p, vp, tau, t0, st);
    if (p->problem == PDE_PROBLEM_FORCE_FREE) return launch_validate<PDE_PROBLEM_FORCE_FREE, true, true>(vp, st);
    return launch_validate<PDE_PROBLEM_KERR, true, true>(vp, st);
}

int pde_fingerprint(const pde_session* s, const uint8_t* code, const uint8_t* len, int64_t n, int L,
                    const double* pts, const double* prim, int n_prim, int P, int spill_slots, int mantissa_bits,
                    double* values, uint64_t* key, int32_t* n_finite, void* stream) {
    static const pde_program dummy{};
    int rc = check_common(s, &dummy, code, len, n, L, pts, pts, P, spill_slots);
    if (rc) return rc;
    if (n == 0) return PDE_OK;
    if (!values || !key || !n_finite) { set_error("pde_fingerprint: null output"); return PDE_E_INVALID; }
    if (mantissa_bits < 8 || mantissa_bits > 51) { set_error("mantissa_bits must be in 8..51"); return PDE_E_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    rc = upload_tables(s, 1e-10, 0.0625, st);
    if (rc) return rc;
    // programs that are not evaluated (empty / malformed / too many spills) leave NaN rows -> key 0
    PDE_CUDA(cudaMemsetAsync(values, 0xff, sizeof(double) * (size_t)n * P, st));
    ValidateParams vp{};
    vp.code = code; vp.len = len; vp.n = n; vp.L = L; vp.pts = pts; vp.tab = pts; vp.prim = prim;
    vp.n_prim = (prim && n_prim > 0) ? (n_prim < PDE_N_PRIM ? n_prim : PDE_N_PRIM) : 0; vp.P = P;
    vp.P_eval = P; vp.ns = spill_slots; vp.t0 = 0.0625f; vp.tau = 1e-10; vp.resid = values;
    rc = launch_validate<kProblemValue, true, false>(vp, st);
    if (rc) return rc;
    int dev = 0, sms = 0;
    PDE_CUDA(cudaGetDevice(&dev));
    PDE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long want = (n + 7) / 8;
    const int grid = (int)(want < (long long)sms * 8 ? want : (long long)sms * 8);
    fingerprint_kernel<<<grid, 256, 0, st>>>(values, n, P, 52 - mantissa_bits,
                                             reinterpret_cast<unsigned long long*>(key), n_finite);
    count_launch();
    PDE_CUDA(cudaGetLastError());
    return PDE_OK;
}

static int fp64_peak_impl(bool three_operands, int iters, double* tflops, void* stream) {
    if (!have_device()) { set_error("no CUDA device"); return PDE_E_NODEVICE; }
    if (!tflops || iters < 1) { set_error("bad argument"); return PDE_E_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    PDE_CUDA(cudaGetDevice(&dev));
    PDE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8, threads = 256;
    double* buf = nullptr;
    PDE_CUDA(cudaMalloc(&buf, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    PDE_CUDA(cudaEventCreate(&e0));
    PDE_CUDA(cudaEventCreate(&e1));
    auto launch = [&](int n) {
        if (three_operands) fp64_peak3_kernel<<<blocks, threads, 0, st>>>(buf, n, 1.0000001, 1e-9);
        else fp64_peak_kernel<<<blocks, threads, 0, st>>>(buf, n, 1.0000001, 1e-9);
    };
    launch(iters / 4 + 1);  // warm-up
    PDE_CUDA(cudaEventRecord(e0, st));
    launch(iters);
    PDE_CUDA(cudaEventRecord(e1, st));
    count_launch(2);
    PDE_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    PDE_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 8.0 * 16.0 * (double)iters * (double)blocks * threads;
    *tflops = flops / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    return PDE_OK;
}

int pde_fp64_peak(int iters, double* tflops, void* stream) { return fp64_peak_impl(false, iters, tflops, stream); }
int pde_fp64_peak_3op(int iters, double* tflops, void* stream) { return fp64_peak_impl(true, iters, tflops, stream); }

}  // extern "C"
