/*
 * pde_b200.h -- C ABI of the B200-native hot path of pde-engine.
 *
 * The reference (PimDeWitte/pde-engine) is pure Python and has NO FFI for this
 * path: its boundary is duck-typed Python.  This header is the
 * C-ABI a maintainer would bind with ctypes (INTEGRATION.md shows the stub);
 * each entry point names the reference code it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, a negative PDE_E_* code on failure;
 *     pde_last_error() gives the message (thread local).  No exceptions cross
 *     the ABI.
 *   - "dev" pointers are device pointers on the current CUDA device, allocated
 *     by the caller (torch tensors); "host" pointers are plain host memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - opaque handles are owned by the library and freed with pde_*_free.
 *   - one host thread per device.
 */
#ifndef PDE_B200_H
#define PDE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDE_B200_ABI_VERSION 6

/* ---- error codes ------------------------------------------------------- */
#define PDE_OK            0
#define PDE_E_INVALID    -1   /* bad argument */
#define PDE_E_CUDA       -2   /* CUDA runtime error (message has the detail) */
#define PDE_E_NOMEM      -3
#define PDE_E_OVERFLOW   -4   /* output buffer too small / index space too large */
#define PDE_E_NODEVICE   -5   /* no CUDA device: there is NO CPU fallback */

/* ---- postfix bytecode (one byte per instruction) ------------------------ *
 * Semantics of every opcode: expression_operations.py:11-77 as seen through
 * sympify with the problem's locals (general_method_paper_reproduction.py:85-93,1257). */
enum {
    PDE_OP_END       = 0x00,
    PDE_OP_VAR0      = 0x01,  /* rho | r */
    PDE_OP_VAR1      = 0x02,  /* z   | x */
    PDE_OP_PRIM0     = 0x08,  /* PRIM(p), p < 8: per-problem primitive jet table */
    PDE_OP_ADD       = 0x10,
    PDE_OP_SUB       = 0x11,
    PDE_OP_MUL       = 0x12,
    PDE_OP_DIV       = 0x13,
    PDE_OP_NEG       = 0x18,  /* infix unary minus */
    PDE_OP_ABS       = 0x19,  /* Abs(x) */
    PDE_OP_SQRT      = 0x1A,  /* sqrt(x)  (EO:38-40) */
    PDE_OP_EXP       = 0x1B,  /* exp(x)   (EO:54-56) */
    PDE_OP_FN_NEG    = 0x20,  /* neg(x)          EO:30-32 */
    PDE_OP_FN_INV    = 0x21,  /* inv(x)          EO:34-36 */
    PDE_OP_FN_SQUARE = 0x22,  /* square(x)       EO:42-44 */
    PDE_OP_FN_POW32  = 0x23,  /* pow_3_2(x)      EO:46-48 */
    PDE_OP_FN_POWN32 = 0x24,  /* pow_neg_3_2(x)  EO:50-52 */
    PDE_OP_FN_EXPNEG = 0x25,  /* exp_neg(x)      EO:58-60 */
    PDE_OP_POW0      = 0x40,  /* POW(k), k < 64: x ** pow_table[k] */
    PDE_OP_CONST0    = 0x80   /* CONST(k), k < 128: const_table[k] */
};
#define PDE_N_PRIM   8
#define PDE_N_POW    64
#define PDE_N_CONST  128
/* reserved slots: CONST(0) = 1; POW(0) = 3/2, POW(1) = -3/2, POW(2) = 2 */

/* per-expression compile flags (0 = device-evaluable) */
#define PDE_FLAG_UNSUPPORTED 1  /* I, zoo, x**y, unknown name ... -> CPU only */
#define PDE_FLAG_TABLE_FULL  2
#define PDE_FLAG_TOO_LONG    4

/* per-expression string predicates used by the prune rules (LBF:134-136,143-152) */
#define PDE_ATTR_HAS_VARS    1  /* 'r' in s or 'x' in s or 'rho' in s or 'z' in s */
#define PDE_ATTR_IS_ONE      2  /* s == '1' */
#define PDE_ATTR_STARTS_INV  4  /* s.startswith('inv(') */

/* enumerator op codes in (op, i, j) triples: the iteration order of the
 * reference's dicts (expression_operations.py:80-106) */
#define PDE_ENUM_N_UNARY  8   /* neg inv sqrt square pow_3_2 pow_neg_3_2 exp exp_neg */
#define PDE_ENUM_N_BINARY 5   /* add sub mul div geom_sum (the 4 special ops emit nothing, LBF:170-195) */

/* problems with a built-in residual operator */
#define PDE_PROBLEM_FORCE_FREE 0  /* problems/force_free/validator.py:305-347, order-4 jets */
#define PDE_PROBLEM_KERR       1  /* problems/kerr_magnetosphere/validator.py:77-91, order-2 jets */

typedef struct pde_session pde_session;
typedef struct pde_exprset pde_exprset;
typedef struct pde_program pde_program;

int         pde_abi_version(void);
const char *pde_last_error(void);
/* number of visible CUDA devices (0 => every compute entry point returns PDE_E_NODEVICE) */
int         pde_device_count(void);

/* ------------------------------------------------------------------------
 * Session = the symbol table sympify works with (GM:85-93): two coordinate
 * names, named constants with the numeric values used by the validator
 * (KV: M_value, a_value), and the append-only constant / exponent tables the
 * bytecode indexes.
 * ---------------------------------------------------------------------- */
int  pde_session_create(const char *var0, const char *var1,
                        const char *const *const_names, const double *const_vals, int n_named,
                        pde_session **out);
void pde_session_free(pde_session *s);
/* copy the current tables out (host): const_vals[PDE_N_CONST], pow_vals[PDE_N_POW] */
int  pde_session_tables(const pde_session *s, double *const_vals, int *n_const,
                        double *pow_vals, int *n_pow);
/* key strings ("1", "1/3", "M", ...) of slot k, for decompilation / debugging */
const char *pde_session_const_key(const pde_session *s, int k);
const char *pde_session_pow_key(const pde_session *s, int k);

/* ------------------------------------------------------------------------
 * Host compiler: expression strings -> term-structured postfix bytecode.
 * Replaces sympify(expr_str, locals) on the hot path (GM:1257) and prepares
 * the operands of the textual splice (LBF:170-195).  Also evaluates the
 * string predicates of the prune rules and the lexicographic rank used for
 * the `a > b` swap (LBF:168-169).  The set is uploaded to the current device.
 * ---------------------------------------------------------------------- */
int  pde_compile_exprs(pde_session *s, const char *const *strs, int n, pde_exprset **out);
/* same, for one NUL-separated blob: string i = blob + offsets[i], offsets[n] = blob size (each
 * string NUL terminated).  The parse runs on up to 16 host threads (PDE_B200_COMPILE_THREADS);
 * table slots are assigned afterwards in string order, so the bytecode does not depend on the
 * thread count.  offsets may be NULL: the blob is then n NUL-terminated strings back to back and
 * the library finds the terminators itself. */
int  pde_compile_exprs_packed(pde_session *s, const char *blob, const uint32_t *offsets, int n, pde_exprset **out);
/* same, for a blob of KNOWN size that must hold exactly n NUL-terminated strings back to back (what
 * travels between the ranks of a sharded batch): the library finds the terminators inside blob_bytes
 * and fails with PDE_E_INVALID if there are fewer or more than n strings -- the caller need not scan
 * the blob itself. */
int  pde_compile_exprs_blob(pde_session *s, const char *blob, size_t blob_bytes, int n, pde_exprset **out);
/* Frees the handle.  Its device mirrors (made by the first pde_enumerate* call, from the library's stream-ordered
 * pool) go back to the pool after the last enumerate kernel that read them has finished: no device-wide
 * synchronisation, safe to call while that work is still queued. */
void pde_exprset_free(pde_exprset *e);
int  pde_exprset_size(const pde_exprset *e, int *n_expr, int *n_terms, int *n_pool_bytes);
/* host copies of the compiled form:
 *   flags[n], attrs[n], rank[n] (dense rank of the string among the set),
 *   term_begin[n+1], term_sign[n_terms] (+1/-1), term_off[n_terms+1], pool[n_pool_bytes] */
int  pde_exprset_export(const pde_exprset *e, uint8_t *flags, uint8_t *attrs, uint32_t *rank,
                        uint32_t *term_begin, int8_t *term_sign, uint32_t *term_off, uint8_t *pool);
/* "whole" programs  t1 [NEG] (tk ADD|SUB)*  into a host [n, L] array (zero padded);
 * len[i] = 0 and flags != 0 for expressions that are not device-evaluable or longer than L */
int  pde_exprset_programs(const pde_exprset *e, int L, uint8_t *code_host, uint8_t *len_host);

/* ------------------------------------------------------------------------
 * Stage 1: the combinatorial generator.
 * Replaces the candidate loops of FastExpressionGenerator.stream_generate
 * (LBF:139-195).  `e` holds E[1] ++ E[2] ++ ... ++ E[depth-1]; depth_begin has
 * `depth` entries + 1 (depth_begin[k-1] = index of the first expression of
 * depth k, depth_begin[depth-1] = total).
 *
 *   pde_enumerate_count : number of candidates the reference would append
 *   pde_enumerate       : for each candidate, in the reference's order:
 *       triple[c] = (op, i, j)  op 0..7 unary / 8..12 binary, i,j indices into e
 *                               (after the add/mul swap; j = -1 for unary)
 *       code[c, L], len[c]      spliced postfix program (len 0 = operand not
 *                               device-compilable or program longer than L)
 *       hash[c]                 64-bit structural hash of (len, code)
 *   pde_dedup           : first_occurrence[c] = 1 iff no earlier candidate has
 *                         the same program (hash match confirmed byte-wise);
 *                         candidates with len 0 are always kept.
 * ---------------------------------------------------------------------- */
int  pde_enumerate_count(const pde_exprset *e, const int32_t *depth_begin, int depth, int prune,
                         int64_t *n_candidates, void *stream);
int  pde_enumerate(const pde_exprset *e, const int32_t *depth_begin, int depth, int prune,
                   int64_t first, int64_t count, int L,
                   int32_t *triple_dev, uint8_t *code_dev, uint8_t *len_dev, uint64_t *hash_dev,
                   void *stream);
int  pde_dedup(const uint8_t *code_dev, const uint8_t *len_dev, const uint64_t *hash_dev,
               int64_t n, int L, uint8_t *first_occurrence_dev, int64_t *n_unique, void *stream);

/* CSR form of the same output: the programs back to back in ONE byte pool, each padded to 16 bytes (an empty
 * program takes none), instead of [count][L] rows -- the mean program of the real candidate sets is 15 bytes, so
 * a depth-5 pass writes ~49 B per candidate instead of 149 (L = 128).  offset[c] (16-byte units, relative to the
 * start of the pool) is the start of candidate first + c, offset[count] the end.  L only bounds a program's length
 * (longer ones are emitted empty, as above).  pde_enumerate_csr_size returns the pool bytes the window needs (the
 * rows of the blocks it touches); triple / len / hash as in pde_enumerate.  pde_dedup_csr / pde_validate_csr read
 * the same layout. */
int  pde_enumerate_csr_size(const pde_exprset *e, const int32_t *depth_begin, int depth, int prune,
                            int64_t first, int64_t count, int L, int64_t *pool_bytes, void *stream);
int  pde_enumerate_csr(const pde_exprset *e, const int32_t *depth_begin, int depth, int prune,
                       int64_t first, int64_t count, int L,
                       int32_t *triple_dev, uint32_t *offset_dev /*[count + 1]*/, uint8_t *pool_dev,
                       uint8_t *len_dev, uint64_t *hash_dev, void *stream);
int  pde_dedup_csr(const uint8_t *pool_dev, const uint32_t *offset_dev, const uint8_t *len_dev, const uint64_t *hash_dev,
                   int64_t n, uint8_t *first_occurrence_dev, int64_t *n_unique, void *stream);

/* synthetic depth-d trees of SURVEY 8d (tree semantics, splitmix64 seeded per tree) */
int  pde_synth_trees(uint64_t seed, int64_t first, int64_t count, int depth, int L,
                     uint8_t *code_dev, uint8_t *len_dev, uint64_t *hash_dev, void *stream);

/* ------------------------------------------------------------------------
 * Residual programs (one per problem, compiled once).
 * Replaces the symbolic construction of det_M (FFV:305-347) / lhs (KV:77-91).
 * `consts`: force-free none; Kerr (M, a) (KV:36-37).
 * ---------------------------------------------------------------------- */
int  pde_compile_residual(int problem_id, const double *consts, int n_consts, pde_program **out);
void pde_program_free(pde_program *p);
int  pde_program_info(const pde_program *p, int *jet_order, int *n_coef, int *n_point_cols);
/* per-point coefficient table (host -> host): force-free [P,1] = 1/rho;
 * Kerr [P,4] = G/(1-x^2), d_r(G/(1-x^2)), G/Delta, d_x(G/Delta)  (column major [cols][P]) */
int  pde_program_point_table(const pde_program *p, const double *pts_host /*[2][P]*/, int P,
                             double *table_host /*[cols][P]*/);

/* ------------------------------------------------------------------------
 * Run-time residual programs (BASELINE north_star item 3: "each problem's PDE
 * residual operator is compiled once into device bytecode").  The reference's
 * plugin seam is ProblemSpec.validator (problems/__init__.py:34-63); a plugin
 * states its PDE as a SymPy formula over a generic u (KerrMagnetosphereValidator._lhs,
 * problems/kerr_magnetosphere/validator.py:77-91; det M, problems/force_free/validator.py:305-347).
 * pde_engine_b200/residual_compiler.py turns such a formula -- a polynomial in
 * the partial derivatives of u whose coefficients depend only on the point --
 * into a straight-line scalar program; this entry takes the program, and
 * pde_validate / pde_eval_points interpret it after the candidate's jet.  A new
 * plugin therefore needs no CUDA and no rebuild (PDE_PROBLEM_FORCE_FREE / _KERR
 * remain as build-time specialisations of the same two residuals).
 *
 * Machine (csrc/validate.cuh): a file F of float64 per lane + one accumulator.
 *   F[0 .. n_coef)              d_g = the partial derivatives of u, g = jidx(i, j) = (i+j)(i+j+1)/2 + j
 *                               for d^(i+j) u / d x0^i d x1^j  (n_coef = 6 for order 2, 15 for order 4)
 *   F[n_coef .. + n_cols)       the point's row of the coefficient table (caller-computed, [n_cols][P])
 *   F[.. + n_consts)            consts
 *   F[.. n_file)                temporaries
 * Word = op | a << 4 | b << 12 | dst << 20 | neg << 28   (neg: the a operand is negated)
 *   MUL   F[dst] = F[a] * F[b]          ACC0  acc = F[a] * F[b]        ACC   acc = fma(F[a], F[b], acc)
 *   LDA   acc = F[a]                    ADDA  acc = acc + F[a]         STA   F[dst] = acc
 *   OUT   R = acc (last word)
 * The same program run on magnitudes (|d_g| + theta_|g|, |column|, |const|, no
 * signs) gives the decision scale S~ (see pde_validate); with theta = 0 the
 * plain scale S = sum of |monomial|.  Compile-time checks: every operand is an
 * input or a temporary written earlier, dst is a temporary, the program ends
 * with OUT, n_file <= PDE_R_MAX_FILE.  jet_order is 2 or 4.
 * ---------------------------------------------------------------------- */
enum pde_residual_op {
    PDE_R_END = 0, PDE_R_MUL = 1, PDE_R_ACC0 = 2, PDE_R_ACC = 3, PDE_R_LDA = 4, PDE_R_STA = 5, PDE_R_ADDA = 6, PDE_R_OUT = 7
};
#define PDE_R_WORD(op, a, b, dst, neg) \
    ((uint32_t)(op) | ((uint32_t)(a) << 4) | ((uint32_t)(b) << 12) | ((uint32_t)(dst) << 20) | ((uint32_t)((neg) ? 1 : 0) << 28))
#define PDE_R_MAX_WORDS  2048
#define PDE_R_MAX_CONSTS 16
#define PDE_R_MAX_COLS   8
#define PDE_R_MAX_FILE   51      /* order 4: three spill slots of 17 elements at 16 warps per CTA */
#define PDE_PROBLEM_PROGRAM 3    /* pde_program_info reports it for programs made by the entry below */
int  pde_compile_residual_program(int jet_order, int n_cols, const double *consts, int n_consts,
                                  const uint32_t *words, int n_words, pde_program **out);

/* ------------------------------------------------------------------------
 * Stage 2: the batched validator.
 * Replaces the per-candidate validate() call of emit_to_db (GM:1302-1316) and
 * the validator worker pool (GM:1672-1824) as a *filter*: a candidate is
 * rejected only when its residual is numerically non-zero: n_finite >= min_finite
 * and |R| > tau * S~ at >= vote_frac * n_finite of the finite points;
 * everything else survives to the CPU validator.
 *
 * S~ is the DECISION SCALE: the residual's majorant (sum of |monomial|, every
 * partial of order n replaced by the sum of the |partials| of that order)
 * evaluated at partials inflated by theta_n = 2 eps W n! / (t0^n tau), where W is
 * the round-off majorant the interpreter carries next to every jet
 * (|computed c_g - exact c_g| <= eps W / t0^|g|; rules and proof sketch in
 * oracle/majorant.py).  For an exact solution |R| <= tau S~ under
 * ANY float64 evaluation order, so a point can only vote when its residual is
 * non-zero beyond the propagated rounding error of u's own jet; t0 (0 < t0 <= 1,
 * the Python layer uses 1/16) is the radius the majorant series are evaluated at: points
 * closer than t0 to a pole of a sub-expression do not vote.
 *
 * Two passes (confirm_points > 0, a multiple of 128; default 128): carrying the
 * majorants costs about a fifth of the kernel's throughput, and a rejection is
 * sound as soon as the majorant rule votes it on ANY sufficiently large set of
 * points.  Pass 1 therefore sweeps all P points WITHOUT majorants (theta = 0)
 * and only PROPOSES rejections; pass 2 re-evaluates the proposed rejections on
 * the first confirm_points points of the same grid WITH the majorants and a
 * rejection stands only if that pass votes it too (n_finite >= min_finite and
 * votes >= vote_frac * n_finite there); every other candidate gets its survivor
 * bit back.  ratio_max / resid_max / scale_at / n_finite / n_votes report pass
 * 1 (whole grid, theta = 0), `confirm` reports pass 2, ref_rs holds (R, S~) of
 * pass 2 for re-examined candidates.  confirm_points = 0: one pass over all P
 * points with the majorants carried (what the parity tests compare against).
 *
 *   code[n, L], len[n]   postfix programs (len 0 = skip: survivor, n_finite 0)
 *   pts[2][P]            collocation grid, SoA, P a multiple of 64
 *   table[cols][P]       pde_program_point_table
 *   prim[n_prim][P/32][16][32]  jets of PRIM(p) leaves in 32-point stripe blocks: coefficient g of
 *                            point q at [p][q / 32][g][q % 32]; row 15 holds the leaf's majorant pair
 *                            (D, W) as two float32 (rows n_coef..14 are padding), so a
 *                            warp reads a leaf with coalesced loads at immediate offsets; may be
 *                            NULL (n_prim = 0) if no program uses PRIM; a program that uses PRIM(p) with
 *                            p >= n_prim is reported as malformed (n_finite = -2), never dereferenced
 * outputs (per candidate):
 *   ratio_max   max |R|/S~ over finite points       resid_max  max |R|
 *   scale_at    S~ at the arg-max of the ratio      n_finite, n_votes
 *   ref_rs[n, n_ref, 2]  (R, S~) at the first n_ref points (the reference's own
 *                        test points, FFV:296-297 / KV:168-172); n_ref <= 4
 *   survivor_bits[(n+31)/32]  bit c = 1 iff candidate c is NOT rejected
 *   n_finite < 0: not evaluated (-1 empty program, -2 malformed, -3 needs more
 *   than `spill_slots` (1..8) spilled jets) -> survivor
 *
 * Every output is addressed by candidate and does not depend on which warp group
 * evaluates a candidate or when: the library deals chunks of candidates to the
 * groups dynamically and walks batches of 2^18 candidates or more in the order
 * of their op signature (instruction-cache locality; work space from the
 * library's stream-ordered pool, 24 B per candidate).  Environment switches
 * for A/B measurements, read once per process: PDE_B200_NO_ORDER (keep the
 * caller's order), PDE_B200_STATIC_DEAL (round-robin chunks); PDE_B200_PROFILE
 * prints host-side phase timings to stderr.
 * ---------------------------------------------------------------------- */
typedef struct pde_validate_out {
    double   *ratio_max;     /* [n] */
    double   *resid_max;     /* [n] */
    double   *scale_at;      /* [n] */
    int32_t  *n_finite;      /* [n] */
    int32_t  *n_votes;       /* [n] */
    double   *ref_rs;        /* [n, n_ref, 2] or NULL */
    uint32_t *survivor_bits; /* [(n+31)/32] */
    int32_t  *confirm;       /* [n, 2] or NULL: (n_finite, n_votes) of the confirmation pass; -1 = not re-examined */
    int32_t  *scratch;       /* [n + 2]: work space of the two-pass mode (may be NULL when confirm_points = 0) */
} pde_validate_out;

int  pde_validate(const pde_session *s, const pde_program *p,
                  const uint8_t *code_dev, const uint8_t *len_dev, int64_t n, int L,
                  const double *pts_dev, const double *table_dev, const double *prim_dev, int n_prim, int P,
                  double tau, int min_finite, double vote_frac, double t0, int confirm_points, int n_ref, int spill_slots,
                  const pde_validate_out *out, void *stream);

/* the same filter on CSR rows (pde_enumerate_csr): program c = pool + 16 * row_off[c], len[c] bytes; L = the
 * staging size per program (>= the longest program, a multiple of 4) */
int  pde_validate_csr(const pde_session *s, const pde_program *p,
                      const uint8_t *pool_dev, const uint32_t *row_off_dev, const uint8_t *len_dev, int64_t n, int L,
                      const double *pts_dev, const double *table_dev, const double *prim_dev, int n_prim, int P,
                      double tau, int min_finite, double vote_frac, double t0, int confirm_points, int n_ref, int spill_slots,
                      const pde_validate_out *out, void *stream);

/* parity / tooling entry: full per-point output for SMALL batches (each may be NULL).
 *   jets[n, n_coef, P]   normalised Taylor coefficients of u
 *   resid[n, P]          R
 *   scale[n, P]          S  = sum of |monomial| of R (the scale per-point parity is quoted against)
 *   scale_maj[n, P]      S~ = the decision scale of pde_validate
 *   maj[n, 3, P] float32 (V, D, W): the majorants of the finished jet */
int  pde_eval_points(const pde_session *s, const pde_program *p,
                     const uint8_t *code_dev, const uint8_t *len_dev, int64_t n, int L,
                     const double *pts_dev, const double *table_dev, const double *prim_dev, int n_prim, int P,
                     double tau, double t0, int spill_slots,
                     double *jets_dev, double *resid_dev, double *scale_dev, double *scale_maj_dev, float *maj_dev,
                     void *stream);

/* ------------------------------------------------------------------------
 * Function fingerprints: a numeric pre-bucketing for the
 * DB-normalisation step of emit_to_db (GM:1256-1286), which runs SymPy
 * simplify(expand(.)) on every row only to have UNIQUE(normalized) (GM:1407)
 * drop rows that denote a function already stored.
 *   values[n, P]   u(x_k) at the P points (device; written by the call, kept
 *                  as evidence / for an exact re-check by the caller)
 *   key[n]         64-bit key of the values rounded to `mantissa_bits` (8..51)
 *                  bits: equal functions -> equal keys up to round-off at a
 *                  rounding boundary (a split bucket costs one redundant CPU
 *                  simplify; different functions do not merge unless they agree
 *                  to 2^-mantissa_bits at all P points); 0 = no finite value
 *                  (not evaluated / non-finite everywhere): leave to the CPU
 *   n_finite[n]    points with a finite value (0 is the device analogue of
 *                  _has_degenerate_denominator, GM:134-199: a suspect, never a verdict)
 * ---------------------------------------------------------------------- */
int  pde_fingerprint(const pde_session *s,
                     const uint8_t *code_dev, const uint8_t *len_dev, int64_t n, int L,
                     const double *pts_dev, const double *prim_dev, int n_prim, int P,
                     int spill_slots, int mantissa_bits,
                     double *values_dev, uint64_t *key_dev, int32_t *n_finite_dev, void *stream);

/* number of kernels launched by this library since load (bench "gpu_launches") */
int64_t pde_launch_count(void);

/* FP64 pipe microbenchmark: register-resident DFMA chains; returns TFLOP/s */
int  pde_fp64_peak(int iters, double *tflops, void *stream);
/* the same for DFMAs that read three different register pairs, acc = fma(y, z, acc) -- the operand pattern of
 * a jet convolution: 3 cycles per scheduler instead of 2 (the register file delivers two new 64-bit operands
 * per DFMA slot), i.e. 2/3 of the rate above */
int  pde_fp64_peak_3op(int iters, double *tflops, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PDE_B200_H */
