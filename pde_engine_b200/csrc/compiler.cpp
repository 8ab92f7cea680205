// compiler.cpp -- host compiler: expression strings -> term-structured postfix bytecode.
//
// Replaces sympify(expr_str, locals) on the hot path
// (general_method_paper_reproduction.py:1257) and prepares the operands of the
// reference's textual splice (lean_normalizer/lean_bridge_fixed.py:170-195).
// The grammar is the subset of Python's expression grammar that SymPy's str()
// printer emits:
//     expr   := term (('+'|'-') term)*
//     term   := factor (('*'|'/') factor)*
//     factor := ('+'|'-') factor | power
//     power  := atom ('**' factor)?
//     atom   := INT | NAME | NAME '(' expr ')' | '(' expr ')'
// Pipeline: parse -> IR with exact-rational constant folding -> split the top
// level into additive terms along the left spine of the +/- chain (leading
// unary minus of a term becomes its sign) -> postfix per term.
// tests/test_compiler.py checks the output byte-for-byte against
// oracle/parser.py (which uses Python's own `ast`).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <memory>
#include <numeric>
#include <string>
#include <chrono>
#include <thread>
#include <vector>

#include "common.h"

namespace {

using i64 = long long;
constexpr i64 LIMIT = 1LL << 53;

struct Unsupported {};
struct TableFull {};

struct Rat {
    i64 n = 0, d = 1;
};

static i64 gcd64(i64 a, i64 b) {
    if (a < 0) a = -a;
    if (b < 0) b = -b;
    while (b) { i64 t = a % b; a = b; b = t; }
    return a ? a : 1;
}

static Rat make_rat(__int128 n, __int128 d) {
    if (d == 0) throw Unsupported();
    if (d < 0) { n = -n; d = -d; }
    if (d == 1 && n <= LIMIT && n >= -LIMIT) { Rat r; r.n = (i64)n; r.d = 1; return r; }      // integers: the common case
    if (n < ((__int128)1 << 62) && n > -((__int128)1 << 62) && d < ((__int128)1 << 62)) {      // 64-bit gcd (128-bit % is a library call)
        i64 a = (i64)(n < 0 ? -n : n), b = (i64)d;
        while (b) { const i64 t = a % b; a = b; b = t; }
        if (a == 0) a = 1;
        const i64 nn = (i64)n / a, dd = (i64)d / a;
        if (nn > LIMIT || nn < -LIMIT || dd > LIMIT) throw Unsupported();
        Rat r; r.n = nn; r.d = dd;
        return r;
    }
    // reduce with 128-bit gcd
    __int128 a = n < 0 ? -n : n, b = d;
    while (b) { __int128 t = a % b; a = b; b = t; }
    if (a == 0) a = 1;
    n /= a; d /= a;
    if (n > LIMIT || n < -LIMIT || d > LIMIT) throw Unsupported();
    Rat r; r.n = (i64)n; r.d = (i64)d;
    return r;
}

// A CONST / POW byte whose table slot is assigned later, in expression order (the tables are
// append-only per session and their numbering is part of the bytecode, so the parallel parse
// must not decide it).
struct Fix {
    uint32_t pos;        // byte position in the worker's pool
    uint8_t is_pow;
    uint8_t named;       // named constant (M, a, E): num = index into session->named
    i64 num, den;
};

static std::string rat_key(const Rat& r) {
    return r.d == 1 ? std::to_string(r.n) : std::to_string(r.n) + "/" + std::to_string(r.d);
}

// ONE-PASS parser + emitter (round 2; the first version built an IR tree per string and walked it twice: 0.9 us per
// string and thread, the bound of the depth-4 wall time at every GPU count).  The recursive descent appends postfix
// bytes to the worker's pool as it goes; what the tree version expressed with nodes lives in the value it returns:
//   sign    a leading unary minus of the sub-expression, NOT yet emitted -- it travels up the left spine of * / chains
//           (-a*b*c -> a b MUL c MUL NEG) and is written as a trailing NEG only where the sub-expression is used as an
//           operand; in an additive chain it merges into the ADD / SUB that joins the term
//   konst   the sub-expression is one exact rational (one placeholder byte at the tail of the pool + the last Fix):
//           constant (op) constant folds in place, -constant negates in place, base ** constant takes the exponent
//   start   where its bytes begin (they always end at the current write position)
// The output is byte-for-byte what the tree version produced (tests/test_compiler.py: 200 k strings against
// oracle/parser.py, which parses with Python's own `ast`).
struct Val {
    int sign = 1;
    bool konst = false;
    Rat rat;
    uint32_t start = 0;
};

struct TermRec { int sign; uint32_t end; };

struct OnePass {
    const char* s;
    size_t pos = 0, n;
    const pde_session* sess;
    uint8_t* base;
    uint8_t*& wp;
    std::vector<Fix>& fixes;
    int depth = 0;

    OnePass(const char* str, size_t len, const pde_session* se, uint8_t* b, uint8_t*& w, std::vector<Fix>& f)
        : s(str), n(len), sess(se), base(b), wp(w), fixes(f) {}

    void ws() { while (pos < n && (s[pos] == ' ' || s[pos] == '\t')) ++pos; }
    bool peek(char c) { ws(); return pos < n && s[pos] == c; }
    bool peek2(const char* t) { ws(); return pos + 1 < n && s[pos] == t[0] && s[pos + 1] == t[1]; }
    void put(uint8_t b) { *wp++ = b; }
    uint32_t here() const { return (uint32_t)(wp - base); }

    Val push_const(const Rat& r) {
        Val v; v.konst = true; v.rat = r; v.start = here();
        fixes.push_back({here(), 0, 0, r.n, r.d});
        put(PDE_OP_CONST0);
        return v;
    }
    void pop_const() { --wp; fixes.pop_back(); }                     // the constant at the tail (byte + Fix)
    void use(const Val& v) { if (v.sign < 0) put(PDE_OP_NEG); }      // v becomes an operand: its pending minus is written now

    static Rat fold(char op, const Rat& l, const Rat& r) {
        const __int128 an = l.n, ad = l.d, bn = r.n, bd = r.d;
        switch (op) {
            case '+': return make_rat(an * bd + bn * ad, ad * bd);
            case '-': return make_rat(an * bd - bn * ad, ad * bd);
            case '*': return make_rat(an * bn, ad * bd);
            default:
                if (bn == 0) throw Unsupported();
                return make_rat(an * bd, ad * bn);
        }
    }

    // expr := term (('+'|'-') term)*.  top != null: the top level of the string -- the additive terms are RECORDED (sign,
    // end of body) instead of joined, because the enumerator splices operands term-wise (LBF:170-195).
    // The terms are those of the left spine of the PARSED tree, and parentheses leave no trace in it: `(a + b) + c` has
    // the three terms a, b, c.  So when the first term of a top-level chain is a complete parenthesised expression
    // (`leading_paren_is_a_term`), its inner chain is continued instead of being closed (`parse_first`).
    bool leading_paren_is_a_term(size_t& open_at) {
        size_t q = pos;
        while (q < n && (s[q] == ' ' || s[q] == '\t' || s[q] == '+')) ++q;        // unary plus is transparent
        if (!(q < n && s[q] == '(')) return false;
        open_at = q;
        int level = 0;
        for (; q < n; ++q) {
            if (s[q] == '(') ++level;
            else if (s[q] == ')' && --level == 0) break;
        }
        if (q >= n) return false;
        ++q;
        while (q < n && (s[q] == ' ' || s[q] == '\t')) ++q;
        return q >= n || s[q] == '+' || s[q] == '-' || s[q] == ')';
    }

    Val parse_first(std::vector<TermRec>* top, bool& chain) {
        size_t open_at = 0;
        if (top && leading_paren_is_a_term(open_at)) {
            if (++depth > 200) throw Unsupported();
            pos = open_at + 1;
            Val acc = parse_first(top, chain);
            additive_loop(acc, chain, top);
            if (!peek(')')) throw Unsupported();
            ++pos;
            --depth;
            return acc;
        }
        return parse_term();
    }

    void additive_loop(Val& acc, bool& chain, std::vector<TermRec>* top) {
        for (;;) {
            ws();
            if (!(pos < n && (s[pos] == '+' || s[pos] == '-'))) break;
            const char op = s[pos++];
            if (!chain && !acc.konst) {               // a non-constant first term: the chain starts here
                if (top) top->push_back({acc.sign, here()});
                else use(acc);
                acc.sign = 1;
                chain = true;
            }
            const Val t = parse_term();
            if (!chain) {                             // acc is a constant so far
                if (t.konst) {                        // constant (+|-) constant: folded in place
                    const Rat r = fold(op, acc.rat, t.rat);
                    pop_const(); pop_const();
                    const uint32_t st = acc.start;
                    acc = push_const(r);
                    acc.start = st;
                    continue;
                }
                if (top) top->push_back({1, t.start});   // the constant is term 1 (it ends where t begins)
                chain = true;
                acc.konst = false;
            }
            const int sg = (op == '+' ? 1 : -1) * t.sign;
            if (top) top->push_back({sg, here()});
            else put(sg > 0 ? PDE_OP_ADD : PDE_OP_SUB);
        }
    }

    Val parse_expr(std::vector<TermRec>* top) {
        if (++depth > 200) throw Unsupported();
        bool chain = false;
        Val acc = parse_first(top, chain);
        additive_loop(acc, chain, top);
        if (top && !chain) top->push_back({acc.sign, here()});
        --depth;
        return acc;
    }

    // term := factor (('*'|'/') factor)*
    Val parse_term() {
        Val l = parse_factor();
        for (;;) {
            ws();
            if (!(pos < n && (s[pos] == '*' || s[pos] == '/') && !(pos + 1 < n && s[pos] == '*' && s[pos + 1] == '*'))) break;
            if (s[pos] == '/' && pos + 1 < n && s[pos + 1] == '/') throw Unsupported();
            const char op = s[pos++];
            const Val r = parse_factor();
            if (l.konst && r.konst) {
                const Rat q = fold(op, l.rat, r.rat);
                pop_const(); pop_const();
                const uint32_t st = l.start;
                l = push_const(q);
                l.start = st;
                continue;
            }
            use(r);
            put(op == '*' ? PDE_OP_MUL : PDE_OP_DIV);
            l.konst = false;                          // the sign of the left spine stays pending
        }
        return l;
    }

    // factor := ('+'|'-') factor | power
    Val parse_factor() {
        ws();
        if (++depth > 200) throw Unsupported();
        Val v;
        if (pos < n && s[pos] == '+') { ++pos; v = parse_factor(); }
        else if (pos < n && s[pos] == '-') {
            ++pos;
            v = parse_factor();
            if (v.konst) { v.rat.n = -v.rat.n; fixes.back().num = v.rat.n; }
            else v.sign = -v.sign;
        } else v = parse_power();
        --depth;
        return v;
    }

    // power := atom ('**' factor)?
    Val parse_power() {
        Val b = parse_atom();
        if (!peek2("**")) return b;
        pos += 2;
        const Val e = parse_factor();
        if (!e.konst) throw Unsupported();
        const Rat k = e.rat;
        pop_const();                                  // the exponent is not an operand: it becomes part of the POW byte
        if (b.konst && k.d == 1) {
            if (k.n > 64 || k.n < -64) throw Unsupported();
            if (b.rat.n == 0 && k.n < 0) throw Unsupported();
            __int128 nn = 1, dd = 1;
            const i64 e2 = k.n < 0 ? -k.n : k.n;
            for (i64 i = 0; i < e2; ++i) {
                nn *= b.rat.n; dd *= b.rat.d;
                if (nn > ((__int128)1 << 100) || nn < -((__int128)1 << 100) || dd > ((__int128)1 << 100)) throw Unsupported();
            }
            const Rat r = k.n < 0 ? make_rat(dd, nn) : make_rat(nn, dd);
            pop_const();
            const uint32_t st = b.start;
            Val v = push_const(r);
            v.start = st;
            return v;
        }
        use(b);
        fixes.push_back({here(), 1, 0, k.n, k.d});
        put(PDE_OP_POW0);
        Val v; v.start = b.start;
        return v;
    }

    // atom := INT | NAME | NAME '(' expr ')' | '(' expr ')'
    Val parse_atom() {
        ws();
        if (pos >= n) throw Unsupported();
        const char c = s[pos];
        if (c == '(') {
            ++pos;
            const Val e = parse_expr(nullptr);
            if (!peek(')')) throw Unsupported();
            ++pos;
            return e;
        }
        if (c >= '0' && c <= '9') {
            // up to 18 digits in 64-bit arithmetic (every literal of the real candidate sets), beyond that 128-bit
            unsigned long long v64 = 0;
            int nd = 0;
            while (pos < n && s[pos] >= '0' && s[pos] <= '9' && nd < 18) { v64 = v64 * 10 + (unsigned)(s[pos] - '0'); ++pos; ++nd; }
            __int128 v = (__int128)v64;
            while (pos < n && s[pos] >= '0' && s[pos] <= '9') {
                v = v * 10 + (s[pos] - '0');
                if (v > ((__int128)1 << 100)) throw Unsupported();
                ++pos;
            }
            if (pos < n && (s[pos] == '.' || s[pos] == 'e' || s[pos] == 'E' || s[pos] == '_' || s[pos] == 'j')) throw Unsupported();
            if (nd < 16) { Rat r; r.n = (i64)v64; r.d = 1; return push_const(r); }       // < 2^53: no reduction needed
            return push_const(make_rat(v, 1));
        }
        if ((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_') {
            const size_t st = pos;
            while (pos < n && ((s[pos] >= 'a' && s[pos] <= 'z') || (s[pos] >= 'A' && s[pos] <= 'Z') || (s[pos] >= '0' && s[pos] <= '9') || s[pos] == '_')) ++pos;
            const char* nm = s + st;
            const size_t nl = pos - st;
            auto is = [&](const std::string& t) { return t.size() == nl && memcmp(t.data(), nm, nl) == 0; };
            if (peek('(')) {
                ++pos;
                const int opc = func_opcode(nm, nl);
                const Val arg = parse_expr(nullptr);
                if (peek(',')) throw Unsupported();
                if (!peek(')')) throw Unsupported();
                ++pos;
                if (opc < 0) throw Unsupported();
                use(arg);
                put((uint8_t)opc);
                Val v; v.start = arg.start;
                return v;
            }
            Val v; v.start = here();
            if (is(sess->var[0])) { put(PDE_OP_VAR0); return v; }
            if (is(sess->var[1])) { put((uint8_t)(PDE_OP_VAR0 + 1)); return v; }
            for (size_t i = 0; i < sess->named.size(); ++i)
                if (is(sess->named[i])) {
                    fixes.push_back({here(), 0, 1, (i64)i, 0});
                    put(PDE_OP_CONST0);
                    return v;
                }
            throw Unsupported();
        }
        throw Unsupported();
    }

    static int func_opcode(const char* f, size_t n) {
        auto eq = [&](const char* t, size_t tn) { return n == tn && memcmp(f, t, tn) == 0; };
        switch (n) {
            case 3:
                if (eq("neg", 3)) return PDE_OP_FN_NEG;
                if (eq("inv", 3)) return PDE_OP_FN_INV;
                if (eq("exp", 3)) return PDE_OP_EXP;
                if (eq("Abs", 3)) return PDE_OP_ABS;
                return -1;
            case 4: return eq("sqrt", 4) ? PDE_OP_SQRT : -1;
            case 6: return eq("square", 6) ? PDE_OP_FN_SQUARE : -1;
            case 7:
                if (eq("pow_3_2", 7)) return PDE_OP_FN_POW32;
                if (eq("exp_neg", 7)) return PDE_OP_FN_EXPNEG;
                return -1;
            case 11: return eq("pow_neg_3_2", 11) ? PDE_OP_FN_POWN32 : -1;
            default: return -1;
        }
    }
};

static bool has_vars(const char* s) {
    // LBF:134-136: ('r' in s) or ('x' in s) or ('rho' in s) or ('z' in s)
    return strchr(s, 'r') || strchr(s, 'x') || strchr(s, 'z');
}

// ---- parallel front end: one worker per contiguous range of strings ----
static int table_find(const pde_session* s, const Fix& f);
constexpr uint32_t kFixResolved = 0xffffffffu;     // Fix::pos of a placeholder whose slot phase A already wrote

struct Worker {
    int lo = 0, hi = 0;
    std::vector<int> pending;           // expressions with a constant / exponent that is not in the tables yet (phase B1)
    std::vector<uint8_t> pool;
    std::vector<int8_t> term_sign;
    std::vector<uint32_t> term_off;     // end offset (in `pool`) of every term
    std::vector<Fix> fixes;
    std::vector<uint32_t> n_terms, pool_end, fix_end;   // per expression (cumulative ends)
    std::vector<uint8_t> flags, attrs;

    void run(const char* blob, const uint32_t* off, const pde_session* sess) {
        const int n = hi - lo;
        n_terms.assign(n, 0); pool_end.assign(n, 0); fix_end.assign(n, 0); flags.assign(n, 0); attrs.assign(n, 0);
        if (n == 0) return;                        // an empty batch has no offsets to read
        // a program is never longer than its source: size the pool for the worker's whole range once
        pool.resize((size_t)(off[hi] - off[lo]) + 64);
        uint8_t* const base = pool.data();
        uint8_t* wp = base;
        term_sign.reserve((size_t)n * 2); term_off.reserve((size_t)n * 2);
        std::vector<TermRec> terms;
        for (int k = 0; k < n; ++k) {
            const char* str = blob + off[lo + k];
            uint8_t* const wp0 = wp;
            const size_t nt0 = term_sign.size(), nf0 = fixes.size();
            try {
                OnePass ps(str, (size_t)(off[lo + k + 1] - off[lo + k] - 1), sess, base, wp, fixes);
                terms.clear();
                ps.parse_expr(&terms);
                ps.ws();
                if (ps.pos != ps.n) throw Unsupported();
                for (const TermRec& t : terms) {
                    term_sign.push_back((int8_t)t.sign);
                    term_off.push_back(t.end);
                }
                // whole program length: bodies + (NEG for a leading minus) + (nterms-1) ADD/SUB
                const size_t whole = (size_t)(wp - wp0) + (terms[0].sign < 0 ? 1 : 0) + (terms.size() - 1);
                if (whole > 255) flags[k] = PDE_FLAG_TOO_LONG;
            } catch (const Unsupported&) {
                flags[k] = PDE_FLAG_UNSUPPORTED;
            }
            if (flags[k]) { wp = wp0; term_sign.resize(nt0); term_off.resize(nt0); fixes.resize(nf0); }
            else {
                // constants / exponents the session tables already hold get their slot here, in parallel (the tables are
                // read-only during phase A); only NEW keys wait for the sequential numbering of phase B1
                bool wait = false;
                for (size_t f = nf0; f < fixes.size(); ++f) {
                    const int slot = table_find(sess, fixes[f]);
                    if (slot < 0) { wait = true; continue; }
                    base[fixes[f].pos] = (uint8_t)((fixes[f].is_pow ? PDE_OP_POW0 : PDE_OP_CONST0) + slot);
                    fixes[f].pos = kFixResolved;
                }
                if (wait) pending.push_back(k);
            }
            n_terms[k] = (uint32_t)(term_sign.size() - nt0);
            pool_end[k] = (uint32_t)(wp - base);
            fix_end[k] = (uint32_t)fixes.size();
            uint8_t attr = 0;                                  // string attributes of the prune predicates (LBF:134-152)
            if (has_vars(str)) attr |= PDE_ATTR_HAS_VARS;
            if (str[0] == '1' && str[1] == '\0') attr |= PDE_ATTR_IS_ONE;
            if (strncmp(str, "inv(", 4) == 0) attr |= PDE_ATTR_STARTS_INV;
            attrs[k] = attr;
        }
        pool.resize((size_t)(wp - base));
    }
};

// numeric mirror of a key table (numerator, denominator; denominator 0 = named constant index), rebuilt when its size
// differs from the key table's
static void ensure_mirror(pde_session* s, bool is_pow) {
    std::vector<std::string>& keys = is_pow ? s->pow_keys : s->const_keys;
    std::vector<long long>& num = is_pow ? s->pow_num : s->const_num;
    std::vector<long long>& den = is_pow ? s->pow_den : s->const_den;
    if (num.size() == keys.size() && den.size() == keys.size()) return;
    num.assign(keys.size(), 0); den.assign(keys.size(), -1);
    for (size_t i = 0; i < keys.size(); ++i) {
        const std::string& k = keys[i];
        bool named = false;
        for (size_t j = 0; j < s->named.size() && !is_pow; ++j) if (s->named[j] == k) { num[i] = (long long)j; den[i] = 0; named = true; }
        if (named) continue;
        const size_t sl = k.find('/');
        num[i] = atoll(k.c_str());
        den[i] = sl == std::string::npos ? 1 : atoll(k.c_str() + sl + 1);
    }
}

// slot of a constant / exponent that is ALREADY in the session tables, or -1 (read-only: phase A calls it from every
// worker while nobody appends)
static int table_find(const pde_session* s, const Fix& f) {
    const std::vector<long long>& num = f.is_pow ? s->pow_num : s->const_num;
    const std::vector<long long>& den = f.is_pow ? s->pow_den : s->const_den;
    const long long fn = f.num, fd = f.named ? 0 : f.den;
    for (size_t i = 0; i < num.size(); ++i) if (num[i] == fn && den[i] == fd) return (int)i;
    return -1;
}

// slot of a constant / exponent in the session tables; appends; -1 = table full
static int table_slot(pde_session* s, const Fix& f) {
    std::vector<std::string>& keys = f.is_pow ? s->pow_keys : s->const_keys;
    std::vector<double>& vals = f.is_pow ? s->pow_vals : s->const_vals;
    std::vector<long long>& num = f.is_pow ? s->pow_num : s->const_num;
    std::vector<long long>& den = f.is_pow ? s->pow_den : s->const_den;
    ensure_mirror(s, f.is_pow);
    const int hit = table_find(s, f);
    if (hit >= 0) return hit;
    const long long fn = f.num, fd = f.named ? 0 : f.den;
    const int cap = f.is_pow ? PDE_N_POW : PDE_N_CONST;
    if ((int)keys.size() >= cap) return -1;
    if (f.named) { keys.push_back(s->named[f.num]); vals.push_back(s->named_vals[f.num]); }
    else { Rat r; r.n = f.num; r.d = f.den; keys.push_back(rat_key(r)); vals.push_back((double)f.num / (double)f.den); }
    num.push_back(fn); den.push_back(fd);
    return (int)keys.size() - 1;
}

static int compile_impl(pde_session* s, const char* blob, const uint32_t* off, int n, pde_exprset** out) {
    std::unique_ptr<pde_exprset> e(new pde_exprset());
    e->n = n;
    e->flags.assign(n, 0);
    e->attrs.assign(n, 0);
    e->term_begin.assign(n + 1, 0);
    e->str_off.assign(off, off + n + (n > 0 ? 1 : 0));
    if (n > 0) e->str_blob.assign(blob, blob + off[n]);
    const bool prof = getenv("PDE_B200_PROFILE") != nullptr;
    auto tnow = [] { return std::chrono::steady_clock::now(); };
    auto t_a = tnow();
    // ---- phase A (parallel): parse + emit; slots of keys the tables already hold, deferred slots for new keys ----
    ensure_mirror(s, false);
    ensure_mirror(s, true);
    int nthreads = 1;
    if (const char* ev = getenv("PDE_B200_COMPILE_THREADS")) nthreads = atoi(ev);
    else nthreads = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
    nthreads = std::max(1, std::min(nthreads, n / 1024 + 1));
    std::vector<Worker> workers(nthreads);
    for (int t = 0; t < nthreads; ++t) {
        workers[t].lo = (int)((long long)n * t / nthreads);
        workers[t].hi = (int)((long long)n * (t + 1) / nthreads);
    }
    if (nthreads == 1) workers[0].run(blob, off, s);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t) th.emplace_back([&, t] { workers[t].run(blob, off, s); });
        for (auto& x : th) x.join();
    }
    auto t_b = tnow();
    // ---- phase B: table slots (sequential, in expression order -- the numbering is part of the bytecode), then the
    //      assembly of the pools at precomputed offsets (parallel again) ----
    // B1: slots of NEW keys.  Only the expressions phase A left pending are visited (none once a session has seen its
    //     constants); an expression whose constants do not fit the tables any more is flagged and its slots are rolled
    //     back (tables stay append-only for successful compiles).
    for (auto& w : workers) {
        for (const int k : w.pending) {
            const uint32_t f0 = k > 0 ? w.fix_end[k - 1] : 0u, f1 = w.fix_end[k];
            const size_t nc0 = s->const_keys.size(), np0 = s->pow_keys.size();
            for (uint32_t f = f0; f < f1; ++f) {
                if (w.fixes[f].pos == kFixResolved) continue;
                const int slot = table_slot(s, w.fixes[f]);
                if (slot < 0) { w.flags[k] = PDE_FLAG_TABLE_FULL; break; }
                w.pool[w.fixes[f].pos] = (uint8_t)((w.fixes[f].is_pow ? PDE_OP_POW0 : PDE_OP_CONST0) + slot);
            }
            if (w.flags[k]) {
                s->const_keys.resize(nc0); s->const_vals.resize(nc0); s->const_num.resize(nc0); s->const_den.resize(nc0);
                s->pow_keys.resize(np0); s->pow_vals.resize(np0); s->pow_num.resize(np0); s->pow_den.resize(np0);
            }
        }
    }
    auto t_b1 = tnow();
    // B2: sizes per worker (expressions flagged in B1 contribute nothing), exclusive offsets
    std::vector<size_t> pool_base(nthreads + 1, 0), term_base(nthreads + 1, 0);
    for (int t = 0; t < nthreads; ++t) {
        Worker& w = workers[t];
        size_t pb = 0, tb = 0;
        uint32_t p0 = 0;
        for (int k = 0; k < w.hi - w.lo; ++k) {
            if (!w.flags[k]) { pb += w.pool_end[k] - p0; tb += w.n_terms[k]; }
            p0 = w.pool_end[k];
        }
        pool_base[t + 1] = pool_base[t] + pb;
        term_base[t + 1] = term_base[t] + tb;
    }
    e->pool.resize(pool_base[nthreads]);
    e->term_sign.resize(term_base[nthreads]);
    e->term_off.resize(term_base[nthreads] + 1);
    e->term_off[0] = 0;
    auto t_b2 = tnow();
    // B3: copy (parallel)
    auto assemble = [&](int t) {
        Worker& w = workers[t];
        size_t pb = pool_base[t], tb = term_base[t];
        uint32_t p0 = 0, t0 = 0;
        for (int k = 0; k < w.hi - w.lo; ++k) {
            const int i = w.lo + k;
            const uint32_t p1 = w.pool_end[k], nt = w.n_terms[k];
            e->attrs[i] = w.attrs[k];
            e->flags[i] = w.flags[k];
            e->term_begin[i] = (uint32_t)tb;
            if (!w.flags[k]) {
                memcpy(e->pool.data() + pb, w.pool.data() + p0, p1 - p0);
                for (uint32_t q = 0; q < nt; ++q) {
                    e->term_sign[tb + q] = w.term_sign[t0 + q];
                    e->term_off[tb + q + 1] = (uint32_t)(pb + (w.term_off[t0 + q] - p0));
                }
                pb += p1 - p0; tb += nt;
            }
            p0 = p1; t0 += nt;
        }
    };
    if (nthreads == 1) assemble(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t) th.emplace_back(assemble, t);
        for (auto& x : th) x.join();
    }
    e->term_begin[n] = (uint32_t)term_base[nthreads];
    if (prof) {
        auto t_c = tnow();
        fprintf(stderr, "[pde_compile] n=%d threads=%d parse %.2f ms, slots+assembly %.2f ms (B1 %.2f B2 %.2f B3 %.2f)\n", n, nthreads,
                std::chrono::duration<double, std::milli>(t_b - t_a).count(), std::chrono::duration<double, std::milli>(t_c - t_b).count(),
                std::chrono::duration<double, std::milli>(t_b1 - t_b).count(), std::chrono::duration<double, std::milli>(t_b2 - t_b1).count(), std::chrono::duration<double, std::milli>(t_c - t_b2).count());
    }
    *out = e.release();
    return PDE_OK;
}

template <typename T>
static int upload(T** dptr, const std::vector<T>& v) {
    size_t bytes = sizeof(T) * (v.empty() ? 1 : v.size());
    if (int rc = pde::scratch_alloc((void**)dptr, bytes, nullptr)) return rc;     // legacy stream: ordered before the copy below
    cudaError_t e;
    if (!v.empty()) {
        e = cudaMemcpy(*dptr, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return pde::cuda_fail((int)e, "cudaMemcpy(exprset)");
    }
    return PDE_OK;
}

}  // namespace

namespace pde {

// Dense lexicographic rank (Python str comparison == byte comparison for ASCII): only the enumerator's
// `a > b` / `a == b` tests need it (LBF:168-195), so it is computed on first use, not by every compile.
void exprset_ensure_rank(pde_exprset* e) {
    if (e->rank_ready) return;
    const int n = e->n;
    e->rank.assign(n, 0);
    const char* blob = e->str_blob.data();
    // sort by the first 8 bytes as one big-endian integer; strcmp only on ties, from byte 8 on (a key with a NUL byte in
    // it is a whole string): 3-4x faster than strcmp for every comparison, and this sort is on the critical path of the
    // first enumerate call of every handle
    struct Key { uint64_t k; int i; };
    std::vector<Key> order(n);
    for (int i = 0; i < n; ++i) {
        const unsigned char* s = reinterpret_cast<const unsigned char*>(blob + e->str_off[i]);
        uint64_t k = 0;
        int j = 0;
        for (; j < 8 && s[j]; ++j) k = (k << 8) | s[j];
        k <<= 8 * (8 - j);
        order[i] = Key{k, i};
    }
    auto tail_cmp = [&](const Key& a, const Key& b) {
        if ((a.k & 0xffULL) == 0) return 0;                       // shorter than 8 bytes: the key is the string
        return strcmp(blob + e->str_off[a.i] + 8, blob + e->str_off[b.i] + 8);
    };
    std::sort(order.begin(), order.end(), [&](const Key& a, const Key& b) {
        if (a.k != b.k) return a.k < b.k;
        const int c = tail_cmp(a, b);
        return c < 0 || (c == 0 && a.i < b.i);
    });
    uint32_t r = 0;
    for (int k = 0; k < n; ++k) {
        if (k > 0 && (order[k].k != order[k - 1].k || tail_cmp(order[k], order[k - 1]) != 0)) ++r;
        e->rank[order[k].i] = r;
    }
    e->rank_ready = true;
}

// Device mirrors of the operand set: needed by the enumerator only, uploaded on first use
// to the device current at that time.
int exprset_ensure_device(pde_exprset* e) {
    if (e->device >= 0) return PDE_OK;
    if (!have_device()) { set_error("exprset has no device mirror: no CUDA device"); return PDE_E_NODEVICE; }
    const bool prof = getenv("PDE_B200_PROFILE") != nullptr;
    auto tnow = [] { return std::chrono::steady_clock::now(); };
    auto t_a = tnow();
    exprset_ensure_rank(e);
    auto t_b = tnow();
    // splice descriptors + whole programs (enumerate.cu): whole(i) = t1 [NEG] (tk ADD|SUB)*
    {
        e->desc.assign((size_t)e->n * 2, 0);
        e->wpool.clear();
        e->wpool.reserve(e->pool.size() + e->term_sign.size() + 8);
        for (int i = 0; i < e->n; ++i) {
            const uint32_t t0 = e->term_begin[i], t1 = e->term_begin[i + 1];
            const uint32_t off = (uint32_t)e->wpool.size();
            unsigned fl = 0, first_len = 0, last_off = 0;
            if (e->flags[i] || t1 == t0) fl = 8;                 // D_BAD
            else {
                for (uint32_t t = t0; t < t1; ++t) {
                    const uint32_t b0 = e->term_off[t], b1 = e->term_off[t + 1];
                    if (t == t1 - 1) last_off = (unsigned)(e->wpool.size() - off);
                    e->wpool.insert(e->wpool.end(), e->pool.begin() + b0, e->pool.begin() + b1);
                    if (t == t0) { first_len = b1 - b0; if (e->term_sign[t] < 0) e->wpool.push_back(PDE_OP_NEG); }
                    else e->wpool.push_back(e->term_sign[t] > 0 ? PDE_OP_ADD : PDE_OP_SUB);
                }
                if (e->term_sign[t0] < 0) fl |= 1;               // D_FIRST_NEG
                if (e->term_sign[t1 - 1] < 0) fl |= 2;           // D_LAST_NEG
                if (t1 - t0 > 1) fl |= 4;                        // D_MULTI
            }
            const unsigned len = (unsigned)(e->wpool.size() - off);   // <= 255 (longer programs are flagged TOO_LONG)
            e->desc[2 * (size_t)i] = off;
            e->desc[2 * (size_t)i + 1] = len | (first_len << 8) | (last_off << 16) | (fl << 24);
        }
        e->wpool.resize(e->wpool.size() + 8, 0);                 // the word-wise reader looks one word ahead
    }
    auto t_c = tnow();
    int dev = 0;
    cudaGetDevice(&dev);
    int rc;
    if ((rc = upload(&e->d_desc, e->desc))) return rc;
    if ((rc = upload(&e->d_wpool, e->wpool))) return rc;
    if ((rc = upload(&e->d_flags, e->flags))) return rc;
    if ((rc = upload(&e->d_attrs, e->attrs))) return rc;
    if ((rc = upload(&e->d_rank, e->rank))) return rc;
    if ((rc = upload(&e->d_term_begin, e->term_begin))) return rc;
    if ((rc = upload(&e->d_term_sign, e->term_sign))) return rc;
    if ((rc = upload(&e->d_term_off, e->term_off))) return rc;
    if ((rc = upload(&e->d_pool, e->pool))) return rc;
    if (prof) {
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "[exprset_ensure_device] n=%d rank %.2f ms, descriptors %.2f ms, uploads %.2f ms\n", e->n, ms(t_a, t_b), ms(t_b, t_c), ms(t_c, tnow()));
    }
    e->device = dev;
    return PDE_OK;
}

}  // namespace pde

extern "C" {

int pde_compile_exprs(pde_session* s, const char* const* strs, int n, pde_exprset** out) {
    if (!s || !out || n < 0 || (n > 0 && !strs)) { pde::set_error("pde_compile_exprs: bad argument"); return PDE_E_INVALID; }
    std::string blob;
    std::vector<uint32_t> off(n + 1, 0);
    size_t total = 0;
    for (int i = 0; i < n; ++i) total += strlen(strs[i]) + 1;
    if (total >= 0xffffffffULL) { pde::set_error("pde_compile_exprs: input too large"); return PDE_E_OVERFLOW; }
    blob.reserve(total);
    for (int i = 0; i < n; ++i) { off[i] = (uint32_t)blob.size(); blob.append(strs[i]); blob.push_back('\0'); }
    off[n] = (uint32_t)blob.size();
    return compile_impl(s, blob.data(), off.data(), n, out);
}

int pde_compile_exprs_blob(pde_session* s, const char* blob, size_t blob_bytes, int n, pde_exprset** out) {
    if (!s || !out || n < 0 || (n > 0 && !blob)) { pde::set_error("pde_compile_exprs_blob: bad argument"); return PDE_E_INVALID; }
    if (blob_bytes >= 0xffffffffULL) { pde::set_error("pde_compile_exprs_blob: input too large"); return PDE_E_OVERFLOW; }
    // exactly n NUL-terminated strings in blob_bytes bytes: the terminators are found here, inside the size the caller
    // states (memchr at memory speed; counting them in Python -- bytes.count -- cost 3-8 ms for the 4.7 MB of the depth-4
    // uniques, as much as parsing a quarter of them)
    std::vector<uint32_t> found((size_t)n + 1);
    const char* p = blob;
    const char* end = blob + blob_bytes;
    for (int i = 0; i < n; ++i) {
        found[i] = (uint32_t)(p - blob);
        const char* q = p < end ? (const char*)memchr(p, 0, (size_t)(end - p)) : nullptr;
        if (!q) { pde::set_error("pde_compile_exprs_blob: %d strings expected, the blob holds %d", n, i); return PDE_E_INVALID; }
        p = q + 1;
    }
    found[n] = (uint32_t)(p - blob);
    if (p != end) { pde::set_error("pde_compile_exprs_blob: the blob holds more than the %d strings expected (a string with a NUL inside?)", n); return PDE_E_INVALID; }
    return pde_compile_exprs_packed(s, blob, found.data(), n, out);
}

int pde_compile_exprs_packed(pde_session* s, const char* blob, const uint32_t* offsets, int n, pde_exprset** out) {
    if (!s || !out || n < 0 || (n > 0 && !blob)) { pde::set_error("pde_compile_exprs_packed: bad argument"); return PDE_E_INVALID; }
    std::vector<uint32_t> found;
    if (!offsets && n > 0) {
        // offsets = NULL: the blob is n NUL-terminated strings back to back; find the terminators here (memchr runs at
        // memory speed; numpy needed 5-10 ms for the 4.7 MB of the depth-4 uniques)
        found.resize((size_t)n + 1);
        const char* p = blob;
        for (int i = 0; i < n; ++i) {
            found[i] = (uint32_t)(p - blob);
            const char* q = (const char*)memchr(p, 0, (size_t)0xffffffffu - (size_t)(p - blob));
            if (!q) { pde::set_error("pde_compile_exprs_packed: string %d is not NUL terminated", i); return PDE_E_INVALID; }
            p = q + 1;
            if ((size_t)(p - blob) >= 0xffffffffULL) { pde::set_error("pde_compile_exprs_packed: input too large"); return PDE_E_OVERFLOW; }
        }
        found[n] = (uint32_t)(p - blob);
        offsets = found.data();
    }
    for (int i = 0; i < n; ++i) {
        if (offsets[i + 1] <= offsets[i] || blob[offsets[i + 1] - 1] != '\0') {
            pde::set_error("pde_compile_exprs_packed: string %d is not NUL terminated at offsets[%d] - 1", i, i + 1);
            return PDE_E_INVALID;
        }
    }
    return compile_impl(s, blob, offsets, n, out);
}

void pde_exprset_free(pde_exprset* e) {
    if (!e) return;
    if (e->device >= 0) {
        // stream-ordered: the legacy stream waits for the last kernel that read the mirrors, then the buffers go back to
        // the library's pool (no device-wide synchronisation, unlike cudaFree)
        int cur = -1;
        cudaGetDevice(&cur);
        if (cur != e->device) cudaSetDevice(e->device);          // the pool and the legacy stream of the mirrors' device
        if (e->used_event) cudaStreamWaitEvent(nullptr, (cudaEvent_t)e->used_event, 0);
        void* bufs[] = {e->d_flags, e->d_attrs, e->d_rank, e->d_term_begin, e->d_term_sign, e->d_term_off, e->d_pool, e->d_desc, e->d_wpool,
                        e->d_count_sums, e->d_count_in_tile, e->d_count_tile, e->d_bytes_sums, e->d_bytes_in_tile, e->d_bytes_tile};
        for (void* b : bufs) pde::scratch_free(b, nullptr);
        if (cur >= 0 && cur != e->device) cudaSetDevice(cur);
    }
    if (e->used_event) cudaEventDestroy((cudaEvent_t)e->used_event);
    delete e;
}

int pde_exprset_size(const pde_exprset* e, int* n_expr, int* n_terms, int* n_pool_bytes) {
    if (!e) { pde::set_error("null exprset"); return PDE_E_INVALID; }
    if (n_expr) *n_expr = e->n;
    if (n_terms) *n_terms = (int)e->term_sign.size();
    if (n_pool_bytes) *n_pool_bytes = (int)e->pool.size();
    return PDE_OK;
}

int pde_exprset_export(const pde_exprset* e, uint8_t* flags, uint8_t* attrs, uint32_t* rank,
                       uint32_t* term_begin, int8_t* term_sign, uint32_t* term_off, uint8_t* pool) {
    if (!e) { pde::set_error("null exprset"); return PDE_E_INVALID; }
    if (flags) memcpy(flags, e->flags.data(), e->flags.size());
    if (attrs) memcpy(attrs, e->attrs.data(), e->attrs.size());
    if (rank) { pde::exprset_ensure_rank(const_cast<pde_exprset*>(e)); memcpy(rank, e->rank.data(), e->rank.size() * 4); }
    if (term_begin) memcpy(term_begin, e->term_begin.data(), e->term_begin.size() * 4);
    if (term_sign) memcpy(term_sign, e->term_sign.data(), e->term_sign.size());
    if (term_off) memcpy(term_off, e->term_off.data(), e->term_off.size() * 4);
    if (pool) memcpy(pool, e->pool.data(), e->pool.size());
    return PDE_OK;
}

int pde_exprset_programs(const pde_exprset* e, int L, uint8_t* code, uint8_t* len) {
    if (!e || !code || !len || L < 1 || L > 256) { pde::set_error("pde_exprset_programs: bad argument"); return PDE_E_INVALID; }
    auto fill = [&](int lo, int hi) {
        memset(code + (size_t)lo * L, 0, (size_t)(hi - lo) * L);
        for (int i = lo; i < hi; ++i) {
            len[i] = 0;
            if (e->flags[i]) continue;
            uint8_t buf[512];
            int w = 0;
            for (uint32_t t = e->term_begin[i]; t < e->term_begin[i + 1]; ++t) {
                uint32_t b0 = e->term_off[t], b1 = e->term_off[t + 1];
                memcpy(buf + w, e->pool.data() + b0, b1 - b0);
                w += (int)(b1 - b0);
                if (t == e->term_begin[i]) { if (e->term_sign[t] < 0) buf[w++] = PDE_OP_NEG; }
                else buf[w++] = e->term_sign[t] > 0 ? PDE_OP_ADD : PDE_OP_SUB;
            }
            if (w > L || w > 255) continue;
            memcpy(code + (size_t)i * L, buf, w);
            len[i] = (uint8_t)w;
        }
    };
    int nthreads = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
    if (const char* ev = getenv("PDE_B200_COMPILE_THREADS")) nthreads = atoi(ev);
    nthreads = std::max(1, std::min(nthreads, e->n / 8192 + 1));
    if (nthreads == 1) fill(0, e->n);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t)
            th.emplace_back(fill, (int)((long long)e->n * t / nthreads), (int)((long long)e->n * (t + 1) / nthreads));
        for (auto& x : th) x.join();
    }
    return PDE_OK;
}

}  // extern "C"
