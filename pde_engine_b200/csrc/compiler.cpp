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
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <memory>
#include <numeric>
#include <string>
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
    // reduce with 128-bit gcd
    __int128 a = n < 0 ? -n : n, b = d;
    while (b) { __int128 t = a % b; a = b; b = t; }
    if (a == 0) a = 1;
    n /= a; d /= a;
    if (n > LIMIT || n < -LIMIT || d > LIMIT) throw Unsupported();
    Rat r; r.n = (i64)n; r.d = (i64)d;
    return r;
}

enum Kind { K_CONST, K_NCONST, K_VAR, K_NEG, K_BIN, K_POW, K_CALL };

struct Node {
    Kind kind;
    Rat rat;            // K_CONST value, K_POW exponent
    int idx = 0;        // K_VAR index, K_NCONST named index, K_CALL opcode
    char op = 0;        // K_BIN: + - * /
    Node* a = nullptr;
    Node* b = nullptr;
};

struct Arena {
    std::vector<std::unique_ptr<Node>> nodes;
    Node* make(Kind k) {
        nodes.emplace_back(new Node());
        nodes.back()->kind = k;
        return nodes.back().get();
    }
};

struct Parser {
    const char* s;
    size_t pos = 0, n;
    const pde_session* sess;
    Arena& ar;
    int depth = 0;
    Parser(const char* str, const pde_session* se, Arena& a) : s(str), n(strlen(str)), sess(se), ar(a) {}

    void ws() { while (pos < n && (s[pos] == ' ' || s[pos] == '\t')) ++pos; }
    bool peek(char c) { ws(); return pos < n && s[pos] == c; }
    bool peek2(const char* t) { ws(); return pos + 1 < n && s[pos] == t[0] && s[pos + 1] == t[1]; }

    Node* constant(Rat r) { Node* x = ar.make(K_CONST); x->rat = r; return x; }

    Node* parse_expr() {
        if (++depth > 200) throw Unsupported();
        Node* l = parse_term();
        for (;;) {
            ws();
            if (pos < n && (s[pos] == '+' || s[pos] == '-')) {
                char op = s[pos++];
                Node* r = parse_term();
                l = binop(op, l, r);
            } else break;
        }
        --depth;
        return l;
    }
    Node* parse_term() {
        Node* l = parse_factor();
        for (;;) {
            ws();
            if (pos < n && (s[pos] == '*' || s[pos] == '/') && !(pos + 1 < n && s[pos] == '*' && s[pos + 1] == '*')) {
                if (s[pos] == '/' && pos + 1 < n && s[pos + 1] == '/') throw Unsupported();
                char op = s[pos++];
                Node* r = parse_factor();
                l = binop(op, l, r);
            } else break;
        }
        return l;
    }
    Node* parse_factor() {
        ws();
        if (++depth > 200) throw Unsupported();
        Node* res;
        if (pos < n && s[pos] == '+') { ++pos; res = parse_factor(); }
        else if (pos < n && s[pos] == '-') {
            ++pos;
            Node* x = parse_factor();
            if (x->kind == K_CONST) { Rat r = x->rat; r.n = -r.n; res = constant(r); }
            else { res = ar.make(K_NEG); res->a = x; }
        } else res = parse_power();
        --depth;
        return res;
    }
    Node* parse_power() {
        Node* base = parse_atom();
        if (peek2("**")) {
            pos += 2;
            Node* e = parse_factor();
            if (e->kind != K_CONST) throw Unsupported();
            Rat k = e->rat;
            if (base->kind == K_CONST && k.d == 1) {
                if (k.n > 64 || k.n < -64) throw Unsupported();
                if (base->rat.n == 0 && k.n < 0) throw Unsupported();
                __int128 nn = 1, dd = 1;
                i64 e2 = k.n < 0 ? -k.n : k.n;
                for (i64 i = 0; i < e2; ++i) {
                    nn *= base->rat.n; dd *= base->rat.d;
                    if (nn > ((__int128)1 << 100) || nn < -((__int128)1 << 100) || dd > ((__int128)1 << 100)) throw Unsupported();
                }
                return constant(k.n < 0 ? make_rat(dd, nn) : make_rat(nn, dd));
            }
            Node* p = ar.make(K_POW);
            p->a = base; p->rat = k;
            return p;
        }
        return base;
    }
    Node* parse_atom() {
        ws();
        if (pos >= n) throw Unsupported();
        char c = s[pos];
        if (c == '(') {
            ++pos;
            Node* e = parse_expr();
            if (!peek(')')) throw Unsupported();
            ++pos;
            return e;
        }
        if (c >= '0' && c <= '9') {
            __int128 v = 0;
            size_t st = pos;
            while (pos < n && s[pos] >= '0' && s[pos] <= '9') {
                v = v * 10 + (s[pos] - '0');
                if (v > ((__int128)1 << 100)) throw Unsupported();
                ++pos;
            }
            if (pos < n && (s[pos] == '.' || s[pos] == 'e' || s[pos] == 'E' || s[pos] == '_' || s[pos] == 'j')) throw Unsupported();
            (void)st;
            return constant(make_rat(v, 1));
        }
        if ((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_') {
            size_t st = pos;
            while (pos < n && ((s[pos] >= 'a' && s[pos] <= 'z') || (s[pos] >= 'A' && s[pos] <= 'Z') || (s[pos] >= '0' && s[pos] <= '9') || s[pos] == '_')) ++pos;
            std::string name(s + st, pos - st);
            if (peek('(')) {
                ++pos;
                int opc = func_opcode(name);
                Node* arg = parse_expr();
                if (peek(',')) throw Unsupported();
                if (!peek(')')) throw Unsupported();
                ++pos;
                if (opc < 0) throw Unsupported();
                Node* x = ar.make(K_CALL);
                x->idx = opc; x->a = arg;
                return x;
            }
            if (name == sess->var[0]) { Node* x = ar.make(K_VAR); x->idx = 0; return x; }
            if (name == sess->var[1]) { Node* x = ar.make(K_VAR); x->idx = 1; return x; }
            for (size_t i = 0; i < sess->named.size(); ++i)
                if (name == sess->named[i]) { Node* x = ar.make(K_NCONST); x->idx = (int)i; return x; }
            throw Unsupported();
        }
        throw Unsupported();
    }
    static int func_opcode(const std::string& f) {
        if (f == "neg") return PDE_OP_FN_NEG;
        if (f == "inv") return PDE_OP_FN_INV;
        if (f == "square") return PDE_OP_FN_SQUARE;
        if (f == "pow_3_2") return PDE_OP_FN_POW32;
        if (f == "pow_neg_3_2") return PDE_OP_FN_POWN32;
        if (f == "exp_neg") return PDE_OP_FN_EXPNEG;
        if (f == "sqrt") return PDE_OP_SQRT;
        if (f == "exp") return PDE_OP_EXP;
        if (f == "Abs") return PDE_OP_ABS;
        return -1;
    }
    Node* binop(char op, Node* l, Node* r) {
        if (l->kind == K_CONST && r->kind == K_CONST) {
            __int128 an = l->rat.n, ad = l->rat.d, bn = r->rat.n, bd = r->rat.d;
            switch (op) {
                case '+': return constant(make_rat(an * bd + bn * ad, ad * bd));
                case '-': return constant(make_rat(an * bd - bn * ad, ad * bd));
                case '*': return constant(make_rat(an * bn, ad * bd));
                default:
                    if (bn == 0) throw Unsupported();
                    return constant(make_rat(an * bd, ad * bn));
            }
        }
        Node* x = ar.make(K_BIN);
        x->op = op; x->a = l; x->b = r;
        return x;
    }
};

static std::string rat_key(const Rat& r) {
    return r.d == 1 ? std::to_string(r.n) : std::to_string(r.n) + "/" + std::to_string(r.d);
}

struct Emitter {
    pde_session* sess;
    Arena& ar;

    int const_slot(const std::string& key, double val) {
        for (size_t i = 0; i < sess->const_keys.size(); ++i)
            if (sess->const_keys[i] == key) return (int)i;
        if ((int)sess->const_keys.size() >= PDE_N_CONST) throw TableFull();
        sess->const_keys.push_back(key);
        sess->const_vals.push_back(val);
        return (int)sess->const_keys.size() - 1;
    }
    int pow_slot(const std::string& key, double val) {
        for (size_t i = 0; i < sess->pow_keys.size(); ++i)
            if (sess->pow_keys[i] == key) return (int)i;
        if ((int)sess->pow_keys.size() >= PDE_N_POW) throw TableFull();
        sess->pow_keys.push_back(key);
        sess->pow_vals.push_back(val);
        return (int)sess->pow_keys.size() - 1;
    }

    struct Term { int sign; Node* body; };

    // leading unary minus: leftmost leaf of the * / chain
    Node* extract_sign(Node* t, int& sign) {
        if (t->kind == K_NEG) { sign = -sign; return extract_sign(t->a, sign); }
        if (t->kind == K_BIN && (t->op == '*' || t->op == '/')) {
            Node* l2 = extract_sign(t->a, sign);
            if (l2 != t->a) {
                Node* x = ar.make(K_BIN);
                x->op = t->op; x->a = l2; x->b = t->b;
                return x;
            }
        }
        return t;
    }

    void split_terms(Node* ir, std::vector<Term>& out) {
        std::vector<Term> chain;
        while (ir->kind == K_BIN && (ir->op == '+' || ir->op == '-')) {
            chain.push_back({ir->op == '+' ? 1 : -1, ir->b});
            ir = ir->a;
        }
        chain.push_back({1, ir});
        for (size_t i = chain.size(); i-- > 0;) {
            int sign = chain[i].sign;
            Node* body = extract_sign(chain[i].body, sign);
            out.push_back({sign, body});
        }
    }

    void emit(Node* ir, std::vector<uint8_t>& out) {
        std::vector<Term> terms;
        split_terms(ir, terms);
        for (size_t k = 0; k < terms.size(); ++k) {
            emit_term(terms[k].body, out);
            if (k == 0) { if (terms[k].sign < 0) out.push_back(PDE_OP_NEG); }
            else out.push_back(terms[k].sign > 0 ? PDE_OP_ADD : PDE_OP_SUB);
        }
    }

    void emit_term(Node* t, std::vector<uint8_t>& out) {
        switch (t->kind) {
            case K_CONST:
                out.push_back((uint8_t)(PDE_OP_CONST0 + const_slot(rat_key(t->rat), (double)t->rat.n / (double)t->rat.d)));
                break;
            case K_NCONST:
                out.push_back((uint8_t)(PDE_OP_CONST0 + const_slot(sess->named[t->idx], sess->named_vals[t->idx])));
                break;
            case K_VAR: out.push_back((uint8_t)(PDE_OP_VAR0 + t->idx)); break;
            case K_NEG: emit(t->a, out); out.push_back(PDE_OP_NEG); break;
            case K_BIN:
                if (t->op == '+' || t->op == '-') emit(t, out);
                else { emit(t->a, out); emit(t->b, out); out.push_back(t->op == '*' ? PDE_OP_MUL : PDE_OP_DIV); }
                break;
            case K_POW:
                emit(t->a, out);
                out.push_back((uint8_t)(PDE_OP_POW0 + pow_slot(rat_key(t->rat), (double)t->rat.n / (double)t->rat.d)));
                break;
            case K_CALL: emit(t->a, out); out.push_back((uint8_t)t->idx); break;
        }
    }
};

static bool has_vars(const char* s) {
    // LBF:134-136: ('r' in s) or ('x' in s) or ('rho' in s) or ('z' in s)
    return strchr(s, 'r') || strchr(s, 'x') || strchr(s, 'z');
}

template <typename T>
static int upload(T** dptr, const std::vector<T>& v) {
    size_t bytes = sizeof(T) * (v.empty() ? 1 : v.size());
    cudaError_t e = cudaMalloc((void**)dptr, bytes);
    if (e != cudaSuccess) return pde::cuda_fail((int)e, "cudaMalloc(exprset)");
    if (!v.empty()) {
        e = cudaMemcpy(*dptr, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return pde::cuda_fail((int)e, "cudaMemcpy(exprset)");
    }
    return PDE_OK;
}

}  // namespace

extern "C" {

int pde_compile_exprs(pde_session* s, const char* const* strs, int n, pde_exprset** out) {
    if (!s || !out || n < 0 || (n > 0 && !strs)) { pde::set_error("pde_compile_exprs: bad argument"); return PDE_E_INVALID; }
    std::unique_ptr<pde_exprset> e(new pde_exprset());
    e->n = n;
    e->flags.assign(n, 0);
    e->attrs.assign(n, 0);
    e->rank.assign(n, 0);
    e->term_begin.assign(n + 1, 0);
    e->term_off.push_back(0);
    for (int i = 0; i < n; ++i) {
        const char* str = strs[i];
        e->term_begin[i] = (uint32_t)e->term_sign.size();
        uint8_t attr = 0;
        if (has_vars(str)) attr |= PDE_ATTR_HAS_VARS;
        if (strcmp(str, "1") == 0) attr |= PDE_ATTR_IS_ONE;
        if (strncmp(str, "inv(", 4) == 0) attr |= PDE_ATTR_STARTS_INV;
        e->attrs[i] = attr;
        const size_t pool0 = e->pool.size(), nt0 = e->term_sign.size();
        const size_t nc0 = s->const_keys.size(), np0 = s->pow_keys.size();
        try {
            Arena ar;
            Parser ps(str, s, ar);
            Node* ir = ps.parse_expr();
            ps.ws();
            if (ps.pos != ps.n) throw Unsupported();
            Emitter em{s, ar};
            std::vector<Emitter::Term> terms;
            em.split_terms(ir, terms);
            size_t total = 0;
            for (auto& t : terms) {
                std::vector<uint8_t> body;
                em.emit_term(t.body, body);
                e->pool.insert(e->pool.end(), body.begin(), body.end());
                e->term_sign.push_back((int8_t)t.sign);
                e->term_off.push_back((uint32_t)e->pool.size());
                total += body.size() + 1;
            }
            // whole program length: bodies + (NEG for a leading minus) + (nterms-1) ADD/SUB
            size_t whole = e->pool.size() - pool0 + (terms[0].sign < 0 ? 1 : 0) + (terms.size() - 1);
            if (whole > 255) {
                e->flags[i] = PDE_FLAG_TOO_LONG;
                throw 0;
            }
        } catch (const TableFull&) {
            e->flags[i] = PDE_FLAG_TABLE_FULL;
        } catch (const Unsupported&) {
            e->flags[i] = PDE_FLAG_UNSUPPORTED;
        } catch (int) {
        }
        if (e->flags[i]) {
            // roll back partial output (tables stay append-only only for successful compiles)
            e->pool.resize(pool0);
            e->term_sign.resize(nt0);
            e->term_off.resize(nt0 + 1);
            s->const_keys.resize(nc0); s->const_vals.resize(nc0);
            s->pow_keys.resize(np0); s->pow_vals.resize(np0);
        }
    }
    e->term_begin[n] = (uint32_t)e->term_sign.size();
    // dense lexicographic rank (Python str comparison == byte comparison for ASCII)
    {
        std::vector<int> order(n);
        std::iota(order.begin(), order.end(), 0);
        std::sort(order.begin(), order.end(), [&](int a, int b) {
            int c = strcmp(strs[a], strs[b]);
            return c < 0 || (c == 0 && a < b);
        });
        uint32_t r = 0;
        for (int k = 0; k < n; ++k) {
            if (k > 0 && strcmp(strs[order[k]], strs[order[k - 1]]) != 0) ++r;
            e->rank[order[k]] = r;
        }
    }
    // device mirrors (skipped when there is no device: host-only use in CPU tests)
    if (pde::have_device()) {
        cudaGetDevice(&e->device);
        int rc;
        if ((rc = upload(&e->d_flags, e->flags))) return rc;
        if ((rc = upload(&e->d_attrs, e->attrs))) return rc;
        if ((rc = upload(&e->d_rank, e->rank))) return rc;
        if ((rc = upload(&e->d_term_begin, e->term_begin))) return rc;
        if ((rc = upload(&e->d_term_sign, e->term_sign))) return rc;
        if ((rc = upload(&e->d_term_off, e->term_off))) return rc;
        if ((rc = upload(&e->d_pool, e->pool))) return rc;
    }
    *out = e.release();
    return PDE_OK;
}

void pde_exprset_free(pde_exprset* e) {
    if (!e) return;
    if (e->device >= 0) {
        cudaFree(e->d_flags); cudaFree(e->d_attrs); cudaFree(e->d_rank); cudaFree(e->d_term_begin);
        cudaFree(e->d_term_sign); cudaFree(e->d_term_off); cudaFree(e->d_pool);
    }
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
    if (rank) memcpy(rank, e->rank.data(), e->rank.size() * 4);
    if (term_begin) memcpy(term_begin, e->term_begin.data(), e->term_begin.size() * 4);
    if (term_sign) memcpy(term_sign, e->term_sign.data(), e->term_sign.size());
    if (term_off) memcpy(term_off, e->term_off.data(), e->term_off.size() * 4);
    if (pool) memcpy(pool, e->pool.data(), e->pool.size());
    return PDE_OK;
}

int pde_exprset_programs(const pde_exprset* e, int L, uint8_t* code, uint8_t* len) {
    if (!e || !code || !len || L < 1 || L > 256) { pde::set_error("pde_exprset_programs: bad argument"); return PDE_E_INVALID; }
    memset(code, 0, (size_t)e->n * L);
    for (int i = 0; i < e->n; ++i) {
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
    return PDE_OK;
}

}  // extern "C"
