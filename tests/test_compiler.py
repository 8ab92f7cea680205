"""Host compiler (C++) vs the oracle parser (Python ast): byte-for-byte (CPU)."""
import numpy as np
import pytest

from conftest import uniques_by_depth
from oracle import enumerate as oe
from oracle import parser as op

EDGE = ["I*rho", "zoo", "rho**z", "2**(1/2)*rho", "-(rho+z)", "- -rho", "+rho", "((rho))", "1/3 - 1/3*rho",
        "rho**-2", "2**-1*z", "foo(rho)", "rho +", "(rho", "rho)", "1.5*rho", "Abs(z)", "exp(-2*z)", "-1",
        "-rho**2*z", "E*rho", "10**30*rho", "3**40", "", "rho//z", "rho**2**2", "-rho**-2", "1/0*rho", "0**-1",
        "rho*(-1 + 1/z)", "(rho**2*(rho - z) - z**3)/(rho - z)", "rho - -z + 1", "exp_neg(pow_3_2(rho/z))"]


def _compare(problem, strs):
    import pde_engine_b200 as pb
    sess = pb.Session.for_problem(problem)
    osess = op.Session.for_problem(problem)
    es = sess.compile(strs)
    ex = es.export()
    for i, s in enumerate(strs):
        c = op.compile_expr(s, osess)
        assert int(ex["flags"][i]) == c.flags, s
        if c.flags:
            continue
        tb, te = ex["term_begin"][i], ex["term_begin"][i + 1]
        terms = [(int(ex["term_sign"][t]), bytes(ex["pool"][ex["term_off"][t]:ex["term_off"][t + 1]])) for t in range(tb, te)]
        assert terms == c.terms, s
    cv, nc, pv, npw = sess.tables()
    assert sess.const_keys() == osess.const_keys and sess.pow_keys() == osess.pow_keys
    assert np.array_equal(cv[:nc], np.array(osess.const_vals)) and np.array_equal(pv[:npw], np.array(osess.pow_vals))
    return es, ex


def test_compiler_force_free(enum_ff):
    E = uniques_by_depth(enum_ff)
    strs = E[1] + E[2] + E[3] + E[4][::37] + enum_ff["depths"]["2"]["candidates"] + enum_ff["depths"]["3"]["candidates"][::5] + EDGE
    es, ex = _compare("force_free", strs)
    # string predicates + lexicographic rank (LBF:134-136,143-152,168-169)
    for i, s in enumerate(strs):
        a = (1 if oe.has_vars(s) else 0) | (2 if s == "1" else 0) | (4 if s.startswith("inv(") else 0)
        assert a == ex["attrs"][i], s
    rk = {s: i for i, s in enumerate(sorted(set(strs)))}
    assert all(rk[s] == ex["rank"][i] for i, s in enumerate(strs))
    # whole programs
    osess = op.Session.for_problem("force_free")
    code, ln = es.programs(64)
    for i, s in enumerate(strs[:600]):
        c = op.compile_expr(s, osess)
        if c.flags or len(c.whole()) > 64:
            assert ln[i] == 0
        else:
            w = c.whole()
            assert ln[i] == len(w) and bytes(code[i, :len(w)]) == w and not code[i, len(w):].any()


def test_compiler_kerr(enum_kerr):
    E = uniques_by_depth(enum_kerr)
    strs = E[1] + E[2] + E[3][::9] + enum_kerr["depths"]["2"]["candidates"] + EDGE
    _compare("kerr_magnetosphere", strs)


def test_table_overflow_is_flagged():
    import pde_engine_b200 as pb
    sess = pb.Session.for_problem("force_free")
    strs = [f"{k}*rho" for k in range(2, 200)]
    es = sess.compile(strs)
    fl = es.flags()
    assert fl[:120].sum() == 0 and (fl[130:] == 2).all()        # PDE_FLAG_TABLE_FULL
    osess = op.Session.for_problem("force_free")
    assert [op.compile_expr(s, osess).flags for s in strs] == list(fl)


def test_bytecode_does_not_depend_on_the_thread_count(enum_ff, monkeypatch):
    """The parse runs on up to 16 host threads; table slots are assigned afterwards in string order,
    so programs, flags and tables are identical for any thread count -- including the roll-back of
    expressions that hit a full constant table (they must not leave slots behind)."""
    import hashlib
    import pde_engine_b200 as pb
    E = uniques_by_depth(enum_ff)
    # > 128 distinct constants spread over the batch: the table fills in the middle of a worker's range
    strs = E[4][:30000] + [f"{k}*rho + z/{k + 1}" for k in range(2, 140)] + E[4][30000:42000] + ["", "rho +", "I*z"]
    ref = None
    for threads in ("1", "3", "16"):
        monkeypatch.setenv("PDE_B200_COMPILE_THREADS", threads)
        sess = pb.Session.for_problem("force_free")
        es = sess.compile(strs)
        code, ln = es.programs(128)
        fl = es.flags()
        got = (hashlib.sha256(code.tobytes() + ln.tobytes() + fl.tobytes()).hexdigest(), sess.const_keys(), sess.pow_keys())
        if ref is None:
            ref = got
            assert (fl == 2).sum() > 10 and sess.tables()[1] == 128       # the table did fill
            assert fl[-3:].tolist() == [1, 1, 1]
        assert got == ref, threads


def test_empty_batch_compiles():
    import pde_engine_b200 as pb
    sess = pb.Session.for_problem("force_free")
    es = sess.compile([])
    assert es.sizes() == (0, 0, 0)
    code, ln = es.programs(16)
    assert code.shape == (0, 16) and ln.shape == (0,)


def test_blob_entry_counts_the_strings_itself():
    """pde_compile_exprs_blob: a blob of known size must hold exactly n NUL-terminated strings -- fewer, more (a NUL
    inside a string), or a missing last terminator are errors, never an over-read; the bytecode equals the list entry's."""
    import pde_engine_b200 as pb
    from pde_engine_b200 import core
    sess = pb.Session.for_problem("force_free")
    strs = ["rho**2*z", "", "rho/z + 3", "sqrt(rho**2 + z**2) - z"]
    blob, n = core.pack_strings(strs)
    a, b = sess.compile(strs).export(), sess.compile_blob(blob, n).export()
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    for bad_blob, bad_n in ((blob, n + 1), (blob, n - 1), (blob[:-1], n), (b"rho\0\0z\0", 2), (b"", 1)):
        with pytest.raises(ValueError):
            sess.compile_blob(bad_blob, bad_n)
    assert sess.compile_blob(b"", 0).n == 0
    with pytest.raises(ValueError):
        sess.compile(["rho\0z"])


def test_warm_session_gives_the_same_bytecode(enum_ff, monkeypatch):
    """Keys the session tables already hold get their slots in the parallel parse phase, only new keys wait for the
    sequential numbering: a batch compiled in pieces (the later pieces into a warm session), in one piece, or again into
    the warm session gives the same programs and the same tables -- also when the table fills up on the way."""
    import pde_engine_b200 as pb
    E = uniques_by_depth(enum_ff)
    a = E[4][:20000] + [f"{k}*rho + z/{k + 1}" for k in range(2, 60)]
    b = E[4][20000:40000] + [f"{k}*rho + z/{k + 1}" for k in range(40, 140)] + E[3][:500]
    for threads in ("1", "5"):
        monkeypatch.setenv("PDE_B200_COMPILE_THREADS", threads)
        one = pb.Session.for_problem("force_free")
        whole = one.compile(a + b)
        cw, lw = whole.programs(128)
        fw = whole.flags()
        two = pb.Session.for_problem("force_free")
        ea, eb = two.compile(a), two.compile(b)
        ca, la = ea.programs(128)
        cb, lb = eb.programs(128)
        assert np.array_equal(np.concatenate([ca, cb]), cw) and np.array_equal(np.concatenate([la, lb]), lw)
        assert np.array_equal(np.concatenate([ea.flags(), eb.flags()]), fw)
        assert (fw == 2).sum() > 10                                     # the constant table did fill in `b`
        assert two.const_keys() == one.const_keys() and two.pow_keys() == one.pow_keys()
        again = two.compile(a + b)                                      # everything known (or known not to fit)
        c2, l2 = again.programs(128)
        assert np.array_equal(c2, cw) and np.array_equal(l2, lw) and np.array_equal(again.flags(), fw)
        assert two.const_keys() == one.const_keys()
