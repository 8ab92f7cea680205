"""Oracle enumeration / normaliser vs the reference's golden vectors (CPU)."""
import hashlib

import pytest

from conftest import load_golden, uniques_by_depth
from oracle import enumerate as oe
from oracle import normalizer as onorm


def test_candidates_match_reference_force_free(enum_ff):
    """Ordered candidate list per depth == what the unmodified reference handed to its
    normaliser (LBF:139-202): full lists at depth 2-3, count + sha256 at depth 4."""
    E = uniques_by_depth(enum_ff)
    for depth in (2, 3, 4):
        rec = enum_ff["depths"][str(depth)]
        cands = oe.candidates_for_depth(E, depth)
        assert len(cands) == rec["n_candidates"]
        assert hashlib.sha256("\n".join(cands).encode()).hexdigest() == rec["candidates_sha256"]
        if "candidates" in rec:
            assert cands == rec["candidates"]
    assert [enum_ff["depths"][str(d)]["n_candidates"] for d in (2, 3, 4)] == [128, 5924, 258285]


def test_candidates_match_reference_kerr(enum_kerr):
    E = uniques_by_depth(enum_kerr)
    for depth in (2, 3):
        rec = enum_kerr["depths"][str(depth)]
        cands = oe.candidates_for_depth(E, depth)
        assert cands == rec["candidates"]
    assert [enum_kerr["depths"][str(d)]["n_candidates"] for d in (2, 3)] == [372, 27936]


def test_triples_render_back_to_strings(enum_ff):
    E = uniques_by_depth(enum_ff)
    cands, triples = oe.candidates_for_depth(E, 3, with_triples=True)
    flat = E[1] + E[2]
    for s, (op, i, j) in zip(cands, triples):
        if op < 8:
            assert s == oe.unary_string(oe.UNARY_NAMES[op], flat[i])
        else:
            assert s == oe.binary_string(oe.BINARY_NAMES[op - 8], flat[i], flat[j])


@pytest.mark.parametrize("which", ["ff", "kerr"])
def test_stream_generate_matches_reference(which, enum_ff, enum_kerr):
    """Dedup + batching (LBF:137,198-215): uniques per depth and the on_batch chunk
    sizes equal the reference run; normalised strings come from the golden pairs
    (the reference's own normaliser output) to keep the test fast."""
    g = enum_ff if which == "ff" else enum_kerr
    memo = {}
    maxd = 3
    for d in range(2, maxd + 1):
        rec = g["depths"][str(d)]
        memo.update(dict(zip(rec["candidates"], rec["normalized"])))
    batches = {}
    E = oe.stream_generate(g["primitives"], lambda s: memo[s], maxd, batch_size=2000,
                           on_batch=lambda d, xs: batches.setdefault(d, []).append(list(xs)))
    for d in range(1, maxd + 1):
        rec = g["depths"][str(d)]
        assert E[d] == rec["uniques"]
        assert [len(b) for b in batches[d]] == rec["batch_sizes"]


def test_normalizer_matches_reference_cache():
    """oracle.normalizer.normalize == the reference's committed normaliser cache
    (lean_normalizer/physics_expressions.db, 113 rows) and the golden pairs."""
    fx = load_golden("ref_fixtures.json")
    for s, norm in fx["normalizer_cache"]:
        assert onorm.normalize(s) == norm, s


def test_normalizer_sample_depth3(enum_ff):
    rec = enum_ff["depths"]["3"]
    for s, norm in list(zip(rec["candidates"], rec["normalized"]))[::97]:
        assert onorm.normalize(s) == norm, s


def test_depth2_run_db_fixture(enum_ff):
    """The committed run DB (112 rows) pins enumeration + normalisation + DB dedup at
    depth 2: the expression column is E[1] ++ E[2] minus the rows dropped by the DB's
    UNIQUE(normalized) constraint (GM:1278-1286, 1407)."""
    fx = load_golden("ref_fixtures.json")
    rows = fx["ff_run_db"]
    E = uniques_by_depth(enum_ff)
    exprs = [r["expression"] for r in rows]
    stream = E[1] + E[2]
    # the DB keeps a subsequence of the stream, in order
    it = iter(stream)
    assert all(any(x == e for x in it) for e in exprs)
    assert len(E[2]) == 110 and len(rows) == 112
    sig = oe.signature
    assert sig("rho") == hashlib.sha256(b"rho").hexdigest()[:16]


def test_closed_form_candidate_count(enum_ff, enum_kerr):
    """oracle.enumerate.count_candidates (used where the list is too long to build) against the reference's
    own candidate counts at every recorded depth, and against the list builder with pruning off."""
    for g in (enum_ff, enum_kerr):
        E = uniques_by_depth(g)
        for d in sorted(E):
            if d < 2:
                continue
            assert oe.count_candidates(E, d) == g["depths"][str(d)]["n_candidates"]
        assert oe.count_candidates(E, 2, prune=False) == len(oe.candidates_for_depth(E, 2, prune=False))
        assert oe.count_candidates(E, 3, prune=False) == len(oe.candidates_for_depth(E, 3, prune=False))
