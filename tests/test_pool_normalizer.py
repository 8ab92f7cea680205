"""PoolNormalizer vs the reference's committed normaliser cache and the oracle restatement."""
import sqlite3
import time

from conftest import load_golden
from oracle.normalizer import OracleNormalizer
from pde_engine_b200.normalizer import PoolNormalizer


def test_matches_the_reference_cache_and_the_oracle(enum_ff, tmp_path):
    fx = load_golden("ref_fixtures.json")
    pairs = fx["normalizer_cache"]                     # candidate string -> normalised string, from the reference's own DB
    cands = enum_ff["depths"]["3"]["candidates"][:1500]
    db = str(tmp_path / "cache.db")
    with PoolNormalizer(n_workers=4, cache_db=db) as pn:
        got = pn.normalize_batch([(s, i) for i, (s, _) in enumerate(pairs)])
        assert [g["normalized"] for g in got] == [n for _, n in pairs]
        assert [g["index"] for g in got] == list(range(len(pairs)))
        t0 = time.time()
        res = pn.normalize_batch([(s, 3) for s in cands])
        dt = time.time() - t0
        assert pn.stats["misses"] >= 1000
    orc = OracleNormalizer().normalize_batch([(s, 3) for s in cands])
    assert res == orc                                      # normalised strings, indices and signatures
    # the cache is the reference's table, written in batches, and is reloaded by a new instance
    n = sqlite3.connect(db).execute("select count(*) from normalized_cache").fetchone()[0]
    assert n == len({s for s, _ in pairs} | set(cands))
    with PoolNormalizer(n_workers=1, cache_db=db) as pn2:
        again = pn2.normalize_batch([(s, 3) for s in cands[:50]])
        assert again == orc[:50] and pn2.stats["misses"] == 0
    print(f"1500 depth-3 candidates on 4 workers: {dt:.1f} s")


def test_unparsable_input_is_returned_unchanged():
    with PoolNormalizer(n_workers=1) as pn:
        assert pn.normalize("rho +* z") == "rho +* z"      # LB:78-79
        assert pn.normalize("(rho * rho**2 + z**2)") == "rho**3 + z**2"
        assert pn.normalize("neg(neg(rho))") == "neg(neg(rho))"   # opaque without locals
