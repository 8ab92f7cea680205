"""CPU: the oracle side of the function fingerprints (oracle/fingerprint.py) -- key function
known answers and properties, the exact partition, and the committed bucket fixtures."""
import hashlib

import numpy as np
import pytest

from conftest import load_golden, uniques_by_depth
from oracle import fingerprint as ofp


def test_key_function_properties():
    rng = np.random.default_rng(7)
    v = rng.standard_normal((50, 64)) * 10.0 ** rng.integers(-8, 8, (50, 64))
    key, nf = ofp.key_from_values(v, 26)
    assert key.dtype == np.uint64 and (nf == 64).all() and len(set(key.tolist())) == 50 and (key != 0).all()
    # round-off far below the kept mantissa does not change the key (values moved away from rounding boundaries first)
    drop = np.uint64(26)
    snapped = ((v.view(np.uint64) >> drop) << drop).view(np.float64)
    k0, _ = ofp.key_from_values(snapped, 26)
    k1, _ = ofp.key_from_values(snapped * (1 + 1e-13), 26)
    assert (k0 == k1).all()
    # a relative change of 2^-20 does
    k2, _ = ofp.key_from_values(snapped * (1 + 2.0 ** -20), 26)
    assert (k0 != k2).all()
    # the order of the points matters (salted), -0 == +0, every non-finite value hashes alike
    assert ofp.key_from_values(v[:, ::-1], 26)[0].tolist() != key.tolist()
    z = np.zeros((1, 64))
    assert ofp.key_from_values(z)[0] == ofp.key_from_values(-z)[0]
    a = v[:1].copy(); a[0, 3] = np.nan
    b = v[:1].copy(); b[0, 3] = -np.inf
    (ka, na), (kb, nb) = ofp.key_from_values(a), ofp.key_from_values(b)
    assert ka == kb and ka != key[0] and na == nb == 63
    # no finite value: key 0 = unknown
    k, n = ofp.key_from_values(np.full((2, 64), np.nan))
    assert k.tolist() == [0, 0] and n.tolist() == [0, 0]


def test_key_function_known_answers():
    """Pinned values of the integer key function (the device must reproduce them bit for bit,
    tests/test_gpu_fingerprint.py feeds it the same array)."""
    v = np.array([[1.0, -2.5, 0.0, 1e300, -1e-300, np.inf, np.nan, 3.0000000001] + [0.5] * 56,
                  [1.0, -2.5, 0.0, 1e300, -1e-300, np.inf, np.nan, 3.0000000002] + [0.5] * 56])
    key, nf = ofp.key_from_values(v, 26)
    assert nf.tolist() == [62, 62]
    assert key[0] == key[1]                       # 1e-10 apart at 26 bits: one bucket
    key40, _ = ofp.key_from_values(v, 40)
    assert key40[0] != key40[1]
    assert [hex(int(k)) for k in key] == KNOWN_KEYS_26
    assert [hex(_key_pure_python(row, 26)) for row in v] == KNOWN_KEYS_26


KNOWN_KEYS_26 = ["0xead97e1242a1e640", "0xead97e1242a1e640"]


def _key_pure_python(row, bits):
    """The definition again with Python integers (independent of the numpy restatement)."""
    import struct
    from oracle.bytecode import mix64, MASK64
    drop = 52 - bits
    h = 0
    nf = 0
    for k, x in enumerate(row):
        b = struct.unpack("<Q", struct.pack("<d", float(x)))[0]
        mag = b & 0x7FFFFFFFFFFFFFFF
        q = 0x7FF8000000000000
        if mag < 0x7FF0000000000000:
            nf += 1
            q = ((mag + (1 << (drop - 1))) >> drop) << drop
            if q:
                q |= b & 0x8000000000000000
        h = (h + mix64(q ^ ((0x9E3779B97F4A7C15 * (k + 1)) & MASK64))) & MASK64
    k64 = mix64(h)
    return (k64 or 1) if nf else 0


def test_exact_partition_semantics():
    from pde_engine_b200.problems import load_problem
    spec = load_problem("force_free", make_gpu=False)
    locs = spec.sympify_locals()
    syms = list(spec.symbols.values())
    pts = np.array([[0.7, 1.3, 0.4, 1.9, 0.9, 1.1], [-1.2, 0.6, -0.3, 1.7, 0.8, -1.5]])
    strs = ["rho", "neg(neg(rho))", "inv(inv(rho))", "z", "sqrt(z**2)", "sqrt(rho**2)", "rho*z", "z*rho + 0",
            "sqrt(-rho)", "exp(rho)/exp(rho)", "1"]
    b = ofp.exact_partition(strs, locs, syms, pts)
    assert b[:3] == [0, 0, 0]
    assert b[3] == 3 and b[4] == 4              # Abs(z) is not z: z is real, not positive (PI:70-71)
    assert b[5] == 0                            # rho is positive
    assert b[6] == b[7] == 6
    assert b[8] == -1                           # imaginary everywhere
    assert b[9] == b[10] == 9


@pytest.mark.parametrize("problem,depth,enum_name", [("force_free", 3, "enum_force_free_d4.json.gz"),
                                                     ("kerr_magnetosphere", 3, "enum_kerr_magnetosphere_d3.json.gz")])
def test_bucket_fixture_is_consistent(problem, depth, enum_name):
    fx = load_golden(f"function_buckets_{problem}_d{depth}.json.gz")
    E = uniques_by_depth(load_golden(enum_name))
    strs = [s for d in range(1, depth + 1) for s in E[d]]
    assert fx["n"] == len(strs) and fx["strings_sha256"] == hashlib.sha256("\n".join(strs).encode()).hexdigest()
    b = fx["bucket"]
    assert all(x == -1 or (0 <= x <= i and b[x] == x) for i, x in enumerate(b))
    assert fx["n_functions"] == len({x for x in b if x >= 0}) < fx["n"]
    # spot check: recompute a slice of the fixture with the oracle
    from pde_engine_b200.problems import load_problem
    import sympy as sp
    spec = load_problem(problem, make_gpu=False)
    extra = {}
    if problem != "force_free":
        extra = {spec.constants["M"]: sp.Rational(1.1378240173), spec.constants["a"]: sp.Rational(0.2718653942)}
    points = ofp.exact_points(list(spec.symbols.values()), np.array(fx["points"]), 6, extra)
    idx = list(range(0, len(strs), max(1, len(strs) // 60)))
    sig = {i: ofp.exact_signature(strs[i], spec.sympify_locals(), points) for i in idx}
    for i in idx:
        if b[i] == -1:
            assert sig[i] is None
        else:
            assert sig[i] == ofp.exact_signature(strs[b[i]], spec.sympify_locals(), points), strs[i]
