"""GPU: function fingerprints (pde_fingerprint) against the oracle:
integer key bit-exact, values against 30-digit SymPy, buckets against the exact partition."""
import ctypes

import numpy as np
import pytest

from conftest import load_golden, uniques_by_depth
from oracle import fingerprint as ofp

pytestmark = pytest.mark.gpu

CASES = [("force_free", 3, "enum_force_free_d4.json.gz"), ("kerr_magnetosphere", 3, "enum_kerr_magnetosphere_d3.json.gz")]


def _strings(enum_name, depth):
    E = uniques_by_depth(load_golden(enum_name))
    return [s for d in range(1, depth + 1) for s in E[d]]


@pytest.mark.parametrize("problem,depth,enum_name", CASES)
def test_keys_values_and_buckets(problem, depth, enum_name, cuda_device):
    import sympy as sp
    from pde_engine_b200.fingerprint import GENERIC_CONSTS, GpuFingerprinter
    from pde_engine_b200.problems import load_problem
    strs = _strings(enum_name, depth)
    fx = load_golden(f"function_buckets_{problem}_d{depth}.json.gz")
    assert fx["n"] == len(strs)
    fpr = GpuFingerprinter(problem, P=64, mantissa_bits=26, keep_values=True)
    f = fpr.fingerprint(strs)

    # (1) integer work: the key and the finite count are bit-exact functions of the device's own values
    key, nf = ofp.key_from_values(f.values, 26)
    assert (key == f.key).all() and (nf == f.n_finite).all()

    # (2) floating point: values against 30-digit SymPy on a sample; tolerance 1e-10 relative to max(|value|, 1e-3)
    # (the floor covers rows whose value is a small difference of O(1) terms)
    spec = load_problem(problem, make_gpu=False)
    locs = spec.sympify_locals()
    syms = list(spec.symbols.values())
    consts = {spec.constants[k]: sp.Rational(v) for k, v in GENERIC_CONSTS[fpr.problem].items()}
    checked = 0
    for i in range(0, len(strs), max(1, len(strs) // 150)):
        if f.key[i] == 0:
            continue
        u = sp.sympify(strs[i], locals=locs)
        for k in (0, 17, 40):
            sub = {syms[0]: sp.Rational(float(fpr.pts_host[0, k])), syms[1]: sp.Rational(float(fpr.pts_host[1, k]))}
            sub.update(consts)
            want = sp.N(u.subs(sub), 30)
            got = f.values[i, k]
            if want.is_real and want.is_finite:
                assert np.isfinite(got) and abs(got - float(want)) <= 1e-10 * max(abs(float(want)), 1e-3), (strs[i], k, got, want)
                checked += 1
            else:
                assert not np.isfinite(got), (strs[i], k, got, want)
    assert checked > 200

    # (3) buckets against the exact partition: NO false merge; few false splits (a value within round-off of a
    # rounding boundary splits a bucket -- bound 0.5 % of the rows); few unknown rows
    gold = np.array(fx["bucket"])
    both = (f.key != 0) & (gold >= 0)
    by_key = {}
    for i in np.nonzero(both)[0]:
        by_key.setdefault(int(f.key[i]), set()).add(int(gold[i]))
    merged = {k: v for k, v in by_key.items() if len(v) > 1}
    assert not merged, [[strs[j] for j in v] for v in list(merged.values())[:5]]
    n_device = len(by_key)
    n_exact = len({int(g) for g in gold[both]})
    assert n_exact <= n_device <= n_exact + max(2, int(0.005 * both.sum())), (n_exact, n_device)
    assert (f.key == 0).mean() < 0.02
    # rows the exact evaluation finds real somewhere are (almost all) fingerprinted
    assert ((f.key == 0) & (gold >= 0)).mean() < 0.01
    b = f.buckets()
    assert ((b == -1) == (f.key == 0)).all() and all(b[i] <= i for i in range(len(b)))


def test_same_function_different_strings(cuda_device):
    from pde_engine_b200.fingerprint import FunctionDedup, GpuFingerprinter
    fpr = GpuFingerprinter("force_free")
    strs = ["rho", "neg(neg(rho))", "inv(inv(rho))", "z", "sqrt(z**2)", "sqrt(rho**2)", "rho*z", "z*rho",
            "sqrt(-rho)", "exp(rho)/exp(rho)", "1", "rho**2*z", "square(rho)*z", "(rho + z)**2 - rho**2 - z**2", "2*rho*z"]
    f = fpr.fingerprint(strs)
    k = f.key
    assert k[0] == k[1] == k[2] == k[5] != 0
    assert k[3] != k[4] and k[3] != 0 and k[4] != 0        # Abs(z) vs z: the grid has both signs of z
    assert k[6] == k[7] and k[9] == k[10] and k[11] == k[12] and k[13] == k[14]
    assert k[8] == 0 and f.n_finite[8] == 0                  # imaginary everywhere: unknown, left to the CPU
    assert len({int(x) for x in k if x}) == 7
    g = f.groups()
    assert [strs[i] for i in g[0]["members"]] == ["rho", "neg(neg(rho))", "inv(inv(rho))", "sqrt(rho**2)"] and strs[g[0]["rep"]] == "rho"
    assert sum(len(x["members"]) for x in g) == len(strs) and len(g) == 8          # 7 functions + the unknown row
    dd = FunctionDedup(fpr)
    keep, same = dd.filter(strs)
    assert keep.tolist() == [True, False, False, True, True, False, True, False, True, True, False, True, False, True, False]
    assert same[1] == "rho" and same[14] == "(rho + z)**2 - rho**2 - z**2"
    keep2, same2 = dd.filter(["neg(neg(neg(neg(rho))))", "rho + 1"])
    assert keep2.tolist() == [False, True] and same2[0] == "rho"


def test_abi_edge_cases(cuda_device):
    import torch
    import pde_engine_b200 as pb
    from pde_engine_b200 import _lib, core
    sess = pb.Session.for_problem("force_free")
    pts = torch.rand((2, 64), dtype=torch.float64, device=cuda_device) + 0.5
    # empty batch: no-op
    rc = _lib.lib.pde_fingerprint(sess._h, None, None, 0, 48, ctypes.c_void_p(pts.data_ptr()), None, 0, 64, 4, 26, None, None, None, None)
    assert rc == 0
    es = sess.compile(["rho*z", "not an expression ((", "rho/(z - z)"])
    code, length = es.programs(48)
    code_t, len_t = torch.from_numpy(code).to(cuda_device), torch.from_numpy(length).to(cuda_device)
    values, key, nfin = core.fingerprint(sess, code_t, len_t, pts)
    assert key[0].item() != 0 and nfin[0].item() == 64
    assert key[1].item() == 0 and nfin[1].item() == 0 and torch.isnan(values[1]).all()      # not compilable
    assert key[2].item() == 0 and nfin[2].item() == 0                                       # x/0 everywhere
    with pytest.raises(_lib.PdeError):
        core.fingerprint(sess, code_t, len_t, pts, mantissa_bits=60)
    with pytest.raises(_lib.PdeError):
        core.fingerprint(sess, code_t, len_t, pts[:, :40].contiguous())


def test_run_with_shared_confirmations(cuda_device, tmp_path):
    """Opt-in: survivors that denote one function are confirmed once; every row is still stored and the verdicts
    equal those of the run that confirms each string on its own."""
    from oracle import symbolic as osym
    from oracle.normalizer import OracleNormalizer
    from pde_engine_b200.confirm import ConfirmationPool
    from pde_engine_b200.engine import run_discovery
    from pde_engine_b200.fingerprint import GpuFingerprinter
    from pde_engine_b200.problems import load_problem
    from pool_factories import oracle_force_free_factory
    runs = {}
    for shared in (False, True):
        spec = load_problem("force_free", cpu_validator=osym.ForceFreeSymbolicValidator())
        with ConfirmationPool(oracle_force_free_factory, n_workers=4, time_cap_s=300) as pool:
            runs[shared] = run_discovery(spec, OracleNormalizer(), 2, run_id="fp%d" % shared,
                                         db_normalize=osym.db_normalize, is_degenerate=osym.has_degenerate_denominator,
                                         confirm_pool=pool,
                                         share_confirmations=GpuFingerprinter("force_free") if shared else None)
    a, b = runs[False]["rows"], runs[True]["rows"]
    assert [r["expression"] for r in a] == [r["expression"] for r in b]
    assert [bool(r["is_valid"]) for r in a] == [bool(r["is_valid"]) for r in b]
    assert [r["paper_solution_name"] for r in a] == [r["paper_solution_name"] for r in b]
    sb = runs[True]["stats"]
    assert sb["confirmations_shared"] > 0
    assert sb["cpu_confirmed"] + sb["confirmations_shared"] == runs[False]["stats"]["cpu_confirmed"]
    assert any('"same_function_as"' in r["validator_evidence"] for r in b)


def test_kerr_parameters_are_generic(cuda_device):
    """The reference's check values are M = 1, a = 1/10 (PI:283): fingerprints use generic values so that
    expressions which only coincide there stay apart."""
    from pde_engine_b200.fingerprint import GpuFingerprinter
    f = GpuFingerprinter("kerr_magnetosphere").fingerprint(
        ["r", "M*r", "r*x", "x*r", "a**2", "1/100", "-2*M*r + a**2 + r**2", "a**2 + r**2 - 2*r", "sqrt(x**2)", "x"])
    k = f.key
    assert (k != 0).all()
    assert k[0] != k[1] and k[2] == k[3] and k[4] != k[5] and k[6] != k[7] and k[8] != k[9]
