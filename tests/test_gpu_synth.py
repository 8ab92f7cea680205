"""Synthetic depth-5 workload: generator parity, PRIM leaves, flop table."""
import numpy as np
import pytest

from oracle import bytecode as bc
from oracle import jets as J
from oracle import parser as op
from oracle import residuals as Rz
from oracle import synth as osyn

pytestmark = pytest.mark.gpu


def test_synth_trees_bit_exact(cuda_device):
    import torch
    import pde_engine_b200 as pb
    n = 4096
    for depth in (1, 2, 5):
        dev = pb.synth_trees(osyn.SEED_TREES, 100, n, depth, 48)
        torch.cuda.synchronize()
        code, ln, hs = dev["code"].cpu().numpy(), dev["len"].cpu().numpy(), dev["hash"].cpu().numpy()
        want = osyn.trees(osyn.SEED_TREES, 100, n, depth)
        for i, w in enumerate(want):
            assert ln[i] == len(w) and bytes(code[i, :len(w)]) == w and not code[i, len(w):].any()
            assert int(hs[i]) & bc.MASK64 == bc.structural_hash(w)
            assert bc.stack_depth(w) >= 1
    assert max(len(w) for w in want) <= 17


def test_flop_table_matches_oracle(cuda_device):
    import torch
    import pde_engine_b200 as pb
    from pde_engine_b200.flops import batch_flops_per_point, op_cost_table
    for order in (4, 2):
        t = osyn.flop_table(order)
        mine = op_cost_table(order)
        for b, c in t.items():
            assert mine[b] == c, hex(b)
    dev = pb.synth_trees(osyn.SEED_TREES, 0, 2000, 5, 48)
    want = osyn.trees(osyn.SEED_TREES, 0, 2000, 5)
    tab = osyn.flop_table(4)
    total = sum(osyn.program_flops(w, tab) for w in want) + 2000 * osyn.RESIDUAL_FLOPS["force_free"]
    assert batch_flops_per_point(dev["code"], "force_free", 4) == total
    # SURVEY 8d expectation: 0.65-0.75 kflop per (candidate, point)
    assert 500 < total / 2000 < 900


def test_synth_validate_matches_oracle(cuda_device):
    """BASELINE config 5's trees (PRIM leaves): the primitive jet table is produced by the interpreter itself
    (pde_eval_points on the primitives' own programs, majorant pairs included) and consumed by PRIM(p).  Same
    assertions as the real-candidate parity tests (test_gpu_validate._compare_points): every finite point, no
    statistical allowance -- ill-conditioned points (poles of 1/(1-b), exp of large arguments) are covered by the
    round-off majorant the device carries, not by an exemption."""
    import torch
    import pde_engine_b200 as pb
    import test_gpu_validate as tv
    from pde_engine_b200.grids import collocation_grid
    from pde_engine_b200.synthetic import primitive_jets, unpack_primitive_table
    P, n = 64, 1500
    sess = pb.Session.for_problem("force_free")
    prog = pb.ResidualProgram.for_problem("force_free")
    pts = collocation_grid("force_free", P)
    pts_t = torch.from_numpy(pts).to(cuda_device)
    tab_t = torch.from_numpy(prog.point_table(pts)).to(cuda_device)
    prim_t = primitive_jets(sess, prog, pts_t, tab_t)
    assert tuple(prim_t.shape) == (2, P // 32, 16, 32)           # packed stripe-block device table
    osess = op.Session.for_problem("force_free")
    opts = np.ascontiguousarray(pts.T)
    from oracle import majorant as Mj
    pm = [Mj.evaluate(op.compile_expr(s, osess).whole(), opts, 4, osess.const_vals, osess.pow_vals) for s in osyn.PRIM_EXPRS]
    prim_np = unpack_primitive_table(prim_t, 15).cpu().numpy()
    np.testing.assert_allclose(prim_np[0], pm[0][0], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(prim_np[1], pm[1][0], rtol=1e-13, atol=1e-15)
    # row 15 of the table: the leaf's (D, W) as two float32
    row15 = prim_t[:, :, 15, :].contiguous().view(torch.float32).reshape(2, P // 32, 32, 2).cpu().numpy()
    for k in range(2):
        np.testing.assert_allclose(row15[k, :, :, 0].reshape(P), pm[k][2], rtol=2e-3)
        np.testing.assert_allclose(row15[k, :, :, 1].reshape(P), pm[k][3], rtol=2e-3)
    dev = pb.synth_trees(osyn.SEED_TREES, 0, n, 5, 48)
    out, _ = tv._device_eval(pb, sess, prog, None, pts_t, tab_t, cuda_device, prim=prim_t, code=(dev["code"], dev["len"]))
    progs = osyn.trees(osyn.SEED_TREES, 0, n, 5)
    oracle = tv._oracle_eval("force_free", progs, pts, prims=[m[0] for m in pm], prim_maj=[(m[2], m[3]) for m in pm])
    n_cmp = tv._compare_points(out, oracle, [str(i) for i in range(n)], 4, fin_agree=0.98)
    assert n_cmp > 0.5 * n * P
