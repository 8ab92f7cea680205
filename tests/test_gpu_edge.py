"""Edge cases of the device path through the C-ABI: empty and ragged batches, programs the kernel must
hand to the CPU (empty / malformed / out of spill space / complex-valued), maximum program length,
smallest and non-power-of-two grids, chunk tails."""
import numpy as np
import pytest

from oracle import bytecode as bc
from oracle import jets as J
from oracle import majorant as Mj
from oracle import parser as op
from oracle import residuals as Rz

pytestmark = pytest.mark.gpu


def _ctx(cuda_device, P, problem="force_free"):
    import torch
    import pde_engine_b200 as pb
    from pde_engine_b200.grids import collocation_grid
    sess = pb.Session.for_problem(problem)
    prog = pb.ResidualProgram.for_problem(problem)
    pts = collocation_grid(problem, P)
    return pb, sess, prog, pts, torch.from_numpy(pts).to(cuda_device), torch.from_numpy(prog.point_table(pts)).to(cuda_device)


def _rows(progs, L, dev):
    import torch
    code = np.zeros((len(progs), L), np.uint8)
    ln = np.zeros(len(progs), np.uint8)
    for i, p in enumerate(progs):
        code[i, :len(p)] = np.frombuffer(bytes(p), np.uint8)
        ln[i] = len(p)
    return torch.from_numpy(code).to(dev), torch.from_numpy(ln).to(dev)


def test_empty_batch_is_a_no_op(cuda_device):
    import torch
    pb, sess, prog, pts, pts_t, tab_t = _ctx(cuda_device, 64)
    n0 = pb.launch_count()
    out = pb.validate(sess, prog, torch.zeros((0, 48), dtype=torch.uint8, device=cuda_device),
                      torch.zeros(0, dtype=torch.uint8, device=cuda_device), pts_t, tab_t, None)
    assert out["n_finite"].numel() == 0 and out["survivor_bits"].numel() == 0 and pb.launch_count() == n0
    first, nu = pb.dedup(torch.zeros((0, 48), dtype=torch.uint8, device=cuda_device),
                         torch.zeros(0, dtype=torch.uint8, device=cuda_device), torch.zeros(0, dtype=torch.int64, device=cuda_device))
    assert nu == 0 and first.numel() == 0


@pytest.mark.parametrize("n", [1, 3, 4, 5, 37])
@pytest.mark.parametrize("P", [64, 192, 1024])
def test_ragged_batches_and_small_grids(cuda_device, n, P):
    """Chunks of 4 candidates per warp group and 32-point stripes dealt to 4 warps: every tail combination
    (n mod 4, P / 32 mod 4) must give the rows the oracle gives."""
    pb, sess, prog, pts, pts_t, tab_t = _ctx(cuda_device, P)
    strs = ["rho**2*z", "rho*z + exp(rho/z)", "sqrt(rho**2 + z**2) - z", "rho**2/z", "(rho + z)/(rho - z**2)"]
    strs = [strs[i % len(strs)] for i in range(n)]
    es = sess.compile(strs)
    code, ln = es.programs(48)
    import torch
    out = pb.validate(sess, prog, torch.from_numpy(code).to(cuda_device), torch.from_numpy(ln).to(cuda_device), pts_t, tab_t, None,
                      confirm_points=0, spill_slots=2)          # one pass: the majorant rule on every point
    nf, nv = out["n_finite"].cpu().numpy(), out["n_votes"].cpu().numpy()
    bits = out["survivor_bits"].cpu().numpy().view(np.uint32)
    osess = op.Session.for_problem("force_free")
    opts = np.ascontiguousarray(pts.T)
    for i, s in enumerate(strs):
        with np.errstate(all="ignore"):
            u, V, D, W = Mj.evaluate(op.compile_expr(s, osess).whole(), opts, 4, osess.const_vals, osess.pow_vals)
            R, S, _ = Rz.force_free_residual(u, opts[:, 0])
            St = Mj.force_free_scale(u, opts[:, 0], W, 1e-10)
            fin = np.isfinite(R) & np.isfinite(St) & (St > 0)
        slack = max(1, P // 100)      # points within float32 round-off of a pole radius / of the threshold
        assert abs(int(nf[i]) - int(fin.sum())) <= slack, (s, nf[i], fin.sum())
        votes = int((np.abs(R[fin]) > 1e-10 * St[fin]).sum())
        assert abs(int(nv[i]) - votes) <= slack, (s, nv[i], votes)
        reject = nf[i] >= 8 and nv[i] > 0 and nv[i] >= 0.5 * nf[i]
        assert ((bits[i >> 5] >> (i & 31)) & 1) == (0 if reject else 1), s
    # bits beyond n stay clear
    assert all(((bits[i >> 5] >> (i & 31)) & 1) == 0 for i in range(n, 32 * len(bits)))


def test_programs_the_kernel_hands_to_the_cpu(cuda_device):
    """n_finite < 0 codes (include/pde_b200.h): -1 empty, -2 malformed, -3 out of spill space; complex-valued
    candidates (NaN in real FP64) have no countable point.  All of them survive."""
    pb, sess, prog, pts, pts_t, tab_t = _ctx(cuda_device, 256)
    V0, V1, ADD, MUL, SQRT, NEG, EXP = bc.OP_VAR0, bc.OP_VAR1, bc.OP_ADD, bc.OP_MUL, bc.OP_SQRT, bc.OP_NEG, bc.OP_EXP
    sub = [V0, V1, MUL, EXP]                                   # a non-leaf sub-tree
    balanced = sub + sub + [MUL] + sub + sub + [MUL] + [ADD]   # (a*b) + (c*d) over sub-trees: needs 2 slots
    progs = [
        [],                                   # empty
        [V0, ADD],                            # malformed: binary op with one operand
        [V0, V1],                             # malformed: two values left
        [0x05],                               # malformed: unknown opcode
        balanced,                             # fine with 2 slots
        [V0, NEG, SQRT],                      # sqrt(-rho): complex everywhere
        [V0, V1, MUL],                        # plain
    ]
    code_t, len_t = _rows(progs, 48, cuda_device)
    out = pb.validate(sess, prog, code_t, len_t, pts_t, tab_t, None, spill_slots=2)
    nf = out["n_finite"].cpu().numpy()
    bits = int(out["survivor_bits"].cpu().numpy().view(np.uint32)[0])
    assert nf[0] == -1 and nf[1] == -2 and nf[2] == -2 and nf[3] == -2
    assert nf[4] == 256 and nf[5] == 0 and nf[6] == 256
    assert (bits & 0b0101111) == 0b0101111                     # everything not evaluated / not countable survives
    assert ((bits >> 6) & 1) == 0                              # rho*z is rejected
    # PRIM leaves without a table row behind them are malformed input, never dereferenced
    prim_prog = [[bc.OP_PRIM0 + 1, V0, MUL], [bc.OP_PRIM0 + 5]]
    cp, lp = _rows(prim_prog, 48, cuda_device)
    outp = pb.validate(sess, prog, cp, lp, pts_t, tab_t, None, spill_slots=2)
    assert outp["n_finite"].cpu().numpy().tolist() == [-2, -2]
    # a length beyond the row (caller error) is reported as malformed, never read
    len_bad = len_t.clone()
    len_bad[6] = 200
    outb = pb.validate(sess, prog, code_t, len_bad, pts_t, tab_t, None, spill_slots=2)
    assert outb["n_finite"].cpu().numpy()[6] == -2
    out1 = pb.validate(sess, prog, code_t, len_t, pts_t, tab_t, None, spill_slots=1)
    nf1 = out1["n_finite"].cpu().numpy()
    assert nf1[4] == -3 and nf1[6] == 256                      # one slot is not enough for (a*b) + (c*d) of sub-trees
    assert (int(out1["survivor_bits"].cpu().numpy().view(np.uint32)[0]) >> 4) & 1


def test_maximum_program_length(cuda_device):
    """L = 256 rows, a 255-byte program: rho followed by 254 negations is rho again."""
    pb, sess, prog, pts, pts_t, tab_t = _ctx(cuda_device, 128)
    long_prog = [bc.OP_VAR0, bc.OP_VAR1, bc.OP_MUL] + [bc.OP_NEG] * 252
    code_t, len_t = _rows([long_prog, [bc.OP_VAR0, bc.OP_VAR1, bc.OP_MUL]], 256, cuda_device)
    jets, resid, scale = pb.eval_points(sess, prog, code_t, len_t, pts_t, tab_t, None, spill_slots=1)
    j = jets.cpu().numpy()
    assert np.array_equal(j[0], j[1]) and np.isfinite(j[0]).all()
    r = resid.cpu().numpy()
    assert np.array_equal(r[0], r[1])


def test_evaluation_order_is_invisible(cuda_device):
    """The translator reorders operand evaluation (Sethi-Ullman); a - b and a / b must keep their operand
    roles whichever side is evaluated first."""
    pb, sess, prog, pts, pts_t, tab_t = _ctx(cuda_device, 128)
    heavy = "exp(rho*z)*sqrt(rho + z**2) + rho/(z**2 + 1)"       # needs a spill itself
    light = "exp(z)*rho"
    strs = [f"({heavy}) - ({light})", f"({light}) - ({heavy})", f"({heavy})/({light})", f"({light})/({heavy})",
            f"({light})*({heavy})", f"({heavy}) + ({light})"]
    es = sess.compile(strs)
    code, ln = es.programs(128)
    assert (ln > 0).all()
    import torch
    jets, resid, scale = pb.eval_points(sess, prog, torch.from_numpy(code).to(cuda_device), torch.from_numpy(ln).to(cuda_device),
                                        pts_t, tab_t, None, spill_slots=2)
    j = jets.cpu().numpy()
    osess = op.Session.for_problem("force_free")
    opts = np.ascontiguousarray(pts.T)
    for i, s in enumerate(strs):
        u = J.evaluate(op.compile_expr(s, osess).whole(), opts, 4, osess.const_vals, osess.pow_vals)
        ok = np.isfinite(u).all(axis=0)
        mag = np.max(np.abs(u[:, ok]), axis=0)
        assert np.all(np.max(np.abs(j[i][:, ok] - u[:, ok]), axis=0) <= 1e-9 * mag), s
