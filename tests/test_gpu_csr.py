"""CSR form of stage 1 (pde_enumerate_csr / pde_dedup_csr / pde_validate_csr): the same candidates, byte for byte, as
the padded-row entry points, which are pinned to the reference's candidate lists in test_gpu_enum.py."""
import numpy as np
import pytest

from conftest import uniques_by_depth

pytestmark = pytest.mark.gpu


def _sets(g, depth):
    E = uniques_by_depth(g)
    flat, db = [], [0]
    for d in range(1, depth):
        flat += E[d]
        db.append(len(flat))
    return flat, db


@pytest.mark.parametrize("problem,depth", [("force_free", 3), ("force_free", 4), ("kerr_magnetosphere", 3)])
def test_csr_equals_padded_rows(problem, depth, cuda_device, enum_ff, enum_kerr):
    import torch
    import pde_engine_b200 as pb
    flat, db = _sets(enum_ff if problem == "force_free" else enum_kerr, depth)
    sess = pb.Session.for_problem(problem)
    es = sess.compile(flat)
    n = pb.enumerate_count(es, db, depth, True)
    ref = pb.enumerate_candidates(es, db, depth, True, 0, n, 128)
    csr = pb.enumerate_candidates_csr(es, db, depth, True, 0, n, 128)
    torch.cuda.synchronize()
    rows, ln = pb.csr_rows(csr, 128)
    np.testing.assert_array_equal(ln, ref["len"].cpu().numpy())
    np.testing.assert_array_equal(rows, ref["code"].cpu().numpy())
    np.testing.assert_array_equal(csr["hash"].cpu().numpy(), ref["hash"].cpu().numpy())
    np.testing.assert_array_equal(csr["triple"].cpu().numpy(), ref["triple"].cpu().numpy())
    off = csr["off"].cpu().numpy().view(np.uint32).astype(np.int64)
    assert off[0] == 0 and np.all(np.diff(off) == (ln.astype(np.int64) + 15) // 16)          # packed, 16-byte aligned, no gaps
    assert off[-1] * 16 == csr["pool"].numel() or n == 0
    # the pool is a third of the padded rows
    assert csr["pool"].numel() < 0.45 * ref["code"].numel()
    # windows: any [first, first + count) reproduces its slice, offsets relative to its own pool
    rng = np.random.default_rng(5)
    for _ in range(6):
        first = int(rng.integers(0, n))
        count = int(rng.integers(0, min(n - first, 5000) + 1))
        w = pb.enumerate_candidates_csr(es, db, depth, True, first, count, 128)
        torch.cuda.synchronize()
        r2, l2 = pb.csr_rows(w, 128)
        np.testing.assert_array_equal(r2, rows[first:first + count])
        np.testing.assert_array_equal(w["hash"].cpu().numpy(), ref["hash"].cpu().numpy()[first:first + count])
        np.testing.assert_array_equal(w["triple"].cpu().numpy(), ref["triple"].cpu().numpy()[first:first + count])
    # dedup on CSR rows == dedup on padded rows
    f1, u1 = pb.dedup(ref["code"], ref["len"], ref["hash"])
    f2, u2 = pb.dedup_csr(csr["pool"], csr["off"], csr["len"], csr["hash"])
    assert u1 == u2
    np.testing.assert_array_equal(f1.cpu().numpy(), f2.cpu().numpy())


def test_validate_on_csr_rows(cuda_device, enum_ff):
    """pde_validate_csr == pde_validate on the same candidates: every output, bit for bit."""
    import torch
    import pde_engine_b200 as pb
    from pde_engine_b200.grids import collocation_grid
    flat, db = _sets(enum_ff, 3)
    sess = pb.Session.for_problem("force_free")
    prog = pb.ResidualProgram.for_problem("force_free")
    es = sess.compile(flat)
    n = pb.enumerate_count(es, db, 3, True)
    ref = pb.enumerate_candidates(es, db, 3, True, 0, n, 128)
    csr = pb.enumerate_candidates_csr(es, db, 3, True, 0, n, 128)
    pts = collocation_grid("force_free", 512)
    pts_t = torch.from_numpy(pts).to(cuda_device)
    tab_t = torch.from_numpy(prog.point_table(pts)).to(cuda_device)
    a = pb.validate(sess, prog, ref["code"], ref["len"], pts_t, tab_t, None, spill_slots=3)
    b = pb.validate(sess, prog, csr["pool"], csr["len"], pts_t, tab_t, None, spill_slots=3, row_off=csr["off"], L=128)
    torch.cuda.synchronize()
    for k in ("survivor_bits", "n_finite", "n_votes", "confirm"):
        np.testing.assert_array_equal(a[k].cpu().numpy(), b[k].cpu().numpy())
    for k in ("ratio_max", "resid_max", "scale_at", "ref_rs"):
        np.testing.assert_array_equal(a[k].cpu().numpy().view(np.int64), b[k].cpu().numpy().view(np.int64))


@pytest.mark.parametrize("problem", ["force_free", "kerr_magnetosphere"])
def test_filter_enumerated_equals_string_prefilter(problem, cuda_device, enum_ff, enum_kerr):
    """GpuBatchValidator.filter_enumerated (stage 1 -> stage 2 without leaving the device) gives every first-occurrence
    candidate the verdict `prefilter` gives the same candidate as a string; windows of the index space (what a rank of
    a sharded run evaluates) concatenate to the whole."""
    import torch
    import pde_engine_b200 as pb
    from pde_engine_b200.generator import candidate_string
    from pde_engine_b200.validator import GpuBatchValidator
    flat, db = _sets(enum_ff if problem == "force_free" else enum_kerr, 3)
    gv = GpuBatchValidator(None, problem=problem, P=512, L=128, spill_slots=3, group=None)
    surv = gv.filter_enumerated(flat, db, 3, True, 128)
    es = gv.session.compile(flat)
    n = pb.enumerate_count(es, db, 3, True)
    assert surv.shape == (n,)
    csr = pb.enumerate_candidates_csr(es, db, 3, True, 0, n, 128)
    first, _ = pb.dedup_csr(csr["pool"], csr["off"], csr["len"], csr["hash"])
    first = first.cpu().numpy().astype(bool)
    tri = csr["triple"].cpu().numpy()
    strs = [candidate_string(int(o), flat[a], flat[b] if b >= 0 else None) for o, a, b in tri]
    bv = gv.prefilter(strs)
    done = first & (csr["len"].cpu().numpy() != 0)   # (rows the device could not splice survive for the CPU path)
    np.testing.assert_array_equal(surv[done], bv.survivor[done])
    assert surv[~first].all()                        # exact duplicates are not evaluated: length 0 = "survives"
    assert 0 < int((~surv).sum()) < n
    # the caller's own enumeration + flags (the generator's case) and windows of it
    again = gv.filter_enumerated(flat, db, 3, True, 128, cand=csr, first_flags=torch.from_numpy(first.astype(np.uint8)).to(cuda_device),
                                 session=gv.session)
    np.testing.assert_array_equal(again, surv)
    words = []
    for lo, cnt in ((0, 2048), (2048, 1024), (3072, n - 3072)):
        words.append(gv._enum_filter_local(es, db, 3, True, 128, lo, cnt).cpu().numpy().view(np.uint32))
    w = np.concatenate(words)
    k = np.arange(n)
    np.testing.assert_array_equal(((w[k >> 5] >> (k & 31).astype(np.uint32)) & 1).astype(bool) | ~first, surv | ~first)


def test_csr_depth5_full_size(cuda_device, enum_ff):
    """BASELINE depth-5 enumeration (11 778 899 candidates) in CSR form: same hashes and lengths as the padded pass,
    pool = sum of the padded lengths, and the window of one rank of eight equals its slice."""
    import torch
    import pde_engine_b200 as pb
    flat, db = _sets(enum_ff, 5)
    sess = pb.Session.for_problem("force_free")
    es = sess.compile(flat)
    n = pb.enumerate_count(es, db, 5, True)
    assert n == 11778899
    ref = pb.enumerate_candidates(es, db, 5, True, 0, n, 128)
    h_ref, l_ref = ref["hash"].clone(), ref["len"].clone()
    del ref
    csr = pb.enumerate_candidates_csr(es, db, 5, True, 0, n, 128)
    torch.cuda.synchronize()
    assert torch.equal(csr["hash"], h_ref) and torch.equal(csr["len"], l_ref)
    padded = ((l_ref.to(torch.int64) + 15) // 16).sum().item() * 16
    assert csr["pool"].numel() == padded and int(csr["off"][-1].item()) * 16 == padded
    first, count = 3 * (n // 8), n // 8
    w = pb.enumerate_candidates_csr(es, db, 5, True, first, count, 128)
    torch.cuda.synchronize()
    assert torch.equal(w["hash"], h_ref[first:first + count])
    lo = int(csr["off"][first].item()) - int(w["off"][0].item())
    assert torch.equal((w["off"][:count].to(torch.int64) + lo), csr["off"][first:first + count].to(torch.int64))
    k = 100000
    a0, a1 = int(w["off"][0].item()) * 16, int(w["off"][k].item()) * 16
    b0 = int(csr["off"][first].item()) * 16
    assert torch.equal(w["pool"][a0:a1], csr["pool"][b0:b0 + (a1 - a0)])
    print(f"depth-5 CSR: {padded / n:.1f} pool bytes per candidate (+ 25 B offset/len/hash/triple) vs 149 B padded")
