"""GPU parity tests (pytest -m gpu): the CUDA path, called through the C ABI, against the CPU oracle and the
golden fixtures of the real reference.  Integers (CSR arrays, negative samples, top-k ids) bit-exact;
fp32 within rtol 1e-5 (north_star)."""
import random

import numpy as np
import pytest
import torch

import laplace_gnn_recommendation_b200 as lg
from laplace_gnn_recommendation_b200.csr import DeviceCSR
from oracle import lightgcn_oracle as lo
from oracle import sampler_oracle as so
from oracle import topk_oracle as to

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def close(a, b, rtol=RTOL, atol=1e-7):
    torch.testing.assert_close(a.detach().cpu(), b.detach().cpu(), rtol=rtol, atol=atol)


def spmm_close(got, rowptr, col, val, X, post=lambda y: y, sequential=False):
    """SpMM parity.  Rows of ordinary length: rtol 1e-5 against the fp32 oracle.  Every row (incl. the heavy
    rows of skewed graphs, where a sequential fp32 sum of thousands of terms -- the oracle's own order -- is
    itself off by more than 1e-5): |got - fp64 value| <= 1e-5*|value| + 1e-6*sum_e|val_e*x_e|, i.e. rtol 1e-5
    plus a few fp32 ulps of the accumulated magnitude.  The fp32 oracle is sanity-checked against the rigorous
    sequential-summation bound deg*2^-24*sum|terms|."""
    got = got.detach().cpu()
    v64 = None if val is None else val.double()
    want32 = post(lo.spmm(rowptr, col, val, X))
    want64 = post(lo.spmm(rowptr, col, v64, X.double()))
    mag = post(lo.spmm(rowptr, col, None if val is None else v64.abs(), X.double().abs()))
    deg = (rowptr[1:] - rowptr[:-1]).double().unsqueeze(1)
    assert bool(((want32.double() - want64).abs() <= (deg + 2) * 2.0 ** -24 * mag + 1e-30).all()), "oracle sanity"
    err = (got.double() - want64).abs()
    bound = 1e-5 * want64.abs() + 1e-6 * mag + 1e-30
    if sequential:  # unsplit rows (chunk=0 / scalar path): one sequential fp32 chain per row, like the oracle itself
        bound = torch.maximum(bound, (deg + 2) * 2.0 ** -24 * mag)
    assert bool((err <= bound).all()), f"max excess {(err - bound).max().item():.3e}"
    short = (deg <= 256).squeeze(1)
    torch.testing.assert_close(got[short], want32[short], rtol=RTOL, atol=1e-6 * float(mag.max()))


def random_graph(seed, n_rows, n_cols, nnz, skew=False):
    g = torch.Generator().manual_seed(seed)
    if skew:  # a few very heavy rows/cols -> exercises the split plan
        row = (torch.rand(nnz, generator=g) ** 4 * n_rows).long().clamp(max=n_rows - 1)
        col = (torch.rand(nnz, generator=g) ** 3 * n_cols).long().clamp(max=n_cols - 1)
    else:
        row = torch.randint(0, n_rows, (nnz,), generator=g)
        col = torch.randint(0, n_cols, (nnz,), generator=g)
    return row, col


# ------------------------------------------------------------------ CSR / CSC / gcn_norm: bit-exact
@pytest.mark.parametrize("shape", [(9, 9, 7), (50, 70, 400), (1000, 1000, 20000), (5, 5, 0), (3, 4, 1)])
def test_csr_build_bit_exact(cuda_dev, shape):
    n_rows, n_cols, nnz = shape
    row, col = random_graph(1, n_rows, n_cols, nnz)
    g = DeviceCSR.from_coo(row.to(cuda_dev), col.to(cuda_dev), n_rows, n_cols, want_perm=True)
    rowptr, c, perm = lo.csr_from_coo(row, col, n_rows, n_cols)
    assert torch.equal(g.rowptr.cpu().long(), rowptr)
    assert torch.equal(g.colidx.cpu().long(), c)
    if nnz:
        keys = row * n_cols + col
        assert torch.equal(keys[g.perm.cpu()], lo.rows_from_rowptr(rowptr) * n_cols + c)
    t = g.transpose()
    colptr, r, csr2csc = lo.csc_from_csr(rowptr, c, n_cols)
    assert torch.equal(t.rowptr.cpu().long(), colptr)
    assert torch.equal(t.colidx.cpu().long(), r)
    assert torch.equal(t.csr2csc.cpu().long(), csr2csc)


def test_csr_reference_fixture_and_gcn_norm_bit_exact(cuda_dev):
    # tests/data_generator.py:144 fixture under wiring (R)
    row, col = torch.tensor([0, 0, 0, 1, 1, 2, 2]), torch.tensor([0, 2, 4, 1, 5, 3, 0])
    g = DeviceCSR.from_coo(row.to(cuda_dev), col.to(cuda_dev), 9, 9)
    assert g.rowptr.tolist() == [0, 3, 5, 7, 7, 7, 7, 7, 7, 7]
    assert g.colidx.tolist() == [0, 2, 4, 1, 5, 0, 3]
    dinv, val = g.gcn_norm()
    o_dinv, o_val = lo.gcn_norm_values(g.rowptr.cpu().long(), g.colidx.cpu().long())
    assert torch.equal(dinv.cpu(), o_dinv) and torch.equal(val.cpu(), o_val)
    row, col = random_graph(2, 3000, 3000, 50000, skew=True)
    g = DeviceCSR.from_coo(row.to(cuda_dev), col.to(cuda_dev), 3000, 3000)
    dinv, val = g.gcn_norm()
    o_dinv, o_val = lo.gcn_norm_values(g.rowptr.cpu().long(), g.colidx.cpu().long())
    assert torch.equal(dinv.cpu(), o_dinv) and torch.equal(val.cpu(), o_val)  # IEEE sqrt/div/mul: identical bits


# ------------------------------------------------------------------ SpMM
@pytest.mark.parametrize("d", [4, 8, 32, 64, 84, 128, 256, 6])
@pytest.mark.parametrize("chunk", [256, 8, 0])
def test_spmm_vs_oracle(cuda_dev, d, chunk, variant=None):
    seq = chunk == 0 or d % 4 != 0
    n, nnz = 700, 30000
    row, col = random_graph(d, n, n, nnz, skew=True)
    g = DeviceCSR.from_coo(row.to(cuda_dev), col.to(cuda_dev), n, n, chunk=chunk)
    if chunk == 8:
        assert g.n_long > 0 and g.n_tasks > g.n_long
    _, val = g.gcn_norm()
    g = g.with_values(val)
    X = torch.randn(n, d, generator=torch.Generator().manual_seed(0))
    rowptr, c = g.rowptr.cpu().long(), g.colidx.cpu().long()
    spmm_close(g.spmm(X.to(cuda_dev), variant=variant), rowptr, c, val.cpu(), X, sequential=seq)
    # unweighted sum and mean (hetero aggregation) on the same structure
    g1 = g.with_values(None)
    spmm_close(g1.spmm(X.to(cuda_dev), variant=variant), rowptr, c, None, X, sequential=seq)
    deg = (rowptr[1:] - rowptr[:-1]).clamp(min=1).float().unsqueeze(1)
    spmm_close(g1.spmm(X.to(cuda_dev), mean=True, variant=variant), rowptr, c, None, X, post=lambda y: y / deg.to(y.dtype),
               sequential=seq)
    # transposed operator (the backward)
    gt = g.transpose()
    colptr, r, csr2csc = lo.csc_from_csr(rowptr, c, n)
    spmm_close(gt.spmm(X.to(cuda_dev), variant=variant), colptr, r, val.cpu()[csr2csc], X, sequential=seq)


@pytest.mark.parametrize("variant", [0, 16, 20, 23])
def test_spmm_sweep_order_bit_identical(cuda_dev, variant):
    """Long-row slices executed in column-sweep order (lgb_csr.task_exec) give the same bits as plan order: partial sums are
    indexed by slice id, only the CTA -> slice assignment changes."""
    n, nnz, d = 400, 30000, 64
    row, col = random_graph(11, n, n, nnz, skew=True)
    g = DeviceCSR.from_coo(row.to(cuda_dev), col.to(cuda_dev), n, n, chunk=64)
    _, val = g.gcn_norm()
    g = g.with_values(val)
    assert g.n_tasks > 10
    X = torch.randn(n, d, generator=torch.Generator().manual_seed(2)).to(cuda_dev)
    y0 = g.spmm(X, variant=variant)
    g.use_sweep_order(True)
    ex = g.task_exec.long().cpu()
    assert torch.equal(torch.sort(ex).values, torch.arange(g.n_tasks))          # a permutation ...
    first = g.colidx.cpu()[g.task_start.cpu().long()][ex]
    assert bool((first[1:] >= first[:-1]).all()) and not torch.equal(ex, torch.arange(g.n_tasks))   # ... sorted by first column
    y1 = g.spmm(X, variant=variant)
    assert torch.equal(y0, y1)
    gt = g.transpose()                                                           # the transposed plan inherits the order
    assert gt.sweep and gt.task_exec is not None
    g.use_sweep_order(False)
    assert g.task_exec is None and torch.equal(g.spmm(X, variant=variant), y0)


@pytest.mark.parametrize("variant,d", [(0, 64), (16, 64), (20, 64), (23, 64), (0, 32), (16, 48), (12, 64)])
def test_spmm_fused_stage2_bit_identical(cuda_dev, variant, d, monkeypatch):
    """LGB_SPMM_FUSED_STAGE2: the stage-2 tree run inside the main launch (by whichever warp stores the last partial row of a
    segment / finishes the last segment of a row) gives the bits of the separate tree kernel, launch after launch (tickets
    reset themselves); kernel families without the fused path (variant 12) ignore the flag."""
    n, nnz = 200, 50000
    row, col = random_graph(31, n, n, nnz, skew=True)
    g = DeviceCSR.from_coo(row.to(cuda_dev), col.to(cuda_dev), n, n, chunk=8)        # row 0: ~13 000 entries -> > 32 segments
    _, val = g.gcn_norm()
    g = g.with_values(val).use_sweep_order(True)
    assert g.n_seg > 40 and int((g.row_seg0[1:] - g.row_seg0[:-1]).max()) > 32 and g.task_seg.numel() == g.n_tasks
    gen = torch.Generator().manual_seed(3)
    X, R, A = (torch.randn(n, d, generator=gen).to(cuda_dev) for _ in range(3))
    monkeypatch.setattr(lg.csr, "STAGE2_FUSED", False)
    y_tree = g.spmm(X, resid=R, variant=variant)
    acc_tree = torch.empty(n, d, device=cuda_dev)
    g.spmm(X, want_y=False, acc_in=A, acc_out=acc_tree, acc_div=4.0, variant=variant)
    monkeypatch.setattr(lg.csr, "STAGE2_FUSED", True)
    for _ in range(3):
        assert torch.equal(g.spmm(X, resid=R, variant=variant), y_tree)
    acc = torch.empty(n, d, device=cuda_dev)
    g.spmm(X, want_y=False, acc_in=A, acc_out=acc, acc_div=4.0, variant=variant)
    assert torch.equal(acc, acc_tree)
    ws = g._partial_ws(d)
    assert not bool(ws[g.n_tasks * d + 64 + g.n_seg * d:].any())                       # every ticket back at zero
    bm = lg.rows_bitmap(n, ((torch.arange(0, n, 7, device=cuda_dev), 0),), cuda_dev)
    Xs = torch.zeros(n, d, device=cuda_dev); Xs[::7] = X[::7]
    if d % 4 == 0 and d <= 64:
        monkeypatch.setattr(lg.csr, "STAGE2_FUSED", False)
        want = g.spmm(Xs, resid=Xs, x_rows=bm, resid_rows=bm)
        monkeypatch.setattr(lg.csr, "STAGE2_FUSED", True)
        assert torch.equal(g.spmm(Xs, resid=Xs, x_rows=bm, resid_rows=bm), want)


@pytest.mark.parametrize("d", [64, 32, 48, 20, 128, 6])
def test_spmm_rowsparse_matches_dense(cuda_dev, d):
    """lgb_spmm_rowsparse (operand zero outside the rows of a bitmap: entries that multiply a zero row are never gathered)
    against lgb_spmm on the same operand, and against the fp64 oracle -- long rows (slices + stage 2), short rows, fused
    residual / accumulate epilogue, every kernel shape (d = 64: 256-bit form; d4 <= 8; d4 <= 16; wider / scalar: dense fallback)."""
    n, nnz = 700, 40000
    row, col = random_graph(21, n, n, nnz, skew=True)
    g = DeviceCSR.from_coo(row.to(cuda_dev), col.to(cuda_dev), n, n, chunk=64)
    _, val = g.gcn_norm()
    g = g.with_values(val).use_sweep_order(True)
    gen = torch.Generator().manual_seed(d)
    for n_hot in (0, 1, 37, n):
        hot = torch.randperm(n, generator=gen)[:n_hot]
        X = torch.zeros(n, d)
        X[hot] = torch.randn(n_hot, d, generator=gen)
        R, A = torch.randn(n, d, generator=gen), torch.randn(n, d, generator=gen)
        # ids outside [0, n) and duplicates are ignored / harmless; the second list uses an offset
        lists = ((torch.cat([hot[: n_hot // 2], hot[: n_hot // 2], torch.tensor([-5, n + 3])]).to(cuda_dev), 0),
                 ((hot[n_hot // 2:] - 11).to(cuda_dev), 11))
        bm = lg.rows_bitmap(n, lists, cuda_dev)
        bits = torch.zeros(bm.numel() * 32, dtype=torch.bool)
        bits[hot] = True
        words = bm.cpu().to(torch.int64) & 0xFFFFFFFF
        got_bits = ((words[:, None] >> torch.arange(32)[None, :]) & 1).bool().view(-1)
        assert torch.equal(got_bits, bits)
        Xd, Rd, Ad = X.to(cuda_dev), R.to(cuda_dev), A.to(cuda_dev)
        dense = g.spmm(Xd, resid=Rd)
        sparse = g.spmm(Xd, resid=Rd, x_rows=bm)
        close(sparse, dense, rtol=1e-5, atol=1e-6)
        # the backward's call shape: resid is the row-sparse operand itself, and its bitmap says which rows to read
        close(g.spmm(Xd, resid=Xd, x_rows=bm, resid_rows=bm), g.spmm(Xd, resid=Xd), rtol=1e-5, atol=1e-6)
        base = lo.spmm(g.rowptr.cpu().long(), g.colidx.cpu().long(), val.cpu().double(), X.double()).float()
        close(sparse, base + R, rtol=1e-5, atol=1e-6)
        acc = torch.empty(n, d, device=cuda_dev)
        g.spmm(Xd, want_y=False, acc_in=Ad, acc_out=acc, acc_div=3.0, x_rows=bm)
        close(acc, (A + base) / 3.0, rtol=1e-5, atol=1e-6)
        assert torch.equal(g.spmm(Xd, x_rows=bm), g.spmm(Xd, x_rows=bm))          # deterministic
    with pytest.raises(RuntimeError):
        g.spmm(Xd, x_rows=bm[:2])                                                  # bitmap shorter than the operand


def test_spmm_rowsparse_random_shapes(cuda_dev):
    """Seeded sweep over small random shapes for lgb_spmm_rowsparse: bitmap sizes that are not multiples of 32, empty rows,
    duplicate entries, no plan / tiny slices (many segments), every hit density from none to all, both slice orders, resid
    with and without its own bitmap -- always the dense kernel's answer."""
    gen = torch.Generator().manual_seed(2024)
    for case in range(24):
        n_rows = int(torch.randint(1, 90, (1,), generator=gen))
        n_cols = int(torch.randint(1, 130, (1,), generator=gen))
        nnz = int(torch.randint(0, 1500, (1,), generator=gen))
        d = [4, 8, 20, 32, 48, 64][case % 6]
        chunk = [0, 4, 16, 64][(case // 6) % 4]
        row = (torch.rand(nnz, generator=gen) ** 3 * n_rows).long().clamp(max=n_rows - 1)
        col = torch.randint(0, n_cols, (nnz,), generator=gen)
        g = DeviceCSR.from_coo(row.to(cuda_dev), col.to(cuda_dev), n_rows, n_cols, chunk=chunk)
        g = g.with_values(torch.rand(nnz, generator=gen).to(cuda_dev) + 0.5)
        g.use_sweep_order(case % 2 == 0)
        density = [0.0, 0.05, 0.5, 1.0][case % 4]
        flagged = torch.rand(n_cols, generator=gen) < density
        X = torch.randn(n_cols, d, generator=gen) * flagged[:, None]
        ids = torch.nonzero(flagged).view(-1)
        bm = lg.rows_bitmap(n_cols, ((ids.to(cuda_dev), 0),), cuda_dev)
        r_flag = torch.rand(n_rows, generator=gen) < 0.3
        R = torch.randn(n_rows, d, generator=gen) * r_flag[:, None]
        rbm = lg.rows_bitmap(n_rows, ((torch.nonzero(r_flag).view(-1).to(cuda_dev), 0),), cuda_dev)
        Xd, Rd = X.to(cuda_dev), R.to(cuda_dev)
        what = f"case {case}: {n_rows}x{n_cols} nnz={nnz} d={d} chunk={chunk} density={density}"
        want = g.spmm(Xd, resid=Rd)
        torch.testing.assert_close(g.spmm(Xd, resid=Rd, x_rows=bm), want, rtol=1e-5, atol=1e-6, msg=what)
        torch.testing.assert_close(g.spmm(Xd, resid=Rd, x_rows=bm, resid_rows=rbm), want, rtol=1e-5, atol=1e-6, msg=what)
        torch.testing.assert_close(g.spmm(Xd, x_rows=bm, mean=True), g.spmm(Xd, mean=True), rtol=1e-5, atol=1e-6, msg=what)


def test_spmm_fused_epilogue_and_degree_order(cuda_dev):
    n, nnz, d = 500, 20000, 64
    row, col = random_graph(5, n, n, nnz, skew=True)
    g = DeviceCSR.from_coo(row.to(cuda_dev), col.to(cuda_dev), n, n, chunk=64)
    _, val = g.gcn_norm()
    g = g.with_values(val)
    gen = torch.Generator().manual_seed(1)
    X, R, A = (torch.randn(n, d, generator=gen) for _ in range(3))
    # float64 evaluation of the oracle: row 0 of this skewed graph has ~4000 entries (see spmm_close)
    base = lo.spmm(g.rowptr.cpu().long(), g.colidx.cpu().long(), val.cpu().double(), X.double()).float()
    for order in (False, True):
        g.use_degree_order(order)
        Y = torch.empty(n, d, device=cuda_dev); acc = torch.empty(n, d, device=cuda_dev)
        g.spmm(X.to(cuda_dev), Y=Y, resid=R.to(cuda_dev), acc_in=A.to(cuda_dev), acc_out=acc, acc_div=4.0)
        close(Y, base + R, atol=1e-6)
        close(acc, (A + (base + R)) / 4.0, atol=1e-6)
        acc2 = A.to(cuda_dev).clone()
        g.spmm(X.to(cuda_dev), want_y=False, acc_in=acc2, acc_out=acc2)       # in-place accumulate, no Y
        close(acc2, A + base, atol=1e-6)
    # determinism: two launches give identical bits (no atomics in the SpMM)
    y1, y2 = g.spmm(X.to(cuda_dev)), g.spmm(X.to(cuda_dev))
    assert torch.equal(y1, y2)


def test_csr_build_rejects_out_of_range_indices(cuda_dev):
    """torch_sparse's SparseTensor asserts row.max() < M / col.max() < N (what catches item ids that were not shifted by
    both_indexes_from_zero, data/lightgcn_loader.py:39-43): the build must raise, and must not write out of bounds."""
    row = torch.tensor([0, 1, 2, 9]); col = torch.tensor([3, 2, 1, 0])
    for bad_row, bad_col in ((torch.tensor([0, 1, 10, 9]), col), (row, torch.tensor([3, 12, 1, 0])), (torch.tensor([0, -1, 2, 9]), col)):
        with pytest.raises(RuntimeError, match="outside sparse_sizes"):
            DeviceCSR.from_coo(bad_row.to(cuda_dev), bad_col.to(cuda_dev), 10, 10)
    g = DeviceCSR.from_coo(torch.tensor([0, 1, 10, 9]).to(cuda_dev), col.to(cuda_dev), 10, 10, validate=False)   # deferred check
    assert int(g.rowptr[-1]) == 4                     # clamped on the device: the arrays stay well-formed
    with pytest.raises(RuntimeError, match="outside sparse_sizes"):
        g.check_built()
    DeviceCSR.from_coo(row.to(cuda_dev), col.to(cuda_dev), 10, 10)                                  # in range: fine
    with pytest.raises(RuntimeError):
        DeviceCSR.from_coo(row.to(cuda_dev), col.to(cuda_dev), 10, 0)                               # entries in an empty matrix


def test_spmm_empty_and_ragged(cuda_dev):
    d = 64
    g = DeviceCSR.from_coo(torch.empty(0, dtype=torch.long, device=cuda_dev), torch.empty(0, dtype=torch.long, device=cuda_dev), 10, 10)
    assert torch.count_nonzero(g.spmm(torch.randn(10, d, device=cuda_dev))) == 0
    # one row holds every entry, all others are empty
    n, nnz = 40, 5000
    row = torch.full((nnz,), 7); col = torch.randint(0, n, (nnz,), generator=torch.Generator().manual_seed(3))
    g = DeviceCSR.from_coo(row.to(cuda_dev), col.to(cuda_dev), n, n, chunk=128)
    X = torch.randn(n, d, generator=torch.Generator().manual_seed(4))
    spmm_close(g.spmm(X.to(cuda_dev)), g.rowptr.cpu().long(), g.colidx.cpu().long(), None, X)
    with pytest.raises(RuntimeError):
        g.spmm(torch.randn(n + 1, d, device=cuda_dev))
    with pytest.raises(RuntimeError):
        g.spmm(torch.randn(n, d))  # CPU tensor: no fallback


def test_spmm_large_properties(cuda_dev):
    """ML-1M-shaped symmetric graph: adjointness <A x, y> == <x, A^T y>, linearity, and a float64 check
    of sampled rows (size-independent properties; the oracle's message buffer would not fit at full scale)."""
    U, I, E, d = 6040, 3706, 1_000_209, 64
    gen = torch.Generator().manual_seed(1234)
    users, items = torch.randint(0, U, (E,), generator=gen), torch.randint(0, I, (E,), generator=gen)
    row, col, n = lo.wiring_symmetric(users, items, U, I)
    g = DeviceCSR.from_coo(row.to(cuda_dev), col.to(cuda_dev), n, n)
    assert g.nnz == 2 * E
    dinv, val = g.gcn_norm()
    g = g.with_values(val)
    x = torch.randn(n, d, generator=gen).to(cuda_dev); y = torch.randn(n, d, generator=gen).to(cuda_dev)
    Ax, Aty = g.spmm(x), g.transpose().spmm(y)
    lhs, rhs = (Ax.double() * y.double()).sum(), (x.double() * Aty.double()).sum()
    # both sides are sums of ~6e5 products of fp32-rounded values: allow 1e-6 of the accumulated magnitude
    assert abs(lhs - rhs) <= 1e-6 * float((Ax.double() * y.double()).abs().sum())
    close(g.spmm(2.0 * x + y), 2.0 * Ax + g.spmm(y), rtol=1e-4, atol=1e-5)
    close(Ax, g.transpose().spmm(x), rtol=1e-4, atol=1e-6)  # symmetric-normalised => A == A^T
    rows = torch.randint(0, n, (64,), generator=gen)
    rp, ci, v = g.rowptr.cpu().long(), g.colidx.cpu().long(), val.cpu().double()
    xc = x.cpu().double()
    for r in rows.tolist():
        s, e = rp[r], rp[r + 1]
        want = (v[s:e, None] * xc[ci[s:e]]).sum(0)
        torch.testing.assert_close(Ax[r].cpu().double(), want, rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------ LightGCN forward/backward + BPR
def run_case(c, dev, fused):
    fused = fused in (True, "rowsparse")
    U, I, d, K = c["U"], c["I"], c["d"], c["K"]
    model = lg.LightGCN(U, I, d, K)
    with torch.no_grad():
        model.users_emb.weight.copy_(c["Wu"]); model.items_emb.weight.copy_(c["Wi"])
    model = model.to(dev)
    adj = lg.SparseTensor(row=c["row"], col=c["col"], sparse_sizes=(U + I, U + I)).to(dev)
    u, p, n = c["u"].to(dev), c["p"].to(dev), c["n"].to(dev)
    if fused:
        loss = model.fused_step(adj, u, p, n, c["lam"])
        Ef = model.last_final
        u_f, i_f = Ef[:U], Ef[U:]
    else:
        u_f, u_0, i_f, i_0 = model(adj)
        assert u_0 is model.users_emb.weight and i_0 is model.items_emb.weight
        loss = lg.bpr_loss(u_f[u], u_0[u], i_f[p], i_0[p], i_f[n], i_0[n], c["lam"])
        loss.backward()
    return loss, u_f, i_f, model.users_emb.weight.grad, model.items_emb.weight.grad


@pytest.mark.parametrize("fused", [False, True, "rowsparse", "rowsparse-autograd"])
def test_lightgcn_against_reference_golden(cuda_dev, golden, fused, monkeypatch):
    if fused in ("rowsparse", "rowsparse-autograd"):      # first backward layer through lgb_spmm_rowsparse whatever the table / batch sizes
        monkeypatch.setattr(lg.lightgcn, "ROWSPARSE_MAX_SHARE", 0)
    for c in golden["lightgcn"]:
        loss, u_f, i_f, gu, gi = run_case(c, cuda_dev, fused)
        close(u_f, c["u_final"]); close(i_f, c["i_final"])
        close(loss, c["loss"])
        close(gu, c["dWu"], atol=1e-9); close(gi, c["dWi"], atol=1e-9)


@pytest.mark.parametrize("wiring", ["R", "S"])
@pytest.mark.parametrize("K,d", [(3, 64), (1, 32), (2, 128), (0, 64), (4, 32)])
def test_lightgcn_against_oracle(cuda_dev, wiring, K, d):
    U, I, E, B, lam = 400, 250, 6000, 512, 1e-4
    gen = torch.Generator().manual_seed(K * 100 + d)
    users, items = torch.randint(0, U, (E,), generator=gen), torch.randint(0, I, (E,), generator=gen)
    row, col, n = (lo.wiring_reference if wiring == "R" else lo.wiring_symmetric)(users, items, U, I)
    Wu, Wi = torch.randn(U, d, generator=gen) * 0.1, torch.randn(I, d, generator=gen) * 0.1
    pick = torch.randint(0, E, (B,), generator=gen)
    c = dict(U=U, I=I, d=d, K=K, row=row, col=col, Wu=Wu, Wi=Wi, lam=lam,
             u=users[pick], p=items[pick], n=torch.randint(0, I, (B,), generator=gen))
    rowptr, cc, _ = lo.csr_from_coo(row, col, n, n)
    o_loss, o_gu, o_gi, o_uf, o_if = lo.train_iteration(Wu, Wi, rowptr, cc, K, c["u"], c["p"], c["n"], lam)
    for fused in (False, True):
        loss, u_f, i_f, gu, gi = run_case(c, cuda_dev, fused)
        close(u_f, o_uf); close(i_f, o_if); close(loss, o_loss)
        close(gu, o_gu, atol=1e-9); close(gi, o_gi, atol=1e-9)


@pytest.mark.parametrize("wiring", ["R", "S"])
@pytest.mark.parametrize("d", [64, 32])
def test_lightgcn_small_batch_rowsparse_backward(cuda_dev, wiring, d, monkeypatch):
    """A batch that is small against the table (the reference's B = 128 on 1.48 M nodes, here 48 on 4 200): the first backward
    layer runs as lgb_spmm_rowsparse on its own accord -- from the batch indices in fused_step, from the non-zero rows
    lgb_scale_rows_nonzero finds in the autograd path -- and the gradients still equal the oracle's."""
    U, I, E, B, K, lam = 3000, 1200, 40000, 48, 3, 1e-4
    gen = torch.Generator().manual_seed(d)
    users = (torch.rand(E, generator=gen) ** 2 * U).long().clamp(max=U - 1)
    items = (torch.rand(E, generator=gen) ** 3 * I).long().clamp(max=I - 1)          # a few heavy items -> sliced rows
    row, col, n = (lo.wiring_reference if wiring == "R" else lo.wiring_symmetric)(users, items, U, I)
    Wu, Wi = torch.randn(U, d, generator=gen) * 0.1, torch.randn(I, d, generator=gen) * 0.1
    pick = torch.randint(0, E, (B,), generator=gen)
    c = dict(U=U, I=I, d=d, K=K, row=row, col=col, Wu=Wu, Wi=Wi, lam=lam,
             u=users[pick], p=items[pick], n=torch.randint(0, I, (B,), generator=gen))
    rowptr, cc, _ = lo.csr_from_coo(row, col, n, n)
    o_loss, o_gu, o_gi, o_uf, o_if = lo.train_iteration(Wu, Wi, rowptr, cc, K, c["u"], c["p"], c["n"], lam)
    calls = []
    real = DeviceCSR.spmm
    monkeypatch.setattr(DeviceCSR, "spmm", lambda self, *a, **k: (calls.append(k.get("x_rows") is not None), real(self, *a, **k))[1])
    for fused in (False, True):
        del calls[:]
        loss, u_f, i_f, gu, gi = run_case(c, cuda_dev, fused)
        assert calls == [False] * K + [True] + [False] * (K - 1)                       # exactly the first backward layer
        close(u_f, o_uf); close(i_f, o_if); close(loss, o_loss)
        close(gu, o_gu, atol=1e-9); close(gi, o_gi, atol=1e-9)
    # the autograd path found exactly the batch rows
    model = lg.LightGCN(U, I, d, K).to(cuda_dev)
    adj = lg.SparseTensor(row=row, col=col, sparse_sizes=(U + I, U + I)).to(cuda_dev)
    u, p, nn_ = c["u"].to(cuda_dev), c["p"].to(cuda_dev), c["n"].to(cuda_dev)
    u_f, u_0, i_f, i_0 = model(adj)
    lg.bpr_loss(u_f[u], u_0[u], i_f[p], i_0[p], i_f[nn_], i_0[nn_], lam).backward()
    want = torch.zeros(model._bitmap.numel() * 32, dtype=torch.bool)
    want[c["u"]] = True; want[c["p"] + U] = True; want[c["n"] + U] = True
    words = model._bitmap.cpu().to(torch.int64) & 0xFFFFFFFF
    got = ((words[:, None] >> torch.arange(32)[None, :]) & 1).bool().view(-1)
    assert torch.equal(got, want) and int(model._rows_count) == int(want.sum()) and model._rows_state["sparse"]


def test_bpr_against_reference_golden(cuda_dev, golden):
    for c in golden["bpr"]:
        xs = [x.clone().to(cuda_dev).requires_grad_(True) for x in c["inputs"]]
        loss = lg.bpr_loss(*xs, c["lam"])
        (loss * 1.0).backward()
        close(loss, c["loss"])
        for x, g in zip(xs, c["grads"]):
            close(x.grad, g, atol=1e-9)
    # upstream gradient is honoured
    xs = [x.clone().to(cuda_dev).requires_grad_(True) for x in golden["bpr"][0]["inputs"]]
    (lg.bpr_loss(*xs, 1e-6) * 3.0).backward()
    close(xs[0].grad, 3.0 * golden["bpr"][0]["grads"][0], atol=1e-9)


def test_bpr_full_split_size(cuda_dev):
    """evaluation() runs bpr_loss over every edge of a split (B = E_split): many CTAs + the finalize pass."""
    gen = torch.Generator().manual_seed(9)
    B, d = 200_003, 64
    xs = [torch.randn(B, d, generator=gen) * 0.3 for _ in range(6)]
    # At this size ATen's fp32 CPU norm() is itself ~1e-3 off (12.8M-term reduction), so the oracle is
    # evaluated in float64; the kernel's deterministic tree reduction must match THAT to rtol 1e-5.
    want = lo.bpr_loss(*[x.double() for x in xs], 1e-6).float()
    got = lg.bpr_loss(*[x.to(cuda_dev) for x in xs], 1e-6)
    close(got, want, rtol=1e-5)
    assert abs(lo.bpr_loss(*xs, 1e-6).item() - want.item()) < 2e-2 * abs(want.item())  # fp32 oracle: same ballpark


def test_module_contract(cuda_dev):
    """Attributes / state_dict / optimizer behaviour the reference driver relies on."""
    m = lg.LightGCN(7, 5, 16, 2).to(cuda_dev)
    assert [k for k in m.state_dict()] == ["users_emb.weight", "items_emb.weight"]
    assert len(list(m.parameters())) == 2
    assert m.users_emb.weight.shape == (7, 16) and m.items_emb.weight.shape == (5, 16)
    assert m._is_fused(m.users_emb.weight, m.items_emb.weight)
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    adj = lg.SparseTensor(row=torch.tensor([0, 1, 7, 8]), col=torch.tensor([7, 8, 0, 1]), sparse_sizes=(12, 12)).to(cuda_dev)
    before = m.items_emb.weight.detach().clone()
    for _ in range(2):
        u_f, u_0, i_f, i_0 = m(adj)
        idx = torch.tensor([0, 1], device=cuda_dev)
        loss = lg.bpr_loss(u_f[idx], u_0[idx], i_f[idx], i_0[idx], i_f[idx + 1], i_0[idx + 1], 1e-3)
        opt.zero_grad(); loss.backward(); opt.step()
    assert not torch.equal(before, m.items_emb.weight.detach())
    assert m._is_fused(m.users_emb.weight, m.items_emb.weight)  # in-place optimizer updates keep the fused table
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m2 = lg.LightGCN(7, 5, 16, 2).to(cuda_dev)
    m2.load_state_dict(sd)
    close(m2(adj)[0], m(adj)[0])


# ------------------------------------------------------------------ sampler: bit-exact indices
def test_sample_mini_batch_bit_exact_vs_reference_golden(cuda_dev, golden):
    g = golden["loader"]
    torch.manual_seed(g["seed"]); random.seed(g["seed"]); np.random.seed(g["seed"])
    tr = g["train"].to(cuda_dev)
    for want in g["batches"]:
        got = torch.stack(lg.sample_mini_batch(g["batch_size"], tr))
        assert got.device.type == "cpu" and got.dtype == torch.int64
        assert torch.equal(got, want)


@pytest.mark.parametrize("loops", [True, False])
def test_structured_negative_sampling_bit_exact(cuda_dev, loops):
    gen = torch.Generator().manual_seed(77)
    U, I, E = 300, 40, 6000   # dense enough that many draws are rejected and redrawn
    ei = torch.stack([torch.randint(0, U, (E,), generator=gen), torch.randint(0, I, (E,), generator=gen)])
    num_nodes = torch.max(ei[1])
    torch.manual_seed(5)
    want = so.structured_negative_sampling(ei, num_nodes=num_nodes, contains_neg_self_loops=loops)
    torch.manual_seed(5)
    got = lg.structured_negative_sampling(ei.to(cuda_dev), num_nodes=num_nodes, contains_neg_self_loops=loops)
    for a, b in zip(got, want):
        assert torch.equal(a.cpu(), b)
    torch.manual_seed(5)
    got_cpu_in = lg.structured_negative_sampling(ei, num_nodes=num_nodes, contains_neg_self_loops=loops)
    assert got_cpu_in[2].device.type == "cpu" and torch.equal(got_cpu_in[2], want[2])


# ------------------------------------------------------------------ top-k ids: bit-exact
def test_topk_against_reference_golden(cuda_dev, golden):
    g = golden["topk"]
    U, I = g["Wu"].shape[0], g["Wi"].shape[0]
    seen = lg.SeenItems(g["exclude"].to(cuda_dev), U, I)
    ids = lg.recommend_topk(g["Wu"].to(cuda_dev), g["Wi"].to(cuda_dev), torch.arange(U, device=cuda_dev), g["k"], seen)
    assert torch.equal(ids.cpu(), g["preds"])
    d = to.create_adj_dict(g["exclude"])
    for u in (0, 5, 39):
        one = lg.make_predictions_for_user(g["Wu"].to(cuda_dev), g["Wi"].to(cuda_dev), u, d, g["k"])
        assert torch.equal(one, g["preds"][u])


@pytest.mark.parametrize("k,I,d", [(12, 3706, 64), (256, 3706, 64), (12, 500, 32), (1000, 1200, 16), (5, 3, 8)])
def test_topk_against_oracle(cuda_dev, k, I, d):
    gen = torch.Generator().manual_seed(k + I)
    U = 64
    Wu, Wi = torch.randn(U, d, generator=gen), torch.randn(I, d, generator=gen)
    excl = torch.stack([torch.randint(0, U, (U * 20,), generator=gen), torch.randint(0, I, (U * 20,), generator=gen)])
    seen = lg.SeenItems(excl.to(cuda_dev), U, I)
    users = torch.arange(U, device=cuda_dev)
    ids, scores = lg.recommend_topk(Wu.to(cuda_dev), Wi.to(cuda_dev), users, k, seen, return_scores=True)
    ids = ids.cpu()
    seen_d = to.create_adj_dict(excl)
    full = Wu @ Wi.T
    for u in range(U):
        mask = torch.zeros(I, dtype=torch.bool)
        if u in seen_d:
            mask[seen_d[u]] = True
        n_avail = int((~mask).sum())
        got = ids[u][ids[u] >= 0]
        assert len(got) == min(k, n_avail)
        if k + len(seen_d.get(u, [])) <= I:
            want = to.predictions_for_user(Wu, Wi, u, seen_d, k)
            if not torch.equal(got, want):  # only legal difference: scores closer than fp32 summation-order noise
                diff = (got != want).nonzero().flatten()
                assert (full[u][got[diff]] - full[u][want[diff]]).abs().max() < 1e-5
        assert not mask[got].any()
        s = full[u][got]
        assert (s[:-1] >= s[1:] - 1e-5).all()


# ------------------------------------------------------------------ sharded engine (CUDA ops)
def test_sharded_engine_single_rank_matches_oracle(cuda_dev):
    """world=1: exercises dist.CudaOps (the two CSR views of the local block, lgb_gcn_values, lgb_mean_rows, the owned-row
    gather / scatter, lgb_bpr on explicit operand tables) on two streams."""
    from laplace_gnn_recommendation_b200.dist import ShardedLightGCN
    from tests.test_dist_gloo import make_problem, single_process_reference
    for K in (3, 1, 2):
        pb = make_problem(seed=K, U=500, I=120, E=9000, d=64, K=K, B=256)
        eng = ShardedLightGCN(pb["U"], pb["I"], pb["d"], K, pb["users"], pb["items"], cuda_dev,
                              init_tables=(pb["Wu"].to(cuda_dev), pb["Wi"].to(cuda_dev)), rank=0, world=1, max_batch=256,
                              schedule="chains" if K != 2 else "layer")
        loss = eng.fused_step(pb["u"], pb["p"], pb["n"], pb["lam"])
        if K == 3:                            # the same step replayed from a CUDA graph (three streams captured)
            gstep = eng.capture(256, pb["lam"])
            loss = gstep(pb["u"].to(cuda_dev), pb["p"].to(cuda_dev), pb["n"].to(cuda_dev)).clone()
        o_loss, o_gu, o_gi, o_uf, o_if = single_process_reference(pb)
        close(loss, o_loss)
        close(eng.E_f[: pb["U"]], o_uf); close(eng.E_f[pb["U"]:], o_if)
        close(eng.grad[: pb["U"]], o_gu, atol=1e-9); close(eng.grad[pb["U"]:], o_gi, atol=1e-9)


def test_sharded_engine_two_gpus_nccl(cuda_dev, tmp_path):
    """2 ranks over NCCL (skipped on a 1-GPU box): every rank's shard equals the single-process oracle."""
    import subprocess, sys, os
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = os.path.join(os.path.dirname(__file__), "dist_gpu_check.py")
    for mode in ("chains,auto", "chains,symm,graph", "layer,nccl"):
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                            "--master-addr", "127.0.0.1", "--master-port", "29617", script], capture_output=True, text=True,
                           timeout=300, env=dict(os.environ, DIST_CHECK_MODE=mode))
        assert r.returncode == 0, mode + r.stdout[-2000:] + r.stderr[-2000:]
        assert r.stdout.count("DIST_OK") == 2, mode


# ------------------------------------------------------------------ the headline configuration itself (BASELINE.json configs[2])
def test_hm_shape_epoch_against_fp64(cuda_dev):
    """The H&M-shaped graph bench.py times (U=1 371 980, I=105 542, E=31 788 324, power-law, d=64, K=3, B=128, same generator and
    seeds), ALL rows: final embeddings, loss and gradients of one fused epoch against a float64 evaluation on the host (scipy
    CSR), with the bound of spmm_close -- 1e-5*|y| + 1e-6*sum|terms| per element (a sequential fp32 sum over the heavy rows of
    this graph is itself further than 1e-5 from the exact value).  ~1.5 minutes of host time."""
    import scipy.sparse as sp
    import bench
    U, I, E = bench.WORKLOADS["hm"]
    d, K, B, lam = 64, 3, 128, 1e-6
    users, items = bench.make_graph(U, I, E, "powerlaw", 1234, cuda_dev)
    gen = torch.Generator(device=cuda_dev).manual_seed(42)
    pick = torch.randint(0, E, (B,), generator=gen, device=cuda_dev)
    ub, pb = users[pick].contiguous(), items[pick].contiguous()
    nb = torch.randint(0, I, (B,), generator=gen, device=cuda_dev)
    N = U + I
    adj = lg.SparseTensor(row=torch.cat([users, items + U]), col=torch.cat([items + U, users]), sparse_sizes=(N, N))
    torch.manual_seed(0)
    model = lg.LightGCN(U, I, d, K).to(cuda_dev)
    loss = model.fused_step(adj, ub, pb, nb, lam)
    Ef = model.last_final.cpu().double()
    G = torch.cat([model.users_emb.weight.grad, model.items_emb.weight.grad]).cpu().double()
    E0 = torch.cat([model.users_emb.weight, model.items_emb.weight]).detach().cpu().double().numpy()
    u, it = users.cpu().numpy(), items.cpu().numpy() + U
    del adj, model
    torch.cuda.empty_cache()
    import numpy as np
    rows, cols = np.concatenate([u, it]), np.concatenate([it, u])
    A = sp.csr_matrix((np.ones(rows.size, dtype=np.float64), (rows, cols)), shape=(N, N))      # duplicates add up, like repeated entries
    deg = np.asarray(A.sum(axis=1)).ravel()
    dinv = np.where(deg > 0, 1.0 / np.sqrt(np.maximum(deg, 1)), 0.0)
    A = sp.diags(dinv) @ A @ sp.diags(dinv)
    absA = abs(A)

    def propagate(x0, resid=None, m0=None):
        """-> (layer outputs [Horner form when resid is given], sum_k |A|^k m0: the magnitude every fp32 rounding of the
        K chained products is relative to; m0 = |x0| unless x0 itself is a difference of larger terms)"""
        m0 = np.abs(x0) if m0 is None else m0
        outs, x, m, mag = [x0], x0, m0, m0
        for _ in range(K):
            x = A @ x if resid is None else A @ x + resid
            m = absA @ m if resid is None else absA @ m + m0
            mag = mag + m
            outs.append(x)
        return outs, mag

    outs, mag = propagate(E0)
    want = sum(outs) / (K + 1)
    bound = 1e-5 * np.abs(want) + 1e-6 * mag
    err = np.abs(Ef.numpy() - want)
    assert (err <= bound).all(), f"E_f: max excess {(err - bound).max():.3e}"
    # loss and dE_f in float64 from the exact final embeddings, then the backward sum_k A^k r (A is symmetric)
    ui, pi, ni = ub.cpu().numpy(), pb.cpu().numpy() + U, nb.cpu().numpy() + U
    x = (want[ui] * (want[pi] - want[ni])).sum(1)
    reg = lam * ((E0[ui] ** 2).sum() + (E0[pi] ** 2).sum() + (E0[ni] ** 2).sum())
    want_loss = -np.mean(np.logaddexp(0.0, x)) + reg
    assert abs(float(loss) - want_loss) <= 1e-5 * abs(want_loss) + 1e-7
    sig = 1.0 / (1.0 + np.exp(-x))
    r = np.zeros_like(want)
    np.add.at(r, ui, (-sig / B)[:, None] * (want[pi] - want[ni]))
    np.add.at(r, pi, (-sig / B)[:, None] * want[ui])
    np.add.at(r, ni, (sig / B)[:, None] * want[ui])
    r /= (K + 1)
    rmag = np.zeros_like(want)                 # r is built from differences of fp32 embedding rows: its rounding is relative to these
    np.add.at(rmag, ui, (sig / B)[:, None] * (np.abs(want[pi]) + np.abs(want[ni])) / (K + 1))
    np.add.at(rmag, pi, (sig / B)[:, None] * np.abs(want[ui]) / (K + 1))
    np.add.at(rmag, ni, (sig / B)[:, None] * np.abs(want[ui]) / (K + 1))
    outs, mag = propagate(r, resid=r, m0=rmag)
    gwant = outs[-1]
    for idx in (ui, pi, ni):
        np.add.at(gwant, idx, 2 * lam * E0[idx])
    gerr = np.abs(G.numpy() - gwant)
    gbound = 1e-5 * np.abs(gwant) + 1e-6 * mag + 1e-6 * np.abs(gwant).max()      # + 1e-6 of the largest gradient entry
    assert (gerr <= gbound).all(), f"grad: max excess {(gerr - gbound).max():.3e}"
