"""CPU tests: the oracle against (a) the golden fixtures produced by the REAL reference modules,
(b) the hand-computed known answers of SURVEY.md Appendix B, (c) independent implementations
(scipy.sparse, dense matmul, literal row loops).  Tolerance for fp32: rtol 1e-5 (north_star)."""
import random

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import hetero_oracle as ho
from oracle import lightgcn_oracle as lo
from oracle import sampler_oracle as so
from oracle import topk_oracle as to
from tests.conftest import FIXTURE_GRAPHS

RTOL, ATOL = 1e-5, 1e-7


def close(a, b, rtol=RTOL, atol=ATOL):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol)


# ---------------------------------------------------------------- Appendix B known answers
def test_appendix_b_csr_and_norm():
    U, I, e = FIXTURE_GRAPHS["random"]
    row, col = torch.tensor(e[0]), torch.tensor(e[1])
    rowptr, c, perm = lo.csr_from_coo(row, col, U + I, U + I)
    assert rowptr.tolist() == [0, 3, 5, 7, 7, 7, 7, 7, 7, 7]
    assert c.tolist() == [0, 2, 4, 1, 5, 0, 3]
    dinv, val = lo.gcn_norm_values(rowptr, c)
    assert dinv[:3].tolist() == pytest.approx([3 ** -0.5, 2 ** -0.5, 2 ** -0.5])
    assert (dinv[3:] == 0).all()
    expect = torch.tensor([1 / 3, 1 / 6 ** 0.5, 0, 1 / 2, 0, 1 / 6 ** 0.5, 0])
    close(val, expect, rtol=1e-6)


def test_appendix_b_forward_and_bpr():
    U, I, e = FIXTURE_GRAPHS["random"]
    rowptr, c, _ = lo.csr_from_coo(torch.tensor(e[0]), torch.tensor(e[1]), U + I, U + I)
    E0 = ((torch.arange(36).view(9, 4) % 7) - 3).float() * 0.1
    Wu, Wi = E0[:U].clone(), E0[U:].clone()
    u_f, _, i_f, _ = lo.lightgcn_forward(Wu, Wi, rowptr, c, 3)
    Ef = torch.cat([u_f, i_f])
    close(Ef[0], torch.tensor([-0.1648311, -0.1044063, -0.0439815, 0.0164433]), rtol=1e-5, atol=1e-7)
    close(Ef[1], torch.tensor([0.046875, 0.09375, 0.140625, -0.140625]), rtol=1e-5, atol=1e-7)
    close(Ef[2], torch.tensor([-0.1104411, -0.0634422, -0.0164433, 0.0305556]), rtol=1e-5, atol=1e-7)
    assert torch.equal(Ef[3:], E0[3:] / 4)  # item rows get no messages under wiring (R)
    loss, dWu, dWi, _, _ = lo.train_iteration(Wu, Wi, rowptr, c, 3, torch.tensor([0, 1, 2, 0]), torch.tensor([0, 1, 3, 4]),
                                              torch.tensor([5, 2, 2, 1]), 1e-6)
    assert loss.item() == pytest.approx(-0.69575399, abs=2e-7)
    dE0 = torch.cat([dWu, dWi])
    close(dE0[0], torch.tensor([0.0029376, -0.0006650, 0.0089576, -0.0006642]), rtol=1e-3, atol=1e-7)
    close(dE0[8], torch.tensor([-0.0051515, -0.0032627, -0.0013740, 0.0005133]), rtol=1e-3, atol=1e-7)


# ---------------------------------------------------------------- golden: real reference code
def test_lightgcn_golden(golden):
    for c in golden["lightgcn"]:
        n = c["U"] + c["I"]
        rowptr, col, _ = lo.csr_from_coo(c["row"], c["col"], n, n)
        loss, dWu, dWi, u_f, i_f = lo.train_iteration(c["Wu"], c["Wi"], rowptr, col, c["K"], c["u"], c["p"], c["n"], c["lam"])
        close(u_f, c["u_final"]); close(i_f, c["i_final"])
        close(loss, c["loss"])
        close(dWu, c["dWu"], atol=1e-9); close(dWi, c["dWi"], atol=1e-9)


def test_bpr_golden(golden):
    for c in golden["bpr"]:
        xs = [x.clone().requires_grad_(True) for x in c["inputs"]]
        loss = lo.bpr_loss(*xs, c["lam"])
        loss.backward()
        close(loss.detach(), c["loss"])
        for x, g in zip(xs, c["grads"]):
            close(x.grad, g, atol=1e-9)


def test_topk_and_metrics_golden(golden):
    g = golden["topk"]
    seen = to.create_adj_dict(g["exclude"])
    preds = torch.stack([to.predictions_for_user(g["Wu"], g["Wi"], u, seen, g["k"]) for u in range(g["Wu"].shape[0])])
    assert torch.equal(preds, g["preds"])
    recall, precision, ndcg, _ = to.metrics_lightgcn(g["Wu"], g["Wi"], g["eval"], [g["exclude"]], g["k"])
    assert (recall, precision, ndcg) == pytest.approx((g["recall"], g["precision"], g["ndcg"]), rel=1e-6)
    m = golden["metrics"]
    gt = [torch.arange(n) for n in m["gt_len"]]
    assert to.recall_precision_at_k(gt, m["r"], m["k"]) == pytest.approx(m["rp"], rel=1e-6)
    assert to.ndcg_at_k(gt, m["r"], m["k"]) == pytest.approx(m["ndcg"], rel=1e-6)
    dd = golden["difference_1d"]
    kept = np.setdiff1d(dd["a"].numpy(), dd["b"].numpy(), assume_unique=True)
    assert kept.tolist() == dd["out"].tolist() == [3, 1, 7]  # order preserved, not sorted


def test_loader_golden(golden):
    g = golden["loader"]
    assert torch.equal(so.both_indexes_from_zero(g["hom"]), g["edge_index"])
    tr, va, te, _ = so.split(g["edge_index"])
    assert torch.equal(tr, g["train"]) and torch.equal(va, g["val"]) and torch.equal(te, g["test"])
    torch.manual_seed(g["seed"]); random.seed(g["seed"]); np.random.seed(g["seed"])
    for want in g["batches"]:
        got = torch.stack(so.sample_mini_batch(g["batch_size"], g["train"]))
        assert torch.equal(got, want)


def test_negative_sampling_properties():
    torch.manual_seed(3)
    U, I, E = 20, 15, 120
    ei = torch.stack([torch.randint(0, U, (E,)), torch.randint(0, I, (E,))])
    num_nodes = int(ei[1].max())
    for loops in (True, False):
        u, p, n = so.structured_negative_sampling(ei, num_nodes=num_nodes, contains_neg_self_loops=loops)
        pos = set((ei[0] * num_nodes + ei[1]).tolist())
        assert all(k not in pos for k in (u * num_nodes + n).tolist())
        assert int(n.max()) < num_nodes  # the top item id is never drawn (reference quirk)
        if not loops:
            assert all(int(a) != int(b) for a, b in zip(u, n) if int(a) < num_nodes)


def test_decoder_and_hetero_golden(golden):
    d = golden["decoder"]
    close(ho.edge_decoder_mlp(d["z_user"], d["z_item"], d["eli"], d["linears"]), d["out"])
    ps = golden["padded_stack"]["out"]
    eli = torch.tensor([[0, 0, 0, 1, 1, 1, 1, 1, 2], [0] * 9])
    out = torch.cat([torch.arange(3.0), torch.arange(5.0), torch.arange(1.0)])
    assert torch.equal(ho.infer_rebatch(out, eli), ps)

    h = golden["hetero"]
    nu, ni, ek, rk = h["node_user"], h["node_item"], h["edge_key"], h["rev_edge_key"]
    for case in h["cases"]:
        sd = case["state_dict"]
        layers = []
        for li in range(2):
            layers.append({et: dict(w_l=sd[f"encoder.layers.{li}.{'__'.join(et)}.lin_l.weight"],
                                    b_l=sd[f"encoder.layers.{li}.{'__'.join(et)}.lin_l.bias"],
                                    w_r=sd[f"encoder.layers.{li}.{'__'.join(et)}.lin_r.weight"]) for et in (ek, rk)})
        z = ho.hetero_encoder(h["x"], h["edge_index"], layers, case["conv_aggr"], "sum", [ek, rk])
        zu = ho.batch_norm_train(z[nu], sd["encoder_layer_norm_customer.weight"], sd["encoder_layer_norm_customer.bias"])
        zi = ho.batch_norm_train(z[ni], sd["encoder_layer_norm_article.weight"], sd["encoder_layer_norm_article.bias"])
        lin = [(sd[f"decoder.layers.{i}.weight"], sd[f"decoder.layers.{i}.bias"]) for i in range(2)]
        logits = ho.edge_decoder_mlp(zu, zi, h["edge_label_index"], lin)
        close(logits, case["logits"], rtol=1e-4, atol=1e-6)
        close(ho.bce_with_logits(logits, case["labels"]), case["loss"], rtol=1e-5)


# ---------------------------------------------------------------- independent cross-checks
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_spmm_cross_checks(seed):
    torch.manual_seed(seed)
    n, nnz, d = 50, 400, 12
    row, col = torch.randint(0, n, (nnz,)), torch.randint(0, n, (nnz,))
    rowptr, c, perm = lo.csr_from_coo(row, col, n, n)
    _, val = lo.gcn_norm_values(rowptr, c)
    X = torch.randn(n, d)
    y = lo.spmm(rowptr, c, val, X)
    close(y, lo.spmm_rowloop(rowptr, c, val, X), atol=1e-6)
    A = sp.csr_matrix((val.numpy(), c.numpy(), rowptr.numpy()), shape=(n, n))
    close(y, torch.from_numpy(A @ X.numpy()), atol=1e-6)
    close(y, lo.spmm_mkl(rowptr, c, val, n) @ X, atol=1e-6)
    # csc_from_csr really is the transpose
    colptr, r, csr2csc = lo.csc_from_csr(rowptr, c, n)
    At = sp.csr_matrix((val[csr2csc].numpy(), r.numpy(), colptr.numpy()), shape=(n, n))
    assert abs(At - A.T).max() == 0
    # autograd backward of the oracle == A^T g
    Xg = X.clone().requires_grad_(True)
    g = torch.randn(n, d)
    lo.spmm(rowptr, c, val, Xg).backward(g)
    close(Xg.grad, torch.from_numpy(A.T @ g.numpy()), atol=1e-6)


def test_csr_keeps_duplicates_and_sorted_input_untouched():
    row = torch.tensor([2, 0, 2, 0, 1, 2]); col = torch.tensor([1, 3, 1, 0, 2, 0])
    rowptr, c, perm = lo.csr_from_coo(row, col, 4, 4)
    assert rowptr.tolist() == [0, 2, 3, 6, 6]
    assert c.tolist() == [0, 3, 2, 0, 1, 1]
    r2 = torch.tensor([0, 0, 1]); c2 = torch.tensor([1, 1, 0])
    _, _, p2 = lo.csr_from_coo(r2, c2, 2, 2)
    assert p2.tolist() == [0, 1, 2]
    rp0, c0, _ = lo.csr_from_coo(torch.empty(0, dtype=torch.long), torch.empty(0, dtype=torch.long), 3, 3)
    assert rp0.tolist() == [0, 0, 0, 0] and c0.numel() == 0


@pytest.mark.parametrize("aggr", ["add", "mean", "max"])
def test_scatter_aggregate_vs_loop(aggr):
    torch.manual_seed(4)
    ns, nd, E, F = 9, 7, 30, 5
    x = torch.randn(ns, F)
    ei = torch.stack([torch.randint(0, ns, (E,)), torch.randint(0, nd - 1, (E,))])  # last dst stays empty
    got = ho.scatter_aggregate(x, ei, nd, aggr)
    want = torch.zeros(nd, F)
    for t in range(nd):
        m = x[ei[0][ei[1] == t]]
        if len(m):
            want[t] = m.sum(0) if aggr == "add" else (m.mean(0) if aggr == "mean" else m.max(0).values)
    close(got, want, atol=1e-6)


def test_hetero_fan_in_order():
    a, b, c = torch.tensor([1.0]), torch.tensor([10.0]), torch.tensor([100.0])
    assert ho.hetero_fan_in([a, b, c], "sum").item() == 111.0
    assert ho.hetero_fan_in([a, b, c], "mean").item() == pytest.approx(37.0)
    assert ho.hetero_fan_in([a, b, c], "max").item() == 100.0
    assert ho.hetero_fan_in([a], "mean").item() == 1.0


def test_cpu_matmul_closure_matches_autograd_oracle():
    """The multi-threaded CPU baseline operator (bench.py) == the sequential oracle, forward and backward."""
    torch.manual_seed(8)
    n, nnz, d = 300, 4000, 16
    rowptr, c, _ = lo.csr_from_coo(torch.randint(0, n, (nnz,)), torch.randint(0, n, (nnz,)), n, n)
    _, val = lo.gcn_norm_values(rowptr, c)
    mm = lo.make_cpu_matmul(rowptr, c, val, n)
    Wu, Wi = torch.randn(200, d) * 0.1, torch.randn(100, d) * 0.1
    u, p, q = torch.randint(0, 200, (64,)), torch.randint(0, 100, (64,)), torch.randint(0, 100, (64,))
    a = lo.train_iteration(Wu, Wi, rowptr, c, 3, u, p, q, 1e-4, val=val)
    b = lo.train_iteration(Wu, Wi, rowptr, c, 3, u, p, q, 1e-4, val=val, matmul=mm)
    for x, y in zip(a, b):
        close(x, y, atol=1e-8)
