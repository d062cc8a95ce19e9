"""GPU parity tests (pytest -m gpu) for the ranking encoder-decoder path: SAGE neighbour aggregation per edge
type, hetero fan-in, decoder (concat-MLP and dot), BatchNorm, BCE loss, gradients, infer -- against the golden
fixtures produced by the real reference model code and against the CPU oracle.  fp32 tolerance rtol 1e-5
(1e-4 after two Linear layers + BatchNorm, where cuBLAS and MKL GEMM summation orders differ)."""
import pytest
import torch
import torch.nn.functional as F

import laplace_gnn_recommendation_b200 as lg
from oracle import hetero_oracle as ho

pytestmark = pytest.mark.gpu


def close(a, b, rtol=1e-5, atol=1e-6):
    torch.testing.assert_close(a.detach().cpu(), b.detach().cpu(), rtol=rtol, atol=atol)


@pytest.mark.parametrize("aggr", ["add", "mean", "max"])
@pytest.mark.parametrize("F_in", [84, 76, 128, 64, 5])
def test_aggregate_fwd_bwd_vs_oracle(cuda_dev, aggr, F_in):
    gen = torch.Generator().manual_seed(F_in)
    ns, nd, E = 300, 200, 5000
    x = torch.randn(ns, F_in, generator=gen)
    ei = torch.stack([torch.randint(0, ns, (E,), generator=gen), torch.randint(0, nd - 3, (E,), generator=gen)])
    xo = x.clone().requires_grad_(True)
    want = ho.scatter_aggregate(xo, ei, nd, aggr)
    w = torch.randn(nd, F_in, generator=gen)
    (want * w).sum().backward()
    xg = x.to(cuda_dev).requires_grad_(True)
    g = lg.build_edge_csr(ei.to(cuda_dev), ns, nd)
    got = lg.aggregate(xg, g, aggr)
    (got * w.to(cuda_dev)).sum().backward()
    close(got, want, atol=1e-5)
    assert torch.count_nonzero(got[nd - 3:]) == 0          # empty destinations give exactly 0
    close(xg.grad, xo.grad, atol=1e-5)


def _load_model(case, h, dev, **kw):
    metadata = ([h["node_user"], h["node_item"]], [h["edge_key"], h["rev_edge_key"]])
    model = lg.Encoder_Decoder_Model(
        encoder_layers=lg.get_SAGEConv_layers(2, 16, 8, case["conv_aggr"]),
        decoder_layers=lg.get_linear_layers(2, 16, 16, 1),
        feature_info={}, metadata=metadata, embedding=False, heterogeneous_prop_agg_type="sum",
        batch_normalize=True, p_dropout_edges=None, p_dropout_features=None, **kw).to(dev)
    x = {k: v.to(dev) for k, v in h["x"].items()}
    ei = {k: v.to(dev) for k, v in h["edge_index"].items()}
    class _D:  # initialize_encoder_input_size only reads these two attributes
        x_dict, edge_index_dict = x, ei
    model.initialize_encoder_input_size(_D)
    return model, x, ei


def test_encoder_decoder_against_reference_golden(cuda_dev, golden):
    h = golden["hetero"]
    eli = h["edge_label_index"].to(cuda_dev)
    for case in h["cases"]:
        model, x, ei = _load_model(case, h, cuda_dev)
        assert sorted(model.state_dict().keys()) == sorted(case["state_dict"].keys())   # checkpoint-compatible names
        model.load_state_dict(case["state_dict"])
        model.train()
        logits = model(dict(x), ei, eli)
        loss = torch.nn.BCEWithLogitsLoss()(logits, case["labels"].to(cuda_dev))
        loss.backward()
        close(logits, case["logits"], rtol=1e-4, atol=1e-5)
        close(loss, case["loss"], rtol=1e-5)
        grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
        assert sorted(grads) == sorted(case["grads"])
        for k, gref in case["grads"].items():
            close(grads[k], gref, rtol=1e-3, atol=1e-5 * float(gref.abs().max()) + 1e-7)


def test_infer_rebatch_matches_oracle(cuda_dev, golden):
    h = golden["hetero"]
    case = h["cases"][0]
    model, x, ei = _load_model(case, h, cuda_dev)
    model.load_state_dict(case["state_dict"])
    eli = h["edge_label_index"].to(cuda_dev)
    out = model.infer(dict(x), ei, eli)
    model.eval()
    flat = model(dict(x), ei, eli).detach()
    want = ho.infer_rebatch(flat.cpu(), h["edge_label_index"])
    assert out.shape == want.shape
    close(out, want)
    assert (out.cpu() == -(1 << 50)).sum() == (want == -(1 << 50)).sum()


def test_decoder_concat_and_dot(cuda_dev, golden):
    d = golden["decoder"]
    zu, zi, eli = d["z_user"].to(cuda_dev), d["z_item"].to(cuda_dev), d["eli"].to(cuda_dev)
    layers = lg.get_linear_layers(2, 16, 16, 1).to(cuda_dev)
    with torch.no_grad():
        for l, (w, b) in zip(layers, d["linears"]):
            l.weight.copy_(w); l.bias.copy_(b)
    dec = lg.EdgeDecoder(layers, None)
    close(dec({"customer": zu, "article": zi}, eli), d["out"], rtol=1e-5, atol=1e-6)   # vs the REAL reference EdgeDecoder
    # gather+concat kernel is an exact copy; its backward equals index_add
    zu_g, zi_g = zu.clone().requires_grad_(True), zi.clone().requires_grad_(True)
    cat = lg.edge_concat(zu_g, zi_g, eli[0], eli[1])
    assert torch.equal(cat, torch.cat([zu[eli[0]], zi[eli[1]]], dim=-1))
    w = torch.randn_like(cat)
    (cat * w).sum().backward()
    zu_o, zi_o = d["z_user"].clone().requires_grad_(True), d["z_item"].clone().requires_grad_(True)
    (torch.cat([zu_o[d["eli"][0]], zi_o[d["eli"][1]]], dim=-1) * w.cpu()).sum().backward()
    close(zu_g.grad, zu_o.grad); close(zi_g.grad, zi_o.grad)
    # dot decoder vs its pure-torch statement, forward + backward, several widths
    for dim in (8, 64, 100, 7):
        gen = torch.Generator().manual_seed(dim)
        a, b = torch.randn(50, dim, generator=gen), torch.randn(70, dim, generator=gen)
        idx = torch.stack([torch.randint(0, 50, (500,), generator=gen), torch.randint(0, 70, (500,), generator=gen)])
        ao, bo = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
        want = ho.edge_decoder_dot(ao, bo, idx)
        gw = torch.randn(500, generator=gen)
        (want * gw).sum().backward()
        ag, bg = a.to(cuda_dev).requires_grad_(True), b.to(cuda_dev).requires_grad_(True)
        got = lg.edge_dot(ag, bg, idx[0].to(cuda_dev), idx[1].to(cuda_dev))
        (got * gw.to(cuda_dev)).sum().backward()
        close(got, want, atol=1e-5); close(ag.grad, ao.grad, atol=1e-5); close(bg.grad, bo.grad, atol=1e-5)
    dot_model = lg.EdgeDecoder(layers, None, mode="dot")
    close(dot_model({"customer": zu, "article": zi}, eli), ho.edge_decoder_dot(d["z_user"], d["z_item"], d["eli"]))


def test_decoder_node_projection_inference_form(cuda_dev, golden):
    """Dropout-off inference form of the two-layer decoder (per-node projections + lgb_edge_mlp2_fwd): equals the REAL reference
    EdgeDecoder output of the golden fixture and the per-edge concat form at the production widths; never used under autograd
    or with active dropout."""
    d = golden["decoder"]
    zu, zi, eli = d["z_user"].to(cuda_dev), d["z_item"].to(cuda_dev), d["eli"].to(cuda_dev)
    layers = lg.get_linear_layers(2, 16, 16, 1).to(cuda_dev)
    with torch.no_grad():
        for l, (w, b) in zip(layers, d["linears"]):
            l.weight.copy_(w); l.bias.copy_(b)
    z = {"customer": zu, "article": zi}
    from laplace_gnn_recommendation_b200 import _lib
    dec = lg.EdgeDecoder(layers, None, node_projection="on")
    n0 = _lib.LAUNCHES
    with torch.no_grad():
        got = dec(z, eli)
    assert _lib.LAUNCHES - n0 == 1                                   # one lgb_edge_mlp2_fwd, no concat launch
    close(got, d["out"], rtol=1e-5, atol=1e-6)
    # under autograd the reference form runs (the inference kernel has no backward)
    out = dec(z, eli)
    assert out.requires_grad
    close(out, d["out"], rtol=1e-5, atol=1e-6)
    # production widths (64 + 64 -> 128 -> 1), a non-multiple-of-4 hidden width, label edges outnumbering the nodes
    for H, du, di, nu, ni, L in ((128, 64, 64, 300, 200, 5000), (6, 8, 4, 30, 20, 400), (132, 64, 64, 50, 40, 1000)):
        gen = torch.Generator().manual_seed(H)
        a, b = torch.randn(nu, du, generator=gen).to(cuda_dev), torch.randn(ni, di, generator=gen).to(cuda_dev)
        idx = torch.stack([torch.randint(0, nu, (L,), generator=gen), torch.randint(0, ni, (L,), generator=gen)]).to(cuda_dev)
        torch.manual_seed(H)
        ls = lg.get_linear_layers(2, du + di, H, 1).to(cuda_dev)
        auto, off = lg.EdgeDecoder(ls, 0.5), lg.EdgeDecoder(ls, 0.5, node_projection="off")
        auto.eval(); off.eval()
        zz = {"customer": a, "article": b}
        with torch.no_grad():
            assert auto._use_node_projection(a, b, L)
            fast, ref = auto(zz, idx), off(zz, idx)
        close(fast, ref, rtol=1e-5, atol=2e-6)
        want = ho.edge_decoder_mlp(a.cpu(), b.cpu(), idx.cpu(), [(l.weight.detach().cpu(), l.bias.detach().cpu()) for l in ls])
        close(fast, want, rtol=1e-5, atol=2e-6)
        auto.train()                                                  # active dropout acts on the per-edge concatenation
        with torch.no_grad():
            assert not auto._use_node_projection(a, b, L)
        auto.eval()
        with torch.no_grad():                                         # few label edges, many nodes: the concat form is cheaper
            assert not auto._use_node_projection(a, b, 3)
    # Encoder_Decoder_Model.infer goes through it (golden infer test covers the values)


@pytest.mark.parametrize("hetero_aggr", ["sum", "mean", "max"])
def test_hetero_fan_in_three_edge_types(cuda_dev, hetero_aggr):
    """Two edge types into the same destination + one reverse type: exercises the pairwise fan-in."""
    gen = torch.Generator().manual_seed(3)
    nu, ni = 40, 30
    x = {"customer": torch.randn(nu, 12, generator=gen), "article": torch.randn(ni, 20, generator=gen)}
    ets = [("customer", "buys", "article"), ("customer", "views", "article"), ("article", "rev_buys", "customer")]
    eid = {ets[0]: torch.stack([torch.randint(0, nu, (200,), generator=gen), torch.randint(0, ni, (200,), generator=gen)]),
           ets[1]: torch.stack([torch.randint(0, nu, (150,), generator=gen), torch.randint(0, ni, (150,), generator=gen)]),
           ets[2]: torch.stack([torch.randint(0, ni, (180,), generator=gen), torch.randint(0, nu, (180,), generator=gen)])}
    enc = lg.to_hetero(lg.GNNEncoder(lg.get_SAGEConv_layers(2, 16, 8, "mean"), None, None),
                       (["customer", "article"], ets), aggr=hetero_aggr).to(cuda_dev)
    xd = {k: v.to(cuda_dev) for k, v in x.items()}
    z = enc(xd, {k: v.to(cuda_dev) for k, v in eid.items()})
    layers = []
    for li in range(2):
        layers.append({et: dict(w_l=enc.layers[li]["__".join(et)].lin_l.weight.detach().cpu(),
                                b_l=enc.layers[li]["__".join(et)].lin_l.bias.detach().cpu(),
                                w_r=enc.layers[li]["__".join(et)].lin_r.weight.detach().cpu()) for et in ets})
    want = ho.hetero_encoder(x, eid, layers, "mean", hetero_aggr, ets)
    for k in want:
        close(z[k], want[k], rtol=1e-4, atol=1e-5)


def test_batch_sized_aggregation_properties(cuda_dev):
    """Config-4 'M' sized batch (E_sub=4e5, widths 84 -> 128): adjointness of aggregate/backward and
    mean == sum / degree, at a size the CPU oracle would take too long for in a unit test."""
    gen = torch.Generator().manual_seed(11)
    ns, nd, E, Fw = 16_000, 90_000, 400_000, 84
    ei = torch.stack([torch.randint(0, ns, (E,), generator=gen), torch.randint(0, nd, (E,), generator=gen)]).to(cuda_dev)
    x = torch.randn(ns, Fw, generator=gen).to(cuda_dev).requires_grad_(True)
    y = torch.randn(nd, Fw, generator=gen).to(cuda_dev)
    g = lg.build_edge_csr(ei, ns, nd)
    s = lg.aggregate(x, g, "add")
    (s * y).sum().backward()
    lhs = (s.detach().double() * y.double()).sum()
    rhs = (x.detach().double() * x.grad.double()).sum()
    # <A x, y> == <x, A^T y>; both are sums of 7.6e6 products of fp32-rounded values (the inner product itself
    # cancels to O(1e2)), so the allowance is 1e-6 of the accumulated magnitude, not of the result
    assert abs(lhs - rhs) <= 1e-6 * float((s.detach().double() * y.double()).abs().sum())
    deg = torch.bincount(ei[1], minlength=nd).clamp(min=1).unsqueeze(1)
    close(lg.aggregate(x.detach(), g, "mean"), s.detach() / deg, atol=1e-6)
    ref = torch.zeros(nd, Fw, device=cuda_dev).index_add_(0, ei[1], x.detach()[ei[0]])
    close(s, ref, atol=1e-5)


# ------------------------------------------------------------------ split-K weight / bias gradient (csrc/dense.cu)
@pytest.mark.parametrize("N,n_in,n_out", [(5000, 84, 128), (777, 76, 64), (40000, 128, 1), (513, 6, 5), (1, 4, 4), (3000, 130, 70)])
def test_linear_wgrad_vs_torch(cuda_dev, N, n_in, n_out):
    """dW = dY^T X, db = column sums of dY against the float64 evaluation: |err| <= 1e-5*|want| + 1e-6*sum|terms| (the same
    bound as the SpMM: a sum over N rows in fp32), incl. widths that are not multiples of 4 / 64 and ragged row counts."""
    from laplace_gnn_recommendation_b200 import hetero
    gen = torch.Generator().manual_seed(N + n_in)
    X, dY = torch.randn(N, n_in, generator=gen), torch.randn(N, n_out, generator=gen)
    lin = torch.nn.Linear(n_in, n_out).to(cuda_dev)
    x = X.to(cuda_dev).requires_grad_(True)
    y = hetero._LinearFn.apply(x, lin.weight, lin.bias)
    torch.testing.assert_close(y.detach().cpu(), torch.nn.functional.linear(X, lin.weight.detach().cpu(), lin.bias.detach().cpu()), rtol=1e-5, atol=1e-5)
    y.backward(dY.to(cuda_dev))
    want_w, want_b = dY.double().t() @ X.double(), dY.double().sum(0)
    mag_w, mag_b = dY.double().abs().t() @ X.double().abs(), dY.double().abs().sum(0)
    for got, want, mag in ((lin.weight.grad, want_w, mag_w), (lin.bias.grad, want_b, mag_b)):
        err = (got.detach().cpu().double() - want).abs()
        assert bool((err <= 1e-5 * want.abs() + 1e-6 * mag + 1e-30).all()), float((err - 1e-5 * want.abs() - 1e-6 * mag).max())
    torch.testing.assert_close(x.grad.cpu(), dY @ lin.weight.detach().cpu(), rtol=1e-4, atol=1e-4)


def test_model_gradients_same_with_and_without_wgrad_kernel(cuda_dev, monkeypatch):
    """The ranking model on a batch large enough for the split-K path: loss and every parameter gradient agree with the plain
    nn.Linear backward (rtol 1e-5 of the gradient's magnitude)."""
    from laplace_gnn_recommendation_b200 import hetero
    gen = torch.Generator().manual_seed(2)
    Nc, Na, E, L = 700, 900, 6000, 800
    x = {"customer": torch.randn(Nc, 20, generator=gen).to(cuda_dev), "article": torch.randn(Na, 12, generator=gen).to(cuda_dev)}
    e = torch.stack([torch.randint(0, Nc, (E,), generator=gen), torch.randint(0, Na, (E,), generator=gen)]).to(cuda_dev)
    ei = {hetero.EDGE_KEY: e, hetero.REV_EDGE_KEY: e.flip(0).contiguous()}
    eli = torch.stack([torch.randint(0, Nc, (L,), generator=gen), torch.randint(0, Na, (L,), generator=gen)]).to(cuda_dev)
    y = (torch.rand(L, generator=gen) > 0.7).float().to(cuda_dev)
    metadata = (["customer", "article"], [hetero.EDGE_KEY, hetero.REV_EDGE_KEY])
    grads = {}
    for rows in (512, 10 ** 9):                       # fused backward on / off
        monkeypatch.setattr(hetero, "WGRAD_MIN_ROWS", rows)
        torch.manual_seed(0)
        model = lg.Encoder_Decoder_Model(encoder_layers=lg.get_SAGEConv_layers(2, 32, 16, "mean"), decoder_layers=lg.get_linear_layers(2, 32, 32, 1),
                                         feature_info={}, metadata=metadata, embedding=False, heterogeneous_prop_agg_type="sum",
                                         batch_normalize=True, p_dropout_edges=None, p_dropout_features=None).to(cuda_dev)
        model.initialize_encoder_input_size(type("B", (), {"x_dict": dict(x), "edge_index_dict": ei})())
        loss = torch.nn.BCEWithLogitsLoss()(model(dict(x), ei, eli), y)
        loss.backward()
        grads[rows] = (loss.detach(), {k: p.grad.clone() for k, p in model.named_parameters()})
    hetero.flush_deferred_checks()
    (la, ga), (lb, gb) = grads[512], grads[10 ** 9]
    torch.testing.assert_close(la, lb, rtol=1e-6, atol=1e-7)
    for k in ga:
        # (a bias in front of the batch norm has a zero gradient: both backward passes return rounding noise there, hence the floor)
        torch.testing.assert_close(ga[k], gb[k], rtol=1e-4, atol=1e-5 * float(gb[k].abs().max()) + 1e-7, msg=lambda m: f"{k}: {m}")


# ------------------------------------------------------------------ device-side sub-graph batch assembly
def _lex_sorted(e):
    order = torch.argsort(e[0] * (int(e[1].max()) + 1 if e.numel() else 1) + e[1], stable=True)
    return e[:, order]


def test_subgraph_sampler_against_reference_golden(cuda_dev):
    """SubgraphSampler.sample(roots) against GraphDataset.__getitem__ of the REAL reference (tests/golden/make_golden_subgraph.py:
    randomization=False, the reference's own test mode) on its manual test graphs and a random one, 1-3 hops: node features
    and label edges identical entry for entry, sub-graph edges identical as multisets (the reference walks Python sets), the
    batch = the disjoint union with PyG-style offsets."""
    import os
    golden = torch.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden_subgraph.pt"))
    for name, gcase in golden.items():
        c, cfg = gcase["inputs"], gcase["config"]
        s = lg.SubgraphSampler(c["edge_index"].to(cuda_dev), c["x_user"].to(cuda_dev), c["x_article"].to(cuda_dev),
                               n_hop_neighbors=cfg["n_hops"], num_neighbors=cfg["num_neighbors"], positive_edges_ratio=cfg["pos_ratio"],
                               negative_edges_ratio=cfg["neg_ratio"], k=cfg["k"], randomization=False)
        batch = s.sample(torch.tensor(c["roots"]).to(cuda_dev))
        uptr, aptr = batch["ptr"]["customer"].cpu(), batch["ptr"]["article"].cpu()
        ei, eli, lab = batch["edge_index_dict"][("customer", "buys", "article")].cpu(), batch["edge_label_index"].cpu(), batch["edge_label"].cpu()
        rev = batch["edge_index_dict"][("article", "rev_buys", "customer")].cpu()
        assert torch.equal(rev, ei.flip(0))
        xu, xa = batch["x_dict"]["customer"].cpu(), batch["x_dict"]["article"].cpu()
        for b, want in enumerate(gcase["items"]):
            what = f"{name} root #{b}"
            u0, u1, a0, a1 = int(uptr[b]), int(uptr[b + 1]), int(aptr[b]), int(aptr[b + 1])
            assert torch.equal(xu[u0:u1], want["x_user"]) and torch.equal(xa[a0:a1], want["x_article"]), what
            mine = (ei[0] >= u0) & (ei[0] < u1)
            assert bool(((ei[1][mine] >= a0) & (ei[1][mine] < a1)).all()), what            # block diagonal
            local = torch.stack([ei[0][mine] - u0, ei[1][mine] - a0])
            assert torch.equal(_lex_sorted(local), _lex_sorted(want["edge_index"])), what
            lm = (eli[0] >= u0) & (eli[0] < u1)
            assert torch.equal(torch.stack([eli[0][lm] - u0, eli[1][lm] - a0]), want["edge_label_index"]), what
            assert torch.equal(lab[lm], want["edge_label"]), what
        # the prebuilt CSR pair describes the same edges as the COO lists
        g = batch["graphs"][("customer", "buys", "article")]
        assert g.n_rows == int(aptr[-1]) and g.n_cols == int(uptr[-1]) and g.nnz == ei.shape[1]
        from laplace_gnn_recommendation_b200 import hetero
        hetero.flush_deferred_checks()


def test_subgraph_sampler_random_mode_and_model_step(cuda_dev):
    """Random mode (device generator): sampled positives are the root's own articles, the cheap-branch negatives lie in
    [0, id_max), fan-out cuts bound every frontier, every edge of the batch is a real edge -- and the batch (with its prebuilt
    CSR pair) drives a training step of the ranking model that equals the step on the same batch without the prebuilt graphs."""
    from laplace_gnn_recommendation_b200 import hetero
    gen = torch.Generator().manual_seed(8)
    U, A, E = 300, 200, 40000
    e = torch.stack([torch.randint(0, U, (E,), generator=gen), torch.randint(0, A, (E,), generator=gen)])
    e[0, :U] = torch.arange(U)
    pos = set(map(tuple, e.t().tolist()))
    xu, xa = torch.randn(U, 6, generator=gen), torch.randn(A, 5, generator=gen)
    s = lg.SubgraphSampler(e.to(cuda_dev), xu.to(cuda_dev), xa.to(cuda_dev), n_hop_neighbors=3, num_neighbors=7, randomization=True)
    roots = torch.randint(0, U, (16,), generator=gen)
    batch = s.sample(roots.to(cuda_dev))
    uid, aid = batch["n_id"]["customer"].cpu(), batch["n_id"]["article"].cpu()
    ei = batch["edge_index_dict"][hetero.EDGE_KEY].cpu()
    assert all((int(uid[a]), int(aid[b])) in pos for a, b in ei.t().tolist())            # remapped ids point at real edges
    eli, lab = batch["edge_label_index"].cpu(), batch["edge_label"].cpu()
    uptr = batch["ptr"]["customer"].cpu()
    for b, r in enumerate(roots.tolist()):
        m = (eli[0] >= uptr[b]) & (eli[0] < uptr[b + 1])
        assert bool((uid[eli[0][m]] == r).all())                                           # label edges start at the root
        assert all((r, int(aid[x])) in pos for x in eli[1][m & (lab == 1)].tolist())       # sampled positives are its articles
        n_pos_r, n_s = int((e[0] == r).sum()), int((m & (lab == 1)).sum())
        assert n_s == max(1, int(n_pos_r * 0.5)) and int((m & (lab == 0)).sum()) == int((11 if n_s <= 1 else 3.0) * n_s)
        # hop structure: 1 + <= 7 + <= 7 users (frontiers are cut to num_neighbors), besides the root's own row
        assert int(uptr[b + 1] - uptr[b]) <= 1 + 7 + 7
    metadata = (["customer", "article"], [hetero.EDGE_KEY, hetero.REV_EDGE_KEY])
    losses = []
    for use_graphs in (True, False):
        torch.manual_seed(0)
        model = lg.Encoder_Decoder_Model(encoder_layers=lg.get_SAGEConv_layers(2, 16, 8, "mean"), decoder_layers=lg.get_linear_layers(2, 16, 16, 1),
                                         feature_info={}, metadata=metadata, embedding=False, heterogeneous_prop_agg_type="sum",
                                         batch_normalize=True, p_dropout_edges=None, p_dropout_features=None).to(cuda_dev)
        out = model(dict(batch["x_dict"]), batch["edge_index_dict"], batch["edge_label_index"], graphs=batch["graphs"] if use_graphs else None)
        loss = torch.nn.BCEWithLogitsLoss()(out, batch["edge_label"].float())
        loss.backward()
        losses.append(float(loss))
    hetero.flush_deferred_checks()
    assert losses[0] == pytest.approx(losses[1], rel=1e-6)
