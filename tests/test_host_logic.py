"""CPU tests of the product's device-agnostic host logic (no kernels involved): loader index paths against the
golden fixtures of the real reference, metric arithmetic against the real reference's numbers, the fused embedding
table behind LightGCN's two Parameters."""
import numpy as np
import pytest
import torch

import laplace_gnn_recommendation_b200 as lg


def test_loader_index_paths_match_reference_golden(golden):
    g = golden["loader"]
    assert torch.equal(lg.both_indexes_from_zero(g["hom"]), g["edge_index"])
    tr, va, te, full = lg.split(g["edge_index"])
    assert torch.equal(tr, g["train"]) and torch.equal(va, g["val"]) and torch.equal(te, g["test"])
    out = lg.make_lightgcn_splits(g["hom"], 30, 50)
    assert torch.equal(out[3], g["train"]) and out[7:] == (30, 50)
    assert out[0].sparse_sizes() == (80, 80) and out[0].nnz() == g["train"].shape[1]


def test_metric_arithmetic_matches_reference_golden(golden):
    """recall/precision/ndcg from the reference's own top-k lists == get_metrics_lightgcn's numbers."""
    g = golden["topk"]
    users = g["eval"][0].unique()
    ids = g["preds"][users]
    recall, precision, ndcg = lg.recall_precision_ndcg(ids, users, g["eval"], g["Wi"].shape[0], g["k"])
    assert (recall, precision, ndcg) == pytest.approx((g["recall"], g["precision"], g["ndcg"]), rel=1e-6)


def test_metric_arithmetic_edge_cases():
    users = torch.tensor([0, 2])
    ei = torch.tensor([[0, 0, 0, 2], [1, 1, 3, 0]])            # user 0 likes item 1 twice (duplicates count) and item 3
    ids = torch.tensor([[1, 5, 3], [4, 5, 6]])
    recall, precision, ndcg = lg.recall_precision_ndcg(ids, users, ei, 10, 3)
    assert recall == pytest.approx((2 / 3 + 0) / 2)
    assert precision == pytest.approx((2 + 0) / 2 / 3)
    d = 1.0 / np.log2(np.arange(2, 5))
    assert ndcg == pytest.approx(((d[0] + d[2]) / d.sum() + 0) / 2, rel=1e-6)


def test_lightgcn_parameters_share_one_table():
    torch.manual_seed(0)
    m = lg.LightGCN(5, 3, 8, 2)
    assert [k for k in m.state_dict()] == ["users_emb.weight", "items_emb.weight"]
    assert len(list(m.parameters())) == 2 and m._is_fused(m.users_emb.weight, m.items_emb.weight)
    assert m._table.shape == (8, 8)
    # same init stream as the reference: Embedding() draws N(0,1), then normal_(std=0.1) users first, items second
    torch.manual_seed(0)
    torch.nn.Embedding(5, 8); torch.nn.Embedding(3, 8)
    wu = torch.nn.init.normal_(torch.empty(5, 8), std=0.1); wi = torch.nn.init.normal_(torch.empty(3, 8), std=0.1)
    assert torch.equal(m.users_emb.weight.detach(), wu) and torch.equal(m.items_emb.weight.detach(), wi)
    # in-place updates (optimizers, load_state_dict) act on the shared table
    with torch.no_grad():
        m.items_emb.weight.add_(1.0)
    assert torch.equal(m._table[5:], wi + 1.0)
    m2 = lg.LightGCN(5, 3, 8, 2)
    m2.load_state_dict(m.state_dict())
    assert torch.equal(m2._table, m._table) and m2._is_fused(m2.users_emb.weight, m2.items_emb.weight)
    m3 = m.double().float()                                    # _apply re-fuses after dtype / device moves
    assert m3._is_fused(m3.users_emb.weight, m3.items_emb.weight)


def test_hetero_module_names_match_reference_checkpoint(golden):
    """state_dict keys of the (not yet materialised) ranking model == the real reference model's keys."""
    h = golden["hetero"]
    metadata = ([h["node_user"], h["node_item"]], [h["edge_key"], h["rev_edge_key"]])
    model = lg.Encoder_Decoder_Model(
        encoder_layers=lg.get_SAGEConv_layers(2, 16, 8, "add"), decoder_layers=lg.get_linear_layers(2, 16, 16, 1),
        feature_info={}, metadata=metadata, embedding=False, heterogeneous_prop_agg_type="sum", batch_normalize=True,
        p_dropout_edges=None, p_dropout_features=None)
    assert sorted(model.state_dict().keys()) == sorted(h["cases"][0]["state_dict"].keys())
    assert lg.get_SAGEConv_layers(3, 16, 8, "mean")[-1].out_channels == 8
    assert [l.out_features for l in lg.get_linear_layers(3, 16, 32, 1)] == [32, 32, 1]


def test_ranking_metrics_match_reference_golden():
    """get_metrics_universal (utils/metrics_encoder_decoder.py:29-86) -- oracle AND the vectorised mirror against values the
    real reference function produced (tests/golden/make_golden_ranking.py), incl. the infer() re-batching that feeds it."""
    import os
    import laplace_gnn_recommendation_b200 as lg
    from oracle import topk_oracle as to
    path = os.path.join(os.path.dirname(__file__), "golden", "reference_golden_ranking.pt")
    for c in torch.load(path, weights_only=False)["universal"]:
        want = (c["recall"], c["precision"], c["ndcg"])
        before = c["infer_out"].clone()
        got_o = to.metrics_universal(c["infer_out"], c["edge_index"], c["edge_label_index"], c["exclude"], c["k"])
        got = lg.get_metrics_universal(c["infer_out"], c["edge_index"], c["edge_label_index"], c["exclude"], c["k"])
        assert got_o == pytest.approx(want, rel=1e-6, abs=1e-7)
        assert got == pytest.approx(want, rel=1e-6, abs=1e-7)
        assert torch.equal(before, c["infer_out"])           # the caller's scores are left alone
        if c["infer_out"].dim() == 2:                        # the re-batching of model.infer (padded_stack of per-user scores)
            users = c["edge_label_index"][0].unique(sorted=True)
            rebatched = lg.hetero.padded_stack([c["scores"][c["edge_label_index"][0] == u] for u in users], value=-(1 << 50))
            assert torch.equal(rebatched, c["infer_out"])


def test_split_does_not_disturb_the_python_random_stream():
    """split() imports sklearn lazily; that import draws from Python's global `random` the first time it happens, which
    would shift the `random.choices` draws of sample_mini_batch depending on import order.  The reference imports sklearn at
    module load (before the caller seeds), so the stream must be the same before and after split()."""
    import random
    ei = torch.stack([torch.arange(50) % 7, torch.arange(50) % 11])
    random.seed(123)
    want = random.random()
    random.seed(123)
    lg.split(ei)
    assert random.random() == want
