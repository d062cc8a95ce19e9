"""The reference's OWN drivers, unmodified, on this package (north_star: "run_pipeline_lightgcn.py and
run_pipeline.py drop it in").  Runs in the build container only (the reference tree is not on the GPU box), on the CPU
emulation of the kernels (tests/emu/): it proves the call surface -- imports, signatures, attribute names, return types,
autograd wiring, state the drivers save -- not speed.

Two drop-in routes, both exercised through `/root/reference/run_pipeline_lightgcn.py::train` and `/root/reference/training.py`:
  (a) install_aliases(): torch_sparse / torch_geometric resolve to this package, the reference's own model code
      (model/lightgcn.py, model/encoder_decoder.py, model/layers.py) runs on the package's SparseTensor / matmul / gcn_norm /
      SAGEConv / to_hetero;
  (b) patch_driver(): the driver keeps its code, the names it imported from model.* / utils.metrics_lightgcn /
      data.lightgcn_loader are swapped for the package's fused drop-ins.
"""
import importlib
import json
import os
import platform
import random
import sys

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = [pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)"),
              pytest.mark.skipif(platform.machine() != "x86_64", reason="the emulator's context switch is x86-64 only")]

REF_TOP = ("config", "model", "utils", "data", "training", "run_pipeline_lightgcn", "reporting")


class FakeGraph:
    """Stands for the pickled PyG HeteroData of data/derived/test_graph.pt: create_dataloaders_lightgcn only calls
    `.to_homogeneous().edge_index` on it (data/lightgcn_loader.py:55)."""

    def __init__(self, edge_index):
        self.edge_index = edge_index

    def to_homogeneous(self):
        return self


@pytest.fixture
def ref_env(tmp_path, monkeypatch):
    """cwd with data/derived fixtures, the reference tree importable, the aliases installed, kernels emulated."""
    from tests.emu.harness import emulated
    import laplace_gnn_recommendation_b200 as lg
    from laplace_gnn_recommendation_b200 import aliases
    U, I, E = 40, 150, 600           # I >= num_recommendations + the longest seen list (the reference's topk asks for k + len(seen))
    gen = torch.Generator().manual_seed(11)
    users = (torch.rand(E, generator=gen) ** 1.5 * U).long().clamp(max=U - 1)
    users[:U] = torch.arange(U)                                     # every user has an edge: max(user id) + 1 == U
    items = torch.randint(0, I, (E,), generator=gen)
    keys = torch.unique(users * I + items)
    homo = torch.stack([keys // I, keys % I + U])                   # to_homogeneous(): item ids follow the user ids
    d = tmp_path / "data" / "derived"
    d.mkdir(parents=True)
    torch.serialization.add_safe_globals([FakeGraph])
    torch.save(FakeGraph(homo), d / "test_graph.pt")
    json.dump({str(i): i for i in range(U)}, open(d / "customer_id_map_forward.json", "w"))
    json.dump({str(i): i for i in range(I)}, open(d / "article_id_map_forward.json", "w"))
    monkeypatch.chdir(tmp_path)
    monkeypatch.syspath_prepend(REF)
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in REF_TOP + ("torch_sparse", "torch_geometric")}
    for k in saved:
        del sys.modules[k]
    with emulated():
        assert aliases.install_aliases(force=True)
        yield dict(U=U, I=I, homo=homo, lg=lg, aliases=aliases, tmp=tmp_path)
    for k in [k for k in sys.modules if k.split(".")[0] in REF_TOP + ("torch_sparse", "torch_geometric")]:
        del sys.modules[k]
    sys.modules.update(saved)


def _seed():
    torch.manual_seed(0); random.seed(0); np.random.seed(0)


def test_reference_lightgcn_driver_runs_unmodified(ref_env):
    """run_pipeline_lightgcn.train(config): the reference's loader, model, loss, evaluation and candidate dump, line for line."""
    driver = importlib.import_module("run_pipeline_lightgcn")
    config = importlib.import_module("config")
    assert driver.__file__.startswith(REF) and sys.modules["model.lightgcn"].__file__.startswith(REF)
    assert sys.modules["torch_sparse"].SparseTensor is ref_env["lg"].SparseTensor        # the alias layer, not the oracle
    cfg = config.LightGCNConfig(epochs=7, k=5, hidden_layer_size=32, learning_rate=1e-2, save_model=False, batch_size=64,
                                num_iterations=3, eval_every=3, lr_decay_every=4, Lambda=1e-6, show_graph=False, num_recommendations=5)
    # (a) the reference's own LightGCN / bpr_loss / sampler on the aliased SparseTensor / matmul / gcn_norm
    _seed()
    stats_a = driver.train(cfg)
    out_a = torch.load("data/derived/lightgcn_output.pt")
    emb_a = torch.load("data/derived/users_emb_final_lightgcn.pt")
    assert len(out_a) == ref_env["U"] and all(v.numel() == 5 for v in out_a.values())
    assert emb_a.shape == (ref_env["U"], 32) and np.isfinite(stats_a.loss)
    # (b) the same driver with the package's drop-ins swapped in: fused LightGCN, fused bpr_loss, device-side sampler
    swapped = ref_env["aliases"].patch_driver(driver)
    assert {"LightGCN", "bpr_loss", "sample_mini_batch", "structured_negative_sampling", "make_predictions_for_user"} <= set(swapped)
    _seed()
    stats_b = driver.train(cfg)
    out_b = torch.load("data/derived/lightgcn_output.pt")
    # same seeds, same draws, same arithmetic up to fp32 summation order: the two routes walk the same trajectory
    assert stats_b.loss == pytest.approx(stats_a.loss, rel=1e-4, abs=1e-6)
    for key in ("recall_val", "recall_test", "precision_val", "precision_test"):
        assert getattr(stats_b, key) == pytest.approx(getattr(stats_a, key), rel=1e-4, abs=1e-6), key
    same = sum(torch.equal(out_a[u], out_b[u]) for u in out_a)
    assert same >= 0.9 * len(out_a)              # top-5 lists agree (ties / 1-ulp score differences may reorder a few)


def _ranking_batches(lg, aliases, n_batches, embedding, seed=5):
    """Small HeteroData batches shaped like the reference's loaders produce (features, buys / rev_buys, label edges)."""
    from laplace_gnn_recommendation_b200 import hetero
    gen = torch.Generator().manual_seed(seed)
    Nc, Na, E, L = 30, 45, 260, 40
    out = []
    for _ in range(n_batches):
        b = aliases.HeteroData()
        if embedding:          # integer categorical columns (utils/get_info.py:17-31 reads their maxima)
            b["customer"].x = torch.stack([torch.randint(0, m, (Nc,), generator=gen) for m in (9, 2, 90, 4)], dim=1)
            b["article"].x = torch.stack([torch.randint(0, m, (Na,), generator=gen) for m in (40, 12, 7)], dim=1)
        else:
            b["customer"].x = torch.randn(Nc, 6, generator=gen)
            b["article"].x = torch.randn(Na, 5, generator=gen)
        e = torch.stack([torch.randint(0, Nc, (E,), generator=gen), torch.randint(0, Na, (E,), generator=gen)])
        b[hetero.EDGE_KEY].edge_index = e
        b[hetero.REV_EDGE_KEY].edge_index = e.flip(0).contiguous()
        b[hetero.EDGE_KEY].edge_label_index = torch.stack([torch.randint(0, Nc, (L,), generator=gen), torch.randint(0, Na, (L,), generator=gen)])
        b[hetero.EDGE_KEY].edge_label = (torch.rand(L, generator=gen) > 0.7).long()
        out.append(b)
    return out


@pytest.mark.parametrize("embedding", [False, True])
def test_reference_ranking_training_loop_runs_unmodified(ref_env, embedding):
    """training.py::train_with_dataloader / test_with_dataloader, unmodified, on (a) the reference's own Encoder_Decoder_Model over
    the aliased SAGEConv / to_hetero and (b) the package's drop-in model -- incl. the feature-embedding route (D1:
    embedding=True, FeatureInfo from utils/get_info.py, Embedding(max_norm=1) renormalising its rows in place)."""
    lg, aliases = ref_env["lg"], ref_env["aliases"]
    training = importlib.import_module("training")
    ref_model = importlib.import_module("model.encoder_decoder")
    ref_layers = importlib.import_module("model.layers")
    get_info = importlib.import_module("utils.get_info")
    assert training.__file__.startswith(REF) and ref_model.__file__.startswith(REF)
    from laplace_gnn_recommendation_b200 import hetero
    batches = _ranking_batches(lg, aliases, 3, embedding)
    metadata = (["customer", "article"], [hetero.EDGE_KEY, hetero.REV_EDGE_KEY])
    feature_info = {}
    if embedding:              # the reference derives it from the FULL graph (run_pipeline.py:41): every category value occurs there
        full = aliases.HeteroData()
        full["customer"].x = torch.tensor([[8, 1, 89, 3], [0, 0, 0, 0]])
        full["article"].x = torch.tensor([[39, 11, 6], [0, 0, 0]])
        full[hetero.EDGE_KEY].edge_index = batches[0][hetero.EDGE_KEY].edge_index
        feature_info = get_info.get_feature_info(full)
        assert feature_info["customer"].num_feat == 4 and feature_info["customer"].embedding_size == [4, 2, 12, 4]
    losses, states, metrics = {}, {}, {}
    for route, (Model, sage, lin) in {"reference model on aliases": (ref_model.Encoder_Decoder_Model, ref_layers.get_SAGEConv_layers, ref_layers.get_linear_layers),
                                      "package drop-in": (lg.Encoder_Decoder_Model, lg.get_SAGEConv_layers, lg.get_linear_layers)}.items():
        _seed()
        model = Model(encoder_layers=sage(2, 16, 8, "mean"), decoder_layers=lin(2, 16, 16, 1), feature_info=feature_info,
                      metadata=metadata, embedding=embedding, heterogeneous_prop_agg_type="sum", batch_normalize=True,
                      p_dropout_edges=None, p_dropout_features=None)
        model.initialize_encoder_input_size(_ranking_batches(lg, aliases, 1, embedding)[0])
        opt = torch.optim.Adam(model.parameters(), lr=1e-2)
        ls = []
        for epoch in range(2):
            ls += training.train_with_dataloader(model, opt, _ranking_batches(lg, aliases, 3, embedding), epoch, "cpu")
        recall, precision = training.test_with_dataloader("VAL", model, _ranking_batches(lg, aliases, 2, embedding, seed=9), "cpu", k=3, break_at=None)
        losses[route], states[route], metrics[route] = ls, {k: v.clone() for k, v in model.state_dict().items()}, (recall, precision)
        assert len(ls) == 6 and all(np.isfinite(ls)) and 0.0 <= recall <= 1.0
        if embedding:              # the quirk of model/encoder_decoder.py:101-114: the embedding tables are NOT parameters / state
            assert not any("embedding_layers" in k for k in model.state_dict())
            w = model.embedding_layers["customer"][0].weight
            assert float(w.norm(dim=1).max()) <= 1.0 + 1e-4          # max_norm=1 renormalised the looked-up rows in place
    a, b = list(losses.values())
    assert a == pytest.approx(b, rel=2e-4, abs=1e-6)             # same seeds, same init order, same arithmetic: same trajectory
    sa, sb = list(states.values())
    assert sa.keys() == sb.keys()                                 # identical state_dict keys (checkpoints interchange)
    for k in sa:
        torch.testing.assert_close(sa[k], sb[k], rtol=2e-3, atol=2e-5)
    assert list(metrics.values())[0] == pytest.approx(list(metrics.values())[1], abs=1e-6)
