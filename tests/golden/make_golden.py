"""Generate tests/golden/*.pt by running the REAL reference modules from /root/reference.

Run once in the build container (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py

torch_sparse / torch_geometric are not installable here, so their names are served by
oracle/ref_shims.py (backed by the oracle's restatement -> those pieces stay "parity unpinned");
every line of the reference's OWN code runs unmodified: model/lightgcn.py, utils/metrics_lightgcn.py,
utils/metrics.py, utils/tensor.py, data/lightgcn_loader.py, model/encoder_decoder.py, model/layers.py.
The fixtures are small tensors + the inputs that produced them, so tests can re-run both the oracle
and the CUDA path on identical inputs.
"""
import os
import random
import sys

import numpy as np
import torch as t

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)
from oracle import ref_shims  # noqa: E402

ref_shims.install()
# the reference's top-level packages (tests/, utils/, data/, model/) must win over ours while importing it
sys.path.insert(0, REF)
for name in [m for m in sys.modules if m == "tests" or m.startswith("tests.")]:
    del sys.modules[name]

from model.lightgcn import LightGCN  # noqa: E402
from utils.metrics_lightgcn import bpr_loss, make_predictions_for_user, create_adj_dict, get_metrics_lightgcn  # noqa: E402
from utils.metrics import RecallPrecision_ATk, NDCGatK_r  # noqa: E402
from utils.tensor import padded_stack, difference_1d  # noqa: E402
from data.lightgcn_loader import split, both_indexes_from_zero, sample_mini_batch  # noqa: E402
from torch_sparse import SparseTensor  # noqa: E402  (shim)

# manual fixtures of the reference's tests/data_generator.py:129-159 (users -> articles)
FIXTURES = {
    "random": (3, 6, [[0, 0, 0, 1, 1, 2, 2], [0, 2, 4, 1, 5, 3, 0]]),
    "star": (5, 4, [[0, 0, 0, 0, 1, 2, 3, 4], [0, 1, 2, 3, 0, 1, 2, 3]]),
}


def seed(s):
    t.manual_seed(s)
    random.seed(s)
    np.random.seed(s)


def lightgcn_case(name, U, I, edges, d, K, wiring, B, lam, s):
    """Reference LightGCN.forward + reference bpr_loss + autograd on one graph."""
    seed(s)
    ei = t.tensor(edges, dtype=t.long)
    if wiring == "R":      # data/lightgcn_loader.py:65-69
        row, col = ei[0], ei[1]
    else:                  # symmetric bipartite
        row, col = t.cat([ei[0], ei[1] + U]), t.cat([ei[1] + U, ei[0]])
    adj = SparseTensor(row=row, col=col, sparse_sizes=(U + I, U + I))
    model = LightGCN(U, I, embedding_dim=d, num_iterations=K)
    Wu0, Wi0 = model.users_emb.weight.detach().clone(), model.items_emb.weight.detach().clone()
    u_f, u_0, i_f, i_0 = model.forward(adj)
    E = ei.shape[1]
    pick = t.randint(E, (B,))
    ub, pb = ei[0][pick], ei[1][pick]
    nb = t.randint(I, (B,))
    loss = bpr_loss(u_f[ub], u_0[ub], i_f[pb], i_0[pb], i_f[nb], i_0[nb], lam)
    loss.backward()
    return dict(name=name, U=U, I=I, d=d, K=K, wiring=wiring, lam=lam, row=row, col=col,
                Wu=Wu0, Wi=Wi0, u=ub, p=pb, n=nb, u_final=u_f.detach(), i_final=i_f.detach(),
                loss=loss.detach(), dWu=model.users_emb.weight.grad.clone(),
                dWi=model.items_emb.weight.grad.clone())


def main():
    out = {}

    # ---- G1: LightGCN forward/backward on the reference's own fixtures + a generated graph
    cases = []
    for name, (U, I, edges) in FIXTURES.items():
        for wiring in ("R", "S"):
            for d in (4, 64):
                cases.append(lightgcn_case(f"{name}_{wiring}_d{d}", U, I, edges, d, 3, wiring, 16, 1e-6, 11))
    seed(5)  # tests/data_generator.py:162-191 style iid randint graph, duplicates present
    U, I = 9, 19
    nconn = int(U * I * 0.5)
    gen = [t.randint(0, U, (nconn,)).tolist(), t.randint(0, I, (nconn,)).tolist()]
    for wiring in ("R", "S"):
        cases.append(lightgcn_case(f"generated_{wiring}_d64", U, I, gen, 64, 3, wiring, 128, 1e-6, 12))
        cases.append(lightgcn_case(f"generated_{wiring}_d32K4", U, I, gen, 32, 4, wiring, 128, 1e-4, 13))
    out["lightgcn"] = cases

    # ---- G2: bpr_loss alone (pure reference code), incl. the softplus threshold branch (x > 20)
    bpr = []
    for (B, d, scale, lam, s) in [(128, 64, 0.1, 1e-6, 1), (37, 32, 1.0, 1e-3, 2), (64, 64, 3.0, 0.0, 3), (1, 8, 0.5, 1e-2, 4)]:
        seed(s)
        xs = [(t.randn(B, d) * scale).requires_grad_(True) for _ in range(6)]
        loss = bpr_loss(*xs, lam)
        loss.backward()
        bpr.append(dict(inputs=[x.detach().clone() for x in xs], lam=lam, loss=loss.detach(),
                        grads=[x.grad.clone() for x in xs]))
    out["bpr"] = bpr

    # ---- G3: top-k with seen-item exclusion + metrics (pure reference code)
    seed(7)
    U, I, d, k = 40, 60, 16, 12
    Wu, Wi = t.randn(U, d), t.randn(I, d)
    excl = t.stack([t.randint(0, U, (300,)), t.randint(0, I, (300,))])
    evale = t.stack([t.randint(0, U, (120,)), t.randint(0, I, (120,))])
    seen = create_adj_dict(excl)
    preds = t.stack([make_predictions_for_user(Wu, Wi, u, seen, k) for u in range(U)])

    class _M:  # get_metrics_lightgcn only touches model.users_emb.weight / items_emb.weight
        pass
    m = _M()
    m.users_emb = t.nn.Embedding.from_pretrained(Wu)
    m.items_emb = t.nn.Embedding.from_pretrained(Wi)
    recall, precision, ndcg = get_metrics_lightgcn(m, evale, [excl], k)
    out["topk"] = dict(Wu=Wu, Wi=Wi, exclude=excl, eval=evale, k=k, preds=preds,
                       recall=recall, precision=precision, ndcg=ndcg)
    r = (t.rand(5, 4) > 0.5)
    gt = [t.arange(n) for n in (1, 3, 2, 6, 4)]
    out["metrics"] = dict(r=r, gt_len=[1, 3, 2, 6, 4], k=4, rp=RecallPrecision_ATk(gt, r, 4), ndcg=NDCGatK_r(gt, r, 4))
    out["difference_1d"] = dict(a=t.tensor([5, 3, 9, 1, 7]), b=t.tensor([9, 5]),
                                out=difference_1d(t.tensor([5, 3, 9, 1, 7]), t.tensor([9, 5]), assume_unique=True))

    # ---- G4: loader integer paths: both_indexes_from_zero, split (sklearn), sample_mini_batch
    seed(21)
    U, I, E = 30, 50, 400
    hom = t.stack([t.randint(0, U, (E,)), t.randint(0, I, (E,)) + U])
    hom[0, 0] = U - 1  # make max(row0)+1 == U as to_homogeneous() guarantees
    ei = both_indexes_from_zero(hom)
    tr, va, te, _ = split(ei)
    seed(42)
    batches = [t.stack(sample_mini_batch(16, tr)) for _ in range(3)]
    out["loader"] = dict(hom=hom, edge_index=ei, train=tr, val=va, test=te, seed=42, batch_size=16,
                         batches=batches)

    # ---- G5: decoder / infer helpers (reference code), encoder through the shim
    from model.encoder_decoder import Encoder_Decoder_Model, EdgeDecoder
    from model.layers import get_SAGEConv_layers, get_linear_layers
    from utils.constants import Constants
    seed(31)
    Nc, Na, Fc, Fa, Es, L = 12, 17, 8, 12, 90, 25
    x = {Constants.node_user: t.randn(Nc, Fc), Constants.node_item: t.randn(Na, Fa)}
    e = t.stack([t.randint(0, Nc, (Es,)), t.randint(0, Na, (Es,))])
    eid = {Constants.edge_key: e, Constants.rev_edge_key: e.flip(0)}
    eli = t.stack([t.randint(0, Nc, (L,)), t.randint(0, Na, (L,))])
    metadata = ([Constants.node_user, Constants.node_item], [Constants.edge_key, Constants.rev_edge_key])
    hetero = []
    for conv_aggr in ("add", "mean", "max"):
        seed(32)
        model = Encoder_Decoder_Model(
            encoder_layers=get_SAGEConv_layers(2, 16, 8, conv_aggr),
            decoder_layers=get_linear_layers(2, 16, 16, 1),
            feature_info={}, metadata=metadata, embedding=False, heterogeneous_prop_agg_type="sum",
            batch_normalize=True, p_dropout_edges=None, p_dropout_features=None)
        model.train()
        logits = model(dict(x), eid, eli)
        labels = (t.rand(L) > 0.5).float()
        loss = t.nn.BCEWithLogitsLoss()(logits, labels)
        loss.backward()
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
        inf = model.infer(dict(x), eid, eli)
        hetero.append(dict(conv_aggr=conv_aggr, state_dict=sd, logits=logits.detach(), labels=labels,
                           loss=loss.detach(), grads=grads, infer=inf))
    out["hetero"] = dict(x=x, edge_index=eid, edge_label_index=eli, cases=hetero,
                         node_user=Constants.node_user, node_item=Constants.node_item,
                         edge_key=Constants.edge_key, rev_edge_key=Constants.rev_edge_key)
    seed(33)
    zc, za = t.randn(Nc, 8), t.randn(Na, 8)
    dec_layers = get_linear_layers(2, 16, 16, 1)
    dec = EdgeDecoder(dec_layers, None)
    out["decoder"] = dict(z_user=zc, z_item=za, eli=eli,
                          linears=[(l.weight.detach().clone(), l.bias.detach().clone()) for l in dec_layers],
                          out=dec({Constants.node_user: zc, Constants.node_item: za}, eli).detach())
    out["padded_stack"] = dict(out=padded_stack([t.arange(3.0), t.arange(5.0), t.arange(1.0)], value=-(1 << 50)))

    path = os.path.join(HERE, "reference_golden.pt")
    t.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
