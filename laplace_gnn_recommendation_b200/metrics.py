"""LightGCN evaluation with the reference's call surface (run_pipeline_lightgcn.py:20-73,
utils/metrics_lightgcn.py:79-122, utils/metrics.py:6-57).

The heavy parts run on the GPU kernels (forward propagation, the O(E) negative-sample rejection test, BPR over every
edge of the split, score + top-k with seen-item exclusion); the recall / precision / NDCG arithmetic is a handful of
device-agnostic torch ops on the [n_users, k] id matrix (the reference does it with per-user Python loops on the CPU).
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from . import _lib
from .bpr import bpr_indexed
from .loader import structured_negative_sampling
from .sparse import SparseTensor, gcn_norm
from .topk import SeenItems, recommend_topk


def recall_precision_ndcg(topk_ids: torch.Tensor, users: torch.Tensor, edge_index: torch.Tensor, num_items: int, k: int
                          ) -> Tuple[float, float, float]:
    """recall@k, precision@k, ndcg@k exactly as RecallPrecision_ATk / NDCGatK_r compute them.

    topk_ids: [n_users, k] recommended item ids for ``users`` (sorted unique user ids of ``edge_index``).
    The per-user ground-truth size counts duplicate edges, like ``len(edge_index[1][edge_index[0] == user])``."""
    dev = topk_ids.device
    users = users.to(dev)
    ei = edge_index.to(dev)
    pos_keys = ei[0] * num_items + ei[1]
    r = torch.isin(users.unsqueeze(1) * num_items + topk_ids, pos_keys) & (topk_ids >= 0)        # hit matrix [n_users, k]
    n_liked = torch.bincount(torch.searchsorted(users, ei[0]), minlength=users.numel()).to(torch.float32)
    num_correct = r.sum(dim=-1).float()
    recall = torch.mean(num_correct / n_liked)
    precision = torch.mean(num_correct) / k
    discount = 1.0 / torch.log2(torch.arange(2, k + 2, device=dev, dtype=torch.float32))
    ideal_len = torch.clamp(n_liked, max=k)
    ideal = (torch.arange(k, device=dev).unsqueeze(0) < ideal_len.unsqueeze(1)).float()
    idcg = (ideal * discount).sum(dim=1)
    dcg = (r.float() * discount).sum(dim=1)
    idcg[idcg == 0.0] = 1.0
    ndcg = dcg / idcg
    ndcg[torch.isnan(ndcg)] = 0.0
    return recall.item(), precision.item(), torch.mean(ndcg).item()


def _recall_precision_ndcg_from_hits(r: torch.Tensor, n_liked: torch.Tensor, k: int) -> Tuple[float, float, float]:
    dev = r.device
    num_correct = r.sum(dim=-1).float()
    recall = torch.mean(num_correct / n_liked)
    precision = torch.mean(num_correct) / k
    discount = 1.0 / torch.log2(torch.arange(2, k + 2, device=dev, dtype=torch.float32))
    ideal = (torch.arange(k, device=dev).unsqueeze(0) < torch.clamp(n_liked, max=k).unsqueeze(1)).float()
    idcg = (ideal * discount).sum(dim=1)
    dcg = (r.float() * discount).sum(dim=1)
    idcg[idcg == 0.0] = 1.0
    ndcg = dcg / idcg
    ndcg[torch.isnan(ndcg)] = 0.0
    return recall.item(), precision.item(), torch.mean(ndcg).item()


def get_metrics_universal(model_output: torch.Tensor, edge_index: torch.Tensor, edge_label_index: torch.Tensor,
                          exclude_edge_indices: List[torch.Tensor], k: int) -> Tuple[float, float, float]:
    """Ranking-model metrics with the reference's signature and semantics (utils/metrics_encoder_decoder.py:29-86, called
    by training.py:38-56 on ``Encoder_Decoder_Model.infer`` output), vectorised on the device of ``model_output`` instead
    of per-edge / per-user Python loops on the CPU.  Reference quirks kept: excluded pairs address row
    ``user`` / column ``item`` of the padded [users, max_candidates] score matrix; the top-k entries are column positions
    and are matched against the item ids of ``edge_index``.  The caller's tensor is never modified."""
    ratings = model_output.detach()
    ratings = (ratings.unsqueeze(0) if ratings.dim() < 2 else ratings).clone()
    dev = ratings.device
    for ex in exclude_edge_indices:
        if ex.numel():
            ex = ex.to(dev)
            ratings[ex[0], ex[1]] = -(1 << 10)
    _, top = torch.topk(ratings, k=k)
    ei = edge_index.detach().to(dev)
    users = edge_label_index.detach().to(dev)[0].unique(sorted=True)
    top = top[: users.numel()]
    m = int(max(ratings.shape[1], int(ei[1].max()) + 1 if ei.numel() else 1))
    r = torch.isin(users.unsqueeze(1) * m + top, ei[0] * m + ei[1])                      # hit matrix [n_users, k]
    slot = torch.searchsorted(users, ei[0]).clamp(max=max(users.numel() - 1, 0))
    mine = users[slot] == ei[0] if users.numel() else torch.zeros_like(ei[0], dtype=torch.bool)
    n_liked = torch.bincount(slot[mine], minlength=users.numel()).to(torch.float32)    # duplicate edges counted, like len()
    return _recall_precision_ndcg_from_hits(r, n_liked, k)


def get_metrics_lightgcn(model, edge_index: torch.Tensor, exclude_edge_indices: List[torch.Tensor], k: int
                         ) -> Tuple[float, float, float]:
    """utils/metrics_lightgcn.py:79-122: top-k on the LAYER-0 tables (reference quirk), seen items excluded."""
    Wu, Wi = model.users_emb.weight.detach(), model.items_emb.weight.detach()
    _lib.require_cuda(Wu, Wi)
    dev = Wu.device
    excl = torch.cat([e.to(dev) for e in exclude_edge_indices], dim=1)
    seen = SeenItems(excl, model.num_users, model.num_items)
    users = edge_index[0].to(dev).unique()
    ids = recommend_topk(Wu, Wi, users, k, seen)
    return recall_precision_ndcg(ids, users, edge_index, model.num_items, k)


@torch.no_grad()
def evaluation(model, edge_index: torch.Tensor, sparse_edge_index: SparseTensor, exclude_edge_indices, k: int,
               lambda_val: float):
    """run_pipeline_lightgcn.py:20-73: (bpr loss over every edge of the split, recall@k, precision@k, ndcg@k)."""
    Wu, Wi = model.users_emb.weight, model.items_emb.weight
    _lib.require_cuda(Wu, Wi)
    dev = Wu.device
    from .lightgcn import propagate_forward
    g = gcn_norm(sparse_edge_index, add_self_loops=model.add_self_loops).csr()
    E0 = model._table_for(Wu, Wi)
    E_f = propagate_forward(g, E0, model.num_iterations)
    u, p, n = structured_negative_sampling(edge_index, num_nodes=int(torch.max(edge_index[1])),
                                           contains_neg_self_loops=False)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    bpr_indexed(E_f, E0, model.num_users, u.to(dev), p.to(dev), n.to(dev), lambda_val, loss=loss)
    recall, precision, ndcg = get_metrics_lightgcn(model, edge_index, exclude_edge_indices, k)
    return loss.item(), recall, precision, ndcg
