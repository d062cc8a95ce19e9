"""CPU oracle (TEST INFRASTRUCTURE, not product) for LightGCN candidate generation and metrics.

* ``make_predictions_for_user``  utils/metrics_lightgcn.py:125-142 (+ ``difference_1d`` utils/tensor.py:16-21)
* ``create_adj_dict``            utils/metrics_lightgcn.py:48-61
* ``get_metrics_lightgcn``       utils/metrics_lightgcn.py:79-122
* ``RecallPrecision_ATk`` / ``NDCGatK_r``  utils/metrics.py:6-57

Pinned against the real reference functions by tests/golden/make_golden.py.
Quirk kept on purpose: ``np.setdiff1d(..., assume_unique=True)`` preserves the top-k ORDER
(it only masks), and candidates are scored on the layer-0 tables.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

Tensor = torch.Tensor


def create_adj_dict(edge_index: Tensor, from_nodes: Optional[Tensor] = None) -> Dict[int, Tensor]:
    users = edge_index[0].unique(sorted=True) if from_nodes is None else from_nodes
    return {int(u): edge_index[1][edge_index[0] == u] for u in users}


def predictions_for_user(user_emb: Tensor, item_emb: Tensor, user_id: int,
                         seen: Dict[int, Tensor], k: int) -> Tensor:
    """scores = Wu[u] @ Wi^T; topk(k + |seen_u|); drop seen ids keeping order; first k."""
    ignore = seen[user_id] if user_id in seen else torch.tensor([])
    scores = user_emb[user_id] @ item_emb.T
    _, idx = torch.topk(scores, k=k + len(ignore))
    kept = np.setdiff1d(idx.detach().numpy(), ignore.detach().numpy(), assume_unique=True)
    return torch.tensor(kept)[:k]


def recall_precision_at_k(ground_truth: List[Tensor], r: Tensor, k: int):
    num_correct = torch.sum(r, dim=-1).float()
    liked = torch.Tensor([len(row) for row in ground_truth])
    recall = torch.mean(num_correct / liked)
    precision = torch.mean(num_correct) / k
    return recall.item(), precision.item()


def ndcg_at_k(ground_truth: List[Tensor], r: Tensor, k: int) -> float:
    test_matrix = torch.zeros((len(r), k))
    for i, items in enumerate(ground_truth):
        test_matrix[i, : min(len(items), k)] = 1
    discount = 1.0 / torch.log2(torch.arange(2, k + 2))
    idcg = torch.sum(test_matrix * discount, dim=1)
    dcg = torch.sum(r * discount, dim=1)
    idcg[idcg == 0.0] = 1.0
    ndcg = dcg / idcg
    ndcg[torch.isnan(ndcg)] = 0.0
    return torch.mean(ndcg).item()


def metrics_lightgcn(user_emb: Tensor, item_emb: Tensor, edge_index: Tensor,
                     exclude_edge_indices: List[Tensor], k: int):
    """recall@k, precision@k, ndcg@k for every user of ``edge_index`` (utils/metrics_lightgcn.py:79-122)."""
    excluded = create_adj_dict(torch.cat(exclude_edge_indices, dim=1))
    users = edge_index[0].unique()
    top = {int(u): predictions_for_user(user_emb, item_emb, int(u), excluded, k) for u in users}
    positives = create_adj_dict(edge_index, from_nodes=users)
    pos_list = [positives[int(u)] for u in users]
    r = torch.stack([torch.isin(top[int(u)], positives[int(u)]) for u in users])
    recall, precision = recall_precision_at_k(pos_list, r, k)
    return recall, precision, ndcg_at_k(pos_list, r, k), top


def metrics_universal(model_output: Tensor, edge_index: Tensor, edge_label_index: Tensor,
                      exclude_edge_indices: List[Tensor], k: int):
    """utils/metrics_encoder_decoder.py:29-86 (ranking model, called from training.py:38-56 on ``model.infer`` output),
    restated with the same per-user loops.  Quirks kept: excluded (user, item) pairs address ROW ``user`` / COLUMN ``item``
    of the padded [users, max_candidates] matrix and are set to -(1 << 10); the top-k entries are column positions and
    are matched against the ITEM IDS of ``edge_index``; the first ``len(users)`` rows are scored."""
    ratings = model_output.detach().clone()
    if ratings.dim() < 2:
        ratings = ratings.unsqueeze(0)
    for ex in exclude_edge_indices:
        for j in range(ex.shape[1]):
            ratings[int(ex[0][j]), int(ex[1][j])] = -(1 << 10)
    _, top = torch.topk(ratings, k=k)
    users = edge_label_index[0].unique(sorted=True)
    truth = [edge_index[1][edge_index[0] == u] for u in users]
    r = torch.stack([torch.isin(top[i], truth[i]) for i in range(len(users))])
    recall, precision = recall_precision_at_k(truth, r, k)
    return recall, precision, ndcg_at_k(truth, r, k)
