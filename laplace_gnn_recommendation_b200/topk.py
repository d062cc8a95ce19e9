"""Candidate generation on the GPU with the reference's call surface
(utils/metrics_lightgcn.py:48-61,125-142; run_pipeline_lightgcn.py:211-222).

``recommend_topk`` scores a block of users against every item on the layer-0 tables (the reference's
quirk), masks each user's seen items and returns the first k unseen item ids in (score desc, id asc)
order -- the same list ``topk(k + |seen|)`` + order-preserving ``setdiff1d`` gives.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import torch

from . import _lib
from ._lib import check, ptr, stream
from .csr import DeviceCSR


class SeenItems:
    """Per-user seen-item lists as a device CSR (replaces the O(U*E) boolean-mask loop of create_adj_dict)."""

    def __init__(self, edge_index: torch.Tensor, num_users: int, num_items: int):
        _lib.require_cuda(edge_index)
        self.csr = DeviceCSR.from_coo(edge_index[0], edge_index[1], num_users, num_items, chunk=0)
        self.num_users, self.num_items = num_users, num_items


def recommend_topk(user_emb: torch.Tensor, item_emb: torch.Tensor, users: torch.Tensor, k: int,
                   seen: Optional[SeenItems] = None, block: int = 2048, return_scores: bool = False):
    """ids[len(users), k] (int64, -1 padded when fewer than k unseen items exist)."""
    _lib.require_cuda(user_emb, item_emb, users)
    user_emb, item_emb = _lib.f32c(user_emb.detach()), _lib.f32c(item_emb.detach())
    users = _lib.i64c(users)
    I, d = item_emb.shape
    nu = users.numel()
    dev = user_emb.device
    out = torch.empty(nu, k, dtype=torch.int64, device=dev)
    scores = torch.empty(nu, k, dtype=torch.float32, device=dev) if return_scores else None
    # keep the score scratch <= ~1 GiB
    block = max(1, min(block, (1 << 28) // max(I, 1)))
    ws = torch.empty(min(block, max(nu, 1)) * I, dtype=torch.float32, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        for s in range(0, nu, block):
            e = min(s + block, nu)
            check(lib.lgb_topk_exclude(ptr(user_emb), ptr(item_emb), users[s:e].data_ptr(), e - s, I, d,
                                       ptr(seen.csr.rowptr) if seen else None, ptr(seen.csr.colidx) if seen else None,
                                       k, out[s:e].data_ptr(), scores[s:e].data_ptr() if scores is not None else None,
                                       ptr(ws), stream()), "topk_exclude")
            _lib.count_launch()
    return (out, scores) if return_scores else out


def make_predictions_for_user(user_embeddings: torch.Tensor, article_embeddings: torch.Tensor, user_id: int,
                              positive_items_for_user: Dict[int, torch.Tensor], num_recommendations: int) -> torch.Tensor:
    """Reference signature (utils/metrics_lightgcn.py:125-142) for a single user; returns a CPU LongTensor."""
    dev = user_embeddings.device
    seen_items = positive_items_for_user.get(user_id)
    seen = None
    if seen_items is not None and len(seen_items) > 0:
        it = seen_items.to(dev).to(torch.int64)
        ei = torch.stack([torch.zeros_like(it), it])
        seen = SeenItems(ei, 1, article_embeddings.shape[0])
        ids = recommend_topk(user_embeddings[user_id:user_id + 1], article_embeddings,
                             torch.zeros(1, dtype=torch.int64, device=dev), num_recommendations, seen)
    else:
        ids = recommend_topk(user_embeddings, article_embeddings, torch.tensor([user_id], device=dev),
                             num_recommendations, None)
    ids = ids[0]
    return ids[ids >= 0].cpu()


def topk_dict(user_emb: torch.Tensor, item_emb: torch.Tensor, users: Iterable[int], k: int,
              seen: Optional[SeenItems]) -> Dict[int, torch.Tensor]:
    """``{user: LongTensor[k]}`` -- the format run_pipeline_lightgcn.py:222 saves for LightGCNMatcher."""
    users_t = torch.as_tensor(list(users), dtype=torch.int64, device=user_emb.device)
    ids = recommend_topk(user_emb, item_emb, users_t, k, seen).cpu()
    return {int(u): row[row >= 0] for u, row in zip(users_t.tolist(), ids)}
