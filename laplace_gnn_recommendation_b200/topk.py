"""Candidate generation on the GPU with the reference's call surface
(utils/metrics_lightgcn.py:48-61,125-142; run_pipeline_lightgcn.py:211-222).

``recommend_topk`` scores a block of users against every item on the layer-0 tables (the reference's
quirk), masks each user's seen items and returns the first k unseen item ids in (score desc, id asc)
order -- the same list ``topk(k + |seen|)`` + order-preserving ``setdiff1d`` gives.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import torch

import os

from . import _lib, csr as _csr
from ._lib import check, ptr, stream
from .csr import DeviceCSR

# scoring kernel of recommend_topk: "cta" = one CTA per user (measured in round 1), "tiled" = 8 users per CTA share every
# item row (FMA-bound instead of L2-bound), "auto" = the first call with enough users runs BOTH on its first block, requires
# bit-identical ids and scores, times them with CUDA events and keeps the faster for that (items, d, k) shape.
TOPK_MODE = os.environ.get("LGB_TOPK_MODE", "auto")
_TOPK_CHOICE: Dict[tuple, str] = {}
_TOPK_AUTO_MIN_USERS = 256      # smaller calls stay on the per-user kernel (nothing to gain, and no probing cost)


class SeenItems:
    """Per-user seen-item lists as a device CSR (replaces the O(U*E) boolean-mask loop of create_adj_dict)."""

    def __init__(self, edge_index: torch.Tensor, num_users: int, num_items: int):
        _lib.require_cuda(edge_index)
        self.csr = DeviceCSR.from_coo(edge_index[0], edge_index[1], num_users, num_items, chunk=0)
        self.num_users, self.num_items = num_users, num_items


def recommend_topk(user_emb: torch.Tensor, item_emb: torch.Tensor, users: torch.Tensor, k: int,
                   seen: Optional[SeenItems] = None, block: int = 2048, return_scores: bool = False,
                   mode: Optional[str] = None):
    """ids[len(users), k] (int64, -1 padded when fewer than k unseen items exist).  ``mode``: see TOPK_MODE."""
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
    sp, si = (ptr(seen.csr.rowptr), ptr(seen.csr.colidx)) if seen else (None, None)

    def run(fn, s, e, ids, sc):
        with torch.cuda.device(dev):
            check(fn(ptr(user_emb), ptr(item_emb), users[s:e].data_ptr(), e - s, I, d, sp, si, k, ids.data_ptr(),
                     sc.data_ptr() if sc is not None else None, ptr(ws), stream()), "topk_exclude")

    mode = mode or TOPK_MODE
    if mode == "auto":
        key = (I, d, k, str(dev))
        if key not in _TOPK_CHOICE and nu >= _TOPK_AUTO_MIN_USERS:
            e = min(block, nu)
            ref_i, ref_s = torch.empty(e, k, dtype=torch.int64, device=dev), torch.empty(e, k, dtype=torch.float32, device=dev)
            got_i, got_s = torch.empty_like(ref_i), torch.empty_like(ref_s)
            run(lib.lgb_topk_exclude, 0, e, ref_i, ref_s)
            try:
                run(lib.lgb_topk_exclude_tiled, 0, e, got_i, got_s)
                same = torch.equal(ref_i, got_i) and torch.equal(ref_s, got_s)
            except RuntimeError:                       # e.g. an item table that is not 16-byte aligned
                same = False
            if same:
                t_cta = _csr._time_ms(lambda: run(lib.lgb_topk_exclude, 0, e, ref_i, ref_s), 2, dev)
                t_tiled = _csr._time_ms(lambda: run(lib.lgb_topk_exclude_tiled, 0, e, got_i, got_s), 2, dev)
                _TOPK_CHOICE[key] = "tiled" if t_tiled < t_cta else "cta"
            else:                                      # never expected: same FMA chains; keep the measured kernel
                _TOPK_CHOICE[key] = "cta"
        mode = _TOPK_CHOICE.get(key, "cta")
    fn = lib.lgb_topk_exclude_tiled if mode == "tiled" else lib.lgb_topk_exclude
    for s in range(0, nu, block):
        e = min(s + block, nu)
        run(fn, s, e, out[s:e], scores[s:e] if scores is not None else None)
        _lib.count_launch(2 if mode == "tiled" else 1)
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
