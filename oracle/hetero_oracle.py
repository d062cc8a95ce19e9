"""CPU oracle (TEST INFRASTRUCTURE, not product) for the ranking encoder-decoder hot path.

* ``SAGEConv((-1,-1,-1), C, aggr, normalize=False, bias=True)`` as built by model/layers.py:6-32 --
  third-party PyG, restated from its published algorithm: **parity unpinned**.
* ``to_hetero(GNNEncoder, metadata, aggr)`` (model/encoder_decoder.py:29-46,93-95) with the
  per-destination pairwise fan-in of temporary_hetero.py:201-228 (the only PyG source in the tree).
* ``EdgeDecoder.forward`` concat -> MLP (model/encoder_decoder.py:55-72) and the literal
  dot-product decoder that BASELINE.json's north_star names.
* per-type ``BatchNorm1d`` (model/encoder_decoder.py:98-99,144-150), ``BCEWithLogitsLoss`` (training.py:26-31),
  ``infer`` re-batching by user (model/encoder_decoder.py:155-164 + utils/tensor.py:24-61).

Everything is functional (weights passed in) so the same tensors can be handed to the CUDA path.
"""
from __future__ import annotations

from collections import defaultdict, deque
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
EdgeType = Tuple[str, str, str]


def scatter_aggregate(x_src: Tensor, edge_index: Tensor, n_dst: int, aggr: str) -> Tensor:
    """agg[t] = (+|mean|max)_{e: dst[e]=t} x_src[src[e]]  (A8).  Empty segments give 0; mean divides by
    max(count, 1).  Messages are materialised and scattered in edge order like PyG's propagate."""
    src, dst = edge_index[0], edge_index[1]
    msg = x_src.index_select(0, src)
    out = torch.zeros(n_dst, x_src.shape[1], dtype=x_src.dtype)
    if aggr in ("add", "sum"):
        return out.index_add_(0, dst, msg)
    if aggr == "mean":
        out = out.index_add_(0, dst, msg)
        cnt = torch.zeros(n_dst, dtype=x_src.dtype).index_add_(0, dst, torch.ones_like(dst, dtype=x_src.dtype))
        return out / cnt.clamp(min=1).unsqueeze(1)
    if aggr == "max":
        idx = dst.unsqueeze(1).expand_as(msg)
        return out.scatter_reduce(0, idx, msg, reduce="amax", include_self=False)
    raise ValueError(aggr)


def sage_conv(x_src: Tensor, x_dst: Tensor, edge_index: Tensor, w_l: Tensor, b_l: Optional[Tensor],
              w_r: Tensor, aggr: str) -> Tensor:
    """out = lin_l(agg) + lin_r(x_dst); lin_l carries the bias, lin_r has none; no L2 normalise."""
    agg = scatter_aggregate(x_src, edge_index, x_dst.shape[0], aggr)
    return F.linear(agg, w_l, b_l) + F.linear(x_dst, w_r)


_PAIR_OPS = {"sum": torch.add, "mean": torch.add, "max": torch.max, "min": torch.min, "mul": torch.mul}


def hetero_fan_in(outs: List[Tensor], aggr: str) -> Tensor:
    """Pairwise queue reduction of temporary_hetero.py:201-228: pop two, combine, push to the BACK."""
    n = len(outs)
    queue = deque(outs)
    while len(queue) >= 2:
        a, b = queue.popleft(), queue.popleft()
        queue.append(_PAIR_OPS[aggr](a, b))
    res = queue.popleft()
    if aggr == "mean" and n > 1:
        res = torch.div(res, n)
    return res


def hetero_encoder(x_dict: Dict[str, Tensor], edge_index_dict: Dict[EdgeType, Tensor],
                   layers: List[Dict[EdgeType, Dict[str, Tensor]]], conv_aggr: str, hetero_aggr: str,
                   edge_types: Optional[List[EdgeType]] = None) -> Dict[str, Tensor]:
    """GNNEncoder.forward lifted over node/edge types (dropout off): non-last layers conv -> relu, last conv.

    Node types that are the destination of no edge type drop out of the dict after a layer, exactly as in
    the traced module (every node type of the default metadata is a destination)."""
    edge_types = edge_types or list(edge_index_dict.keys())
    x = dict(x_dict)
    for li, params in enumerate(layers):
        per_dst: Dict[str, List[Tensor]] = defaultdict(list)
        for et in edge_types:
            s, _, d = et
            p = params[et]
            per_dst[d].append(sage_conv(x[s], x[d], edge_index_dict[et], p["w_l"], p["b_l"], p["w_r"], conv_aggr))
        x = {d: hetero_fan_in(v, hetero_aggr) for d, v in per_dst.items()}
        if li != len(layers) - 1:
            x = {k: v.relu() for k, v in x.items()}
    return x


def edge_decoder_mlp(z_user: Tensor, z_item: Tensor, edge_label_index: Tensor,
                     linears: List[Tuple[Tensor, Tensor]]) -> Tensor:
    """cat(z_user[row], z_item[col]) -> (Linear -> relu)* -> Linear -> view(-1)   (dropout off)."""
    row, col = edge_label_index[0], edge_label_index[1]
    z = torch.cat([z_user[row], z_item[col]], dim=-1)
    for i, (w, b) in enumerate(linears):
        z = F.linear(z, w, b)
        if i != len(linears) - 1:
            z = z.relu()
    return z.view(-1)


def edge_decoder_dot(z_user: Tensor, z_item: Tensor, edge_label_index: Tensor) -> Tensor:
    """score_e = <z_user[row_e], z_item[col_e]> -- the decoder north_star names (no reference counterpart)."""
    return (z_user[edge_label_index[0]] * z_item[edge_label_index[1]]).sum(dim=-1)


def batch_norm_train(z: Tensor, weight: Tensor, bias: Tensor, eps: float = 1e-5) -> Tensor:
    return F.batch_norm(z, None, None, weight, bias, training=True, eps=eps)


def bce_with_logits(logits: Tensor, labels: Tensor) -> Tensor:
    return F.binary_cross_entropy_with_logits(logits, labels.float())


def infer_rebatch(out: Tensor, edge_label_index: Tensor, value=-(1 << 50)) -> Tensor:
    """Scores regrouped per user and right-padded with -2^50 (model/encoder_decoder.py:161-164)."""
    users = edge_label_index[0].unique(sorted=True)
    per_user = [out[edge_label_index[0] == u] for u in users]
    width = max(x.size(-1) for x in per_user)
    return torch.stack([F.pad(x, (0, width - x.size(-1)), value=value) if width > x.size(-1) else x
                        for x in per_user], dim=0)
