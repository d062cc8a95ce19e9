"""CPU oracle (TEST INFRASTRUCTURE, not product) for the LightGCN loader / sampler.

* ``both_indexes_from_zero``  data/lightgcn_loader.py:39-43
* ``split``                   data/lightgcn_loader.py:13-31 (sklearn ``train_test_split``, seed 1)
* ``structured_negative_sampling``  PyG utility called at data/lightgcn_loader.py:105-107 and
  run_pipeline_lightgcn.py:40-44; source absent from /root/reference -> restated from its
  published algorithm: **parity unpinned** for this function.
* ``sample_mini_batch``       data/lightgcn_loader.py:95-112

All randomness comes from torch's global CPU generator and Python's ``random`` in the same
order as the reference, so seeded runs are reproducible draw for draw.
"""
from __future__ import annotations

import random
from typing import Tuple

import numpy as np
import torch

Tensor = torch.Tensor


def both_indexes_from_zero(edge_index: Tensor) -> Tensor:
    """Item ids shifted down by ``max(user id)+1`` (data/lightgcn_loader.py:39-43)."""
    out = edge_index.clone()
    out[1] = out[1] - (out[0].max() + 1)
    return out


def split(edge_index: Tensor):
    """80/10/10 edge split with sklearn, random_state=1 twice (data/lightgcn_loader.py:13-31)."""
    from sklearn.model_selection import train_test_split

    idx = list(range(edge_index.shape[1]))
    train_idx, rest = train_test_split(idx, test_size=0.2, random_state=1)
    val_idx, test_idx = train_test_split(rest, test_size=0.5, random_state=1)
    return edge_index[:, train_idx], edge_index[:, val_idx], edge_index[:, test_idx], edge_index


def structured_negative_sampling(edge_index: Tensor, num_nodes=None,
                                 contains_neg_self_loops: bool = True
                                 ) -> Tuple[Tensor, Tensor, Tensor]:
    """For every edge (i, j) draw k with (i, k) not an edge (A6).

    pos = row*num_nodes + col (+ the self-loop keys i*(num_nodes+1) when
    ``contains_neg_self_loops`` is False); rand = randint(num_nodes, (E,)); reject with
    ``np.isin`` and redraw ONLY the rejected positions until none is left.
    ``num_nodes`` may be a 0-dim tensor, as the reference passes ``t.max(edge_index[1])``.
    """
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1 if edge_index.numel() else 0
    row, col = edge_index.cpu()
    pos_idx = row * num_nodes + col
    if not contains_neg_self_loops:
        loop_idx = torch.arange(num_nodes) * (num_nodes + 1)
        pos_idx = torch.cat([pos_idx, loop_idx], dim=0)

    rand = torch.randint(num_nodes, (row.size(0),), dtype=torch.long)
    neg_idx = row * num_nodes + rand
    mask = torch.from_numpy(np.isin(neg_idx.numpy(), pos_idx.numpy())).to(torch.bool)
    rest = mask.nonzero(as_tuple=False).view(-1)
    while rest.numel() > 0:
        tmp = torch.randint(num_nodes, (rest.size(0),), dtype=torch.long)
        rand[rest] = tmp
        neg_idx = row[rest] * num_nodes + tmp
        mask = torch.from_numpy(np.isin(neg_idx.numpy(), pos_idx.numpy())).to(torch.bool)
        rest = rest[mask]
    return edge_index[0], edge_index[1], rand.to(edge_index.device)


def sample_mini_batch(batch_size: int, edge_index: Tensor):
    """Negatives for ALL edges, then ``random.choices`` picks the batch (data/lightgcn_loader.py:95-112).

    Note the reference's ``num_nodes = max(item id)`` (not +1): the top item is never a negative.
    """
    edges = structured_negative_sampling(edge_index.to("cpu"), num_nodes=torch.max(edge_index[1]).to("cpu"))
    edges = torch.stack(edges, dim=0)
    indices = random.choices([i for i in range(edges[0].shape[0])], k=batch_size)
    batch = edges[:, indices]
    return batch[0], batch[1], batch[2]
