"""LightGCN mini-batch / negative sampling with the reference's call surface
(data/lightgcn_loader.py:95-112, PyG ``structured_negative_sampling``), membership tests on the GPU.

Bit-exactness contract: every random number is drawn from the SAME generators in the SAME order as the
reference (torch's global CPU generator for ``torch.randint``, Python's ``random`` for
``random.choices``); only the O(E log E) ``np.isin`` rejection test moves to the device
(lgb_neg_reject_mask: binary search in the sorted positive-key array, which is built once per edge
list and cached).  The returned indices are therefore identical to the reference's for the same seeds.
"""
from __future__ import annotations

import ctypes as C
import random
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import check, ptr, stream

_KEY_CACHE: Dict[tuple, torch.Tensor] = {}
_KEY_CACHE_MAX = 8


def _device_for(t: torch.Tensor) -> torch.device:
    if t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("structured_negative_sampling needs a CUDA device (no CPU fallback in this package)")
    return torch.device("cuda", torch.cuda.current_device())


def _sorted_pos_keys(row: torch.Tensor, col: torch.Tensor, num_nodes: int, cache_key: Optional[tuple]) -> torch.Tensor:
    if cache_key is not None and cache_key in _KEY_CACHE:
        return _KEY_CACHE[cache_key]
    lib = _lib.load()
    n = row.numel()
    need = C.c_size_t(0)
    check(lib.lgb_sort_keys_ws_bytes(n, C.byref(need)), "sort_keys_ws_bytes")
    ws = torch.empty(max(need.value, 256), dtype=torch.uint8, device=row.device)
    keys = torch.empty(n, dtype=torch.int64, device=row.device)
    with torch.cuda.device(row.device):
        check(lib.lgb_edge_keys_sorted(ptr(row), ptr(col), n, num_nodes, ptr(keys), ptr(ws), ws.numel(), stream()),
              "edge_keys_sorted")
    _lib.count_launch(2)
    if cache_key is not None:
        if len(_KEY_CACHE) >= _KEY_CACHE_MAX:
            _KEY_CACHE.pop(next(iter(_KEY_CACHE)))
        _KEY_CACHE[cache_key] = keys
    return keys


def _reject_mask(row: torch.Tensor, cand: torch.Tensor, num_nodes: int, keys: torch.Tensor, loops: bool) -> torch.Tensor:
    mask = torch.empty(row.numel(), dtype=torch.uint8, device=row.device)
    with torch.cuda.device(row.device):
        check(_lib.load().lgb_neg_reject_mask(ptr(row), ptr(cand), row.numel(), num_nodes, ptr(keys), keys.numel(),
                                              0 if loops else 1, ptr(mask), stream()), "neg_reject_mask")
    _lib.count_launch()
    return mask


def both_indexes_from_zero(edge_index: torch.Tensor) -> torch.Tensor:
    """Item ids of a to_homogeneous() edge list shifted down by max(user id)+1 (data/lightgcn_loader.py:39-43)."""
    out = torch.clone(edge_index)
    out[1] = out[1] - (torch.max(out[0]) + 1)
    return out


def split(edge_index: torch.Tensor):
    """80/10/10 edge split, sklearn ``train_test_split`` with random_state=1 twice (data/lightgcn_loader.py:13-31):
    the same index permutation as the reference, hence bit-identical splits."""
    state = random.getstate()      # importing sklearn for the first time draws from Python's global `random`: the reference
    from sklearn.model_selection import train_test_split   # imports it at module load, i.e. BEFORE the caller seeds, so the
    random.setstate(state)         # lazy import here must not move the stream that sample_mini_batch's random.choices uses
    n = edge_index.shape[1]
    train_idx, rest = train_test_split(list(range(n)), test_size=0.2, random_state=1)
    val_idx, test_idx = train_test_split(rest, test_size=0.5, random_state=1)
    dev = edge_index.device
    pick = lambda idx: edge_index[:, torch.tensor(idx, dtype=torch.long, device=dev)]   # noqa: E731
    return pick(train_idx), pick(val_idx), pick(test_idx), edge_index


def make_lightgcn_splits(homogeneous_edge_index: torch.Tensor, num_users: int, num_items: int):
    """create_dataloaders_lightgcn (data/lightgcn_loader.py:54-91) from an in-memory edge list instead of the
    data/derived/*.pt files: the same 9-tuple, with the reference's wiring (rows = user ids, cols = item ids in
    [0, I), matrix (U+I)^2)."""
    from .sparse import SparseTensor
    edge_index = both_indexes_from_zero(homogeneous_edge_index)
    train_ei, val_ei, test_ei, edge_index = split(edge_index)
    n = num_users + num_items
    mk = lambda ei: SparseTensor(row=ei[0], col=ei[1], sparse_sizes=(n, n))   # noqa: E731
    return mk(train_ei), mk(val_ei), mk(test_ei), train_ei, val_ei, test_ei, edge_index, num_users, num_items


def structured_negative_sampling(edge_index: torch.Tensor, num_nodes=None, contains_neg_self_loops: bool = True
                                 ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """PyG semantics: for every edge (i, j) a node k with (i, k) not an edge.

    Returns ``(edge_index[0], edge_index[1], neg)`` with ``neg`` on ``edge_index.device``.
    """
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1 if edge_index.numel() else 0
    num_nodes = int(num_nodes)
    dev = _device_for(edge_index)
    E = edge_index.shape[1]
    ei = edge_index if edge_index.is_cuda else edge_index.to(dev, non_blocking=True)
    row, col = _lib.i64c(ei[0]), _lib.i64c(ei[1])
    cache_key = None
    if edge_index.is_cuda and E > 0:
        # the sorted key array of a resident edge list is reused across iterations; the key carries a small
        # content fingerprint so a recycled allocation with other contents cannot alias it
        probe = torch.tensor([0, E // 3, E // 2, (2 * E) // 3, E - 1], device=dev)
        finger = tuple(ei[:, probe].flatten().tolist())
        cache_key = (edge_index.data_ptr(), E, num_nodes, str(edge_index.device), edge_index._version, finger)
    keys = _sorted_pos_keys(row, col, num_nodes, cache_key)

    rand = torch.randint(num_nodes, (E,), dtype=torch.long)            # CPU generator: same draw as the reference
    rand_dev = rand.to(dev, non_blocking=True)
    mask = _reject_mask(row, rand_dev, num_nodes, keys, contains_neg_self_loops)
    rest = mask.nonzero(as_tuple=False).view(-1)
    while rest.numel() > 0:                                             # .numel() synchronises: the redraw count must be exact
        tmp = torch.randint(num_nodes, (rest.numel(),), dtype=torch.long)
        tmp_dev = tmp.to(dev, non_blocking=True)
        rand_dev[rest] = tmp_dev
        sub_mask = _reject_mask(row[rest].contiguous(), tmp_dev, num_nodes, keys, contains_neg_self_loops)
        rest = rest[sub_mask.bool()]
    return edge_index[0], edge_index[1], rand_dev.to(edge_index.device)


def sample_mini_batch(batch_size: int, edge_index: torch.Tensor):
    """Reference data/lightgcn_loader.py:95-112: negatives for ALL edges (num_nodes = max item id, as the
    reference passes it), then ``random.choices`` picks ``batch_size`` triples.  Returns three CPU int64
    tensors like the reference; only the selected triples cross PCIe."""
    num_nodes = int(torch.max(edge_index[1]))
    u_all, p_all, n_all = structured_negative_sampling(edge_index, num_nodes=num_nodes)
    E = edge_index.shape[1]
    indices = random.choices(range(E), k=batch_size)                    # same draws as choices([0..E-1], k)
    idx = torch.tensor(indices, dtype=torch.long).to(edge_index.device)
    batch = torch.stack([u_all[idx], p_all[idx], n_all[idx]], dim=0).to("cpu")
    return batch[0], batch[1], batch[2]


class DeviceSampler:
    """Opt-in, NOT bit-identical to the reference: the same DISTRIBUTION as ``sample_mini_batch`` -- B edges picked uniformly
    with replacement, each with a negative drawn uniformly from [0, max item id) and redrawn while (user, negative) is a
    positive edge -- at O(B) cost per call instead of O(E).  The reference (data/lightgcn_loader.py:95-112) draws a negative
    for EVERY edge on the CPU and then keeps B = 128 of them: at the H&M shape 0.1-0.2 s of host work per iteration against a
    6 ms GPU step.  Here the sorted positive-key array is built once per edge list; a call picks its B edges first and draws
    only their negatives, on the device (torch's CUDA Philox generator), testing membership with the same kernel
    (lgb_neg_reject_mask).  Because it consumes another random stream than the reference it is never selected implicitly."""

    def __init__(self, edge_index: torch.Tensor, generator: Optional[torch.Generator] = None):
        dev = _device_for(edge_index)
        ei = edge_index if edge_index.device == dev else edge_index.to(dev)
        self.E = ei.shape[1]
        self.num_nodes = int(torch.max(ei[1])) if self.E else 0     # the reference's num_nodes = max item id (:105-107)
        self.row, self.col = _lib.i64c(ei[0]).contiguous(), _lib.i64c(ei[1]).contiguous()
        self.keys = _sorted_pos_keys(self.row, self.col, self.num_nodes, None)
        self.generator, self.device = generator, dev

    def sample(self, batch_size: int):
        """-> (users, positives, negatives): three DEVICE int64 tensors."""
        dev, g, hi = self.device, self.generator, max(self.num_nodes, 1)
        pick = torch.randint(self.E, (batch_size,), device=dev, generator=g)
        u, p = self.row[pick], self.col[pick]
        n = torch.randint(hi, (batch_size,), device=dev, generator=g)
        for _ in range(100000):                            # P(redraw) = deg(u) / num_nodes per round: normally a handful of rounds
            mask = _reject_mask(u, n, self.num_nodes, self.keys, True).bool()
            if not bool(mask.any()):
                return u, p, n
            n = torch.where(mask, torch.randint(hi, (batch_size,), device=dev, generator=g), n)
        raise RuntimeError("DeviceSampler: a user of the batch has (almost) every item as a positive -- no negative found "
                           "(the reference's structured_negative_sampling loops forever on such a user)")


def sample_mini_batch_device(batch_size: int, edge_index: torch.Tensor, generator: Optional[torch.Generator] = None):
    """One-shot form of ``DeviceSampler`` (builds the key array on every call: keep a DeviceSampler in a training loop)."""
    return DeviceSampler(edge_index, generator).sample(batch_size)
