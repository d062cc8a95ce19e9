"""Import shims that let the REAL reference modules under /root/reference be imported in the build
container, where torch_sparse / torch_geometric are not installed (TEST INFRASTRUCTURE, not product).

Used only by tests/golden/make_golden.py to generate the committed golden fixtures.  The third-party
names are backed by the oracle's restatement (so those pieces stay "parity unpinned"), while every line of
the reference's own code (model/lightgcn.py, utils/metrics_lightgcn.py, utils/metrics.py, utils/tensor.py,
data/lightgcn_loader.py, model/encoder_decoder.py, model/layers.py) runs unmodified on top of them.
"""
from __future__ import annotations

import copy
import sys
import types

import torch
from torch import nn

from . import hetero_oracle as ho
from . import lightgcn_oracle as lo
from . import sampler_oracle as so


class SparseTensor:
    """CPU stand-in for torch_sparse.SparseTensor (CSR + optional values), oracle A1."""

    def __init__(self, row=None, col=None, value=None, sparse_sizes=None, _csr=None):
        if _csr is not None:
            self.rowptr, self.col, self.value, self._sizes = _csr
            return
        m, n = sparse_sizes
        self.rowptr, self.col, perm = lo.csr_from_coo(row, col, m, n)
        self.value = None if value is None else value[perm]
        self._sizes = (m, n)

    def to(self, *a, **k):
        return self

    def sparse_sizes(self):
        return self._sizes

    def has_value(self):
        return self.value is not None

    def with_value(self, value):
        return SparseTensor(_csr=(self.rowptr, self.col, value, self._sizes))


def matmul(adj: SparseTensor, x: torch.Tensor) -> torch.Tensor:
    return lo.spmm(adj.rowptr, adj.col, adj.value, x)


def gcn_norm(adj: SparseTensor, edge_weight=None, num_nodes=None, improved=False, add_self_loops=True,
             dtype=None):
    assert not add_self_loops, "shim covers the reference's add_self_loops=False call only"
    _, val = lo.gcn_norm_values(adj.rowptr, adj.col)
    return adj.with_value(val)


class MessagePassing(nn.Module):
    def __init__(self, aggr="add", **kwargs):
        super().__init__()
        self.aggr = aggr

    def propagate(self, edge_index, size=None, **kwargs):
        assert isinstance(edge_index, SparseTensor)
        return self.message_and_aggregate(edge_index, kwargs["x"])


class SAGEConv(nn.Module):
    """PyG SAGEConv stand-in (lazy in-channels, lin_l with bias, lin_r without), oracle A8."""

    def __init__(self, in_channels, out_channels, aggr="mean", normalize=False, root_weight=True, bias=True):
        super().__init__()
        self.in_channels, self.out_channels, self.aggr = in_channels, out_channels, aggr
        self.lin_l = nn.LazyLinear(out_channels, bias=bias)
        self.lin_r = nn.LazyLinear(out_channels, bias=False)

    def forward(self, x, edge_index):
        x_src, x_dst = (x, x) if isinstance(x, torch.Tensor) else x
        agg = ho.scatter_aggregate(x_src, edge_index, x_dst.shape[0], self.aggr)
        return self.lin_l(agg) + self.lin_r(x_dst)


class _HeteroEncoder(nn.Module):
    """Stand-in for to_hetero(GNNEncoder): the layer loop of model/encoder_decoder.py:29-46 lifted over
    edge types with the fan-in of temporary_hetero.py:201-228.  Module names follow PyG's
    ``layers.<i>.<src>__<rel>__<dst>`` convention."""

    def __init__(self, module, metadata, aggr):
        super().__init__()
        self.node_types, self.edge_types = metadata
        self.aggr = aggr
        self.p_dropout_features = module.p_dropout_features
        self.layers = nn.ModuleList([
            nn.ModuleDict({"__".join(et): copy.deepcopy(layer) for et in self.edge_types})
            for layer in module.layers])

    def forward(self, x_dict, edge_index_dict):
        x = dict(x_dict)
        for li, convs in enumerate(self.layers):
            last = li == len(self.layers) - 1
            if not last and self.p_dropout_features is not None:
                # FX tracing bakes training=True
                x = {k: torch.nn.functional.dropout(v, p=self.p_dropout_features, training=True)
                     for k, v in x.items()}
            per_dst = {}
            for et in self.edge_types:
                s, _, d = et
                per_dst.setdefault(d, []).append(convs["__".join(et)]((x[s], x[d]), edge_index_dict[et]))
            x = {d: ho.hetero_fan_in(v, self.aggr) for d, v in per_dst.items()}
            if not last:
                x = {k: v.relu() for k, v in x.items()}
        return x


def to_hetero(module, metadata, aggr="sum"):
    return _HeteroEncoder(module, metadata, aggr)


class _Placeholder:
    def __init__(self, *a, **k):
        pass


def install():
    """Register the stand-ins under the third-party module names (idempotent)."""
    def mod(name, **attrs):
        m = sys.modules.get(name) or types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    mod("torch_sparse", SparseTensor=SparseTensor, matmul=matmul)
    tg = mod("torch_geometric", seed_everything=lambda s: (torch.manual_seed(s), __import__("random").seed(s),
                                                          __import__("numpy").random.seed(s)))
    tg.nn = mod("torch_geometric.nn", SAGEConv=SAGEConv, to_hetero=to_hetero, MessagePassing=MessagePassing)
    tg.nn.conv = mod("torch_geometric.nn.conv", MessagePassing=MessagePassing)
    tg.nn.conv.gcn_conv = mod("torch_geometric.nn.conv.gcn_conv", gcn_norm=gcn_norm)
    tg.data = mod("torch_geometric.data", Data=_Placeholder, HeteroData=_Placeholder)
    tg.utils = mod("torch_geometric.utils", structured_negative_sampling=so.structured_negative_sampling,
                   dropout_adj=None)
    tg.loader = mod("torch_geometric.loader", DataLoader=_Placeholder)
