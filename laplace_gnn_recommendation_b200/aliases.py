"""``install_aliases()``: register this package's classes under the third-party module names the reference imports
(``torch_sparse``, ``torch_geometric.nn / .nn.conv / .nn.conv.gcn_conv / .utils / .data / .loader``), so that the
reference's OWN modules and drivers -- ``model/lightgcn.py``, ``model/encoder_decoder.py``, ``data/lightgcn_loader.py``,
``run_pipeline_lightgcn.py``, ``training.py`` -- import and run UNMODIFIED on the sm_100a kernels (SURVEY.md section 7
step 0: the reference's hot path lives entirely in torch_sparse / PyG, which are not installable here).

Nothing in this file computes: every name maps to a class or function of this package (never to the oracle).  What the
reference imports but never calls on the hot path (``Data``, the loaders, ``dropout_adj``, ``transforms``) is a named
placeholder that raises when used.  ``HeteroData`` is the small typed container the ranking path needs from PyG
(``x_dict``, ``edge_index_dict``, ``metadata()``, ``data[edge_type].edge_label_index``, ``.to(device)``).

``patch_driver(module)`` is the other drop-in route north_star describes: the reference driver keeps its own code and
only the names it imported from ``model.lightgcn`` / ``utils.metrics_lightgcn`` / ``data.lightgcn_loader`` are swapped for
this package's implementations (fused kernels instead of the reference's composition of library calls).
"""
from __future__ import annotations

import importlib.util
import random
import sys
import types
from typing import Dict

import torch
from torch import nn


class MessagePassing(nn.Module):
    """The slice of PyG's MessagePassing that ``model/lightgcn.py:11-87`` uses: ``propagate`` with a SparseTensor
    adjacency dispatches to ``message_and_aggregate`` (PyG's fused path)."""

    def __init__(self, aggr="add", **kwargs):
        super().__init__()
        self.aggr = aggr

    def propagate(self, edge_index, size=None, **kwargs):
        from .sparse import SparseTensor
        if not isinstance(edge_index, SparseTensor):
            raise RuntimeError("MessagePassing.propagate: only the SparseTensor form the reference uses is supported")
        return self.message_and_aggregate(edge_index, kwargs["x"])


class _Store:
    """Attribute bag of one node / edge type."""

    def __init__(self):
        object.__setattr__(self, "_d", {})

    def __getattr__(self, name):
        try:
            return object.__getattribute__(self, "_d")[name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name, value):
        self._d[name] = value

    def __contains__(self, name):
        return name in self._d


class HeteroData:
    """Minimal typed graph container with PyG's HeteroData surface used by ``training.py`` / ``utils/get_info.py``:
    ``data['customer'].x``, ``data['customer', 'buys', 'article'].edge_index`` / ``.edge_label_index`` / ``.edge_label``,
    ``x_dict``, ``edge_index_dict``, ``metadata()``, ``to(device)``."""

    def __init__(self):
        self._stores: Dict[object, _Store] = {}

    def __getitem__(self, key):
        key = tuple(key) if isinstance(key, (tuple, list)) else key
        if key not in self._stores:
            self._stores[key] = _Store()
        return self._stores[key]

    @property
    def node_types(self):
        return [k for k in self._stores if isinstance(k, str)]

    @property
    def edge_types(self):
        return [k for k in self._stores if isinstance(k, tuple)]

    def metadata(self):
        return self.node_types, self.edge_types

    @property
    def x_dict(self):
        return {k: self._stores[k].x for k in self.node_types if "x" in self._stores[k]}

    @property
    def edge_index_dict(self):
        return {k: self._stores[k].edge_index for k in self.edge_types if "edge_index" in self._stores[k]}

    def to(self, device, *args, **kwargs):
        out = HeteroData()
        for k, st in self._stores.items():
            for name, v in st._d.items():
                setattr(out[k], name, v.to(device, *args, **kwargs) if torch.is_tensor(v) else v)
        return out


def seed_everything(seed: int) -> None:
    import numpy as np
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def _placeholder(name: str):
    class _P:
        def __init__(self, *a, **k):
            raise RuntimeError(f"{name} is a named placeholder of laplace_gnn_recommendation_b200.install_aliases(): the "
                               "reference imports it, the propagation hot path never uses it")
    _P.__name__ = _P.__qualname__ = name.rsplit(".", 1)[-1]
    return _P


def _third_party_present() -> bool:
    try:
        return importlib.util.find_spec("torch_sparse") is not None and importlib.util.find_spec("torch_geometric") is not None
    except (ImportError, ValueError):
        return False


def install_aliases(force: bool = False) -> bool:
    """Register the aliases (idempotent).  A real torch_sparse / torch_geometric installation is left alone unless
    ``force`` -- then the reference's modules run on this package instead.  Returns True when the aliases are active."""
    if _third_party_present() and not force and not getattr(sys.modules.get("torch_sparse"), "__lgb_alias__", False):
        return False
    from . import hetero, loader, sparse

    def mod(name, **attrs):
        m = sys.modules.get(name)
        if m is None or not getattr(m, "__lgb_alias__", False):
            m = types.ModuleType(name)
            m.__lgb_alias__ = True
            m.__doc__ = "alias of laplace_gnn_recommendation_b200 (install_aliases)"
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    mod("torch_sparse", SparseTensor=sparse.SparseTensor, matmul=sparse.matmul)
    tg = mod("torch_geometric", seed_everything=seed_everything)
    tg.nn = mod("torch_geometric.nn", SAGEConv=hetero.SAGEConv, to_hetero=hetero.to_hetero, MessagePassing=MessagePassing)
    tg.nn.conv = mod("torch_geometric.nn.conv", MessagePassing=MessagePassing, SAGEConv=hetero.SAGEConv)
    tg.nn.conv.gcn_conv = mod("torch_geometric.nn.conv.gcn_conv", gcn_norm=sparse.gcn_norm)
    tg.data = mod("torch_geometric.data", HeteroData=HeteroData, Data=_placeholder("torch_geometric.data.Data"))
    tg.utils = mod("torch_geometric.utils", structured_negative_sampling=loader.structured_negative_sampling,
                   dropout_adj=_placeholder("torch_geometric.utils.dropout_adj"))
    tg.loader = mod("torch_geometric.loader", **{n: _placeholder(f"torch_geometric.loader.{n}")
                                                 for n in ("DataLoader", "NeighborLoader", "LinkNeighborLoader")})
    tg.transforms = mod("torch_geometric.transforms", ToUndirected=_placeholder("torch_geometric.transforms.ToUndirected"))
    return True


def patch_driver(module) -> list:
    """Swap the names a reference driver module imported from its own model / loss / sampler modules for this package's
    drop-ins (same signatures): LightGCN, bpr_loss, sample_mini_batch, structured_negative_sampling,
    make_predictions_for_user, Encoder_Decoder_Model, get_SAGEConv_layers, get_linear_layers, get_metrics_universal.
    Returns the names that were replaced."""
    import laplace_gnn_recommendation_b200 as lg
    swapped = []
    for name in ("LightGCN", "bpr_loss", "sample_mini_batch", "structured_negative_sampling", "make_predictions_for_user",
                 "Encoder_Decoder_Model", "get_SAGEConv_layers", "get_linear_layers", "get_metrics_universal"):
        if hasattr(module, name) and hasattr(lg, name):
            setattr(module, name, getattr(lg, name))
            swapped.append(name)
    return swapped
