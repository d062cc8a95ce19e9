"""laplace_gnn_recommendation_b200 -- B200-native (sm_100a) graph-propagation hot path behind the
call surface of dream-faster/laplace-gnn-recommendation's ``model/lightgcn.py`` and
``model/encoder_decoder.py``.

Only the hot path lives here: the CUDA kernels + C ABI (``csrc/``, ``include/laplace_b200.h``) and the
host-side mirror of the reference interface (``LightGCN``, ``bpr_loss``, ``SparseTensor``, ``matmul``,
``gcn_norm``, ``sample_mini_batch``, ``SAGEConv``, ``to_hetero``, ``Encoder_Decoder_Model`` ...).
There is no CPU fallback: the kernels fail loudly when the shared library or a CUDA device is missing.
"""
from . import _lib  # noqa: F401
from .aliases import HeteroData, install_aliases, patch_driver  # noqa: F401
from .bpr import bpr_indexed, bpr_loss  # noqa: F401
from . import csr, lightgcn  # noqa: F401
from .csr import DeviceCSR, rows_bitmap  # noqa: F401
from .hetero import (EdgeDecoder, Encoder_Decoder_Model, GNNEncoder, HeteroEncoder, SAGEConv, aggregate,  # noqa: F401
                     build_edge_csr, edge_concat, edge_dot, get_linear_layers, get_SAGEConv_layers, to_hetero)
from .lightgcn import LightGCN  # noqa: F401
from .loader import (DeviceSampler, both_indexes_from_zero, make_lightgcn_splits, sample_mini_batch, sample_mini_batch_device, split,  # noqa: F401
                     structured_negative_sampling)
from .metrics import evaluation, get_metrics_lightgcn, get_metrics_universal, recall_precision_ndcg  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .sparse import SparseTensor, gcn_norm, matmul  # noqa: F401
from .subgraph import SubgraphSampler  # noqa: F401
from .topk import SeenItems, make_predictions_for_user, recommend_topk, topk_dict  # noqa: F401

__all__ = [
    "LightGCN", "bpr_loss", "bpr_indexed", "SparseTensor", "matmul", "gcn_norm", "DeviceCSR",
    "sample_mini_batch", "sample_mini_batch_device", "DeviceSampler", "structured_negative_sampling", "recommend_topk", "make_predictions_for_user",
    "SeenItems", "topk_dict", "SAGEConv", "to_hetero", "GNNEncoder", "HeteroEncoder", "EdgeDecoder",
    "Encoder_Decoder_Model", "get_SAGEConv_layers", "get_linear_layers", "aggregate", "build_edge_csr",
    "edge_concat", "edge_dot", "both_indexes_from_zero", "split", "make_lightgcn_splits", "evaluation",
    "get_metrics_lightgcn", "get_metrics_universal", "recall_precision_ndcg", "FusedAdam", "install_aliases", "patch_driver",
    "HeteroData", "SubgraphSampler", "rows_bitmap",
]
