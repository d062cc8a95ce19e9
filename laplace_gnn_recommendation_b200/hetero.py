"""Ranking encoder-decoder with the reference's call surface (model/encoder_decoder.py:17-164,
model/layers.py:6-56) on the sm_100a kernels.

What runs where
  * per edge type neighbour aggregation (PyG SAGEConv -> propagate -> gather + torch_scatter): one per-batch
    CSR build (lgb_csr_build, dst-major, stable => deterministic) + the fused segment-reduce kernel
    lgb_spmm (sum / mean) or lgb_segment_max; backward = the same kernel on the transposed arrays;
  * edge decoder: gather+concat kernel (reference-faithful concat -> MLP) or the fused dot-product
    decoder north_star names, both with hand-written backward;
  * the dense pieces (SAGE lin_l / lin_r, decoder Linear layers, BatchNorm1d, feature Embedding) stay
    torch / cuBLAS library calls -- they are not gather/scatter work.

Reference quirks kept on purpose: the feature-embedding ModuleLists live in a plain dict
(not in state_dict(), not in parameters()) -- they only follow the module across devices here so that the
model can run on the GPU at all; encoder dropout is applied with training=True even in eval mode (what FX
tracing bakes in) unless ``bake_dropout_training=False``.
"""
from __future__ import annotations

import copy
import os
import ctypes as C
from collections import defaultdict, deque
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn import BatchNorm1d, Embedding, Linear, ModuleList

from . import _lib
from ._lib import check, ptr, stream
from .csr import DeviceCSR

NODE_USER, NODE_ITEM = "customer", "article"                       # utils/constants.py:4-5
EDGE_KEY = (NODE_USER, "buys", NODE_ITEM)                          # utils/constants.py:14
REV_EDGE_KEY = (NODE_ITEM, "rev_buys", NODE_USER)                  # utils/constants.py:15


def key2str(key) -> str:
    return "__".join(key) if isinstance(key, tuple) else key


# --------------------------------------------------------------------------------------------
# neighbour aggregation
# --------------------------------------------------------------------------------------------
# Per-batch graphs are built WITHOUT a host synchronisation (the batch set-up used to stall the stream five times per step):
#   * no long-row split plan (chunk = 0; its slice count would have to be read back) -- neighbour lists of sampled batches are
#     bounded by the loader's fan-out, a row is walked by one warp / lane group;
#   * the index-range check of the CSR build and the "is this edge type the flip of that one" test leave their verdicts on the
#     device; they are copied to pinned memory asynchronously and examined at the START OF THE NEXT BATCH (or by
#     ``flush_deferred_checks()``), when they have long been computed.  Out-of-range indices are clamped on the device, so a
#     bad batch cannot write out of bounds in the meantime; it raises one step late.
BATCH_GRAPH_CHUNK = int(os.environ.get("LGB_HETERO_CHUNK", "0"))
_deferred: List[tuple] = []          # (event, pinned int32 slot, message)
_slots: Dict[bool, torch.Tensor] = {}    # one pinned ring per process (a cudaHostAlloc per check would serialise with the copies in flight)
_RING = 256


def _defer_check(flag: torch.Tensor, message: str) -> None:
    """flag: device int32 scalar-like tensor that must be 0."""
    ring = _slots.get(flag.is_cuda)
    if ring is None:
        ring = torch.zeros(_RING, dtype=torch.int32)
        ring = ring.pin_memory() if flag.is_cuda else ring
        _slots[flag.is_cuda] = ring
        _slots["next"] = 0
    if len(_deferred) >= _RING - 1:      # nobody flushed for a long time: do it now rather than overwrite a pending slot
        flush_deferred_checks()
    i = _slots["next"] = (_slots.get("next", 0) + 1) % _RING
    host = ring[i:i + 1]
    host.copy_(flag.reshape(-1)[:1], non_blocking=True)
    ev = None
    if flag.is_cuda:
        ev = torch.cuda.Event()
        ev.record()
    _deferred.append((ev, host, message))


def flush_deferred_checks() -> None:
    """Examine the verdicts the previous batches left behind (waits for their copies only, not for the stream)."""
    while _deferred:
        ev, host, message = _deferred.pop(0)
        if ev is not None:
            ev.synchronize()
        if int(host[0]) != 0:
            _deferred.clear()
            raise RuntimeError(message)


def build_edge_csr(edge_index: torch.Tensor, n_src: int, n_dst: int, deferred: bool = False) -> DeviceCSR:
    """dst-major CSR of one edge type: row = edge_index[1] (target), col = edge_index[0] (source).
    deferred=True (the per-batch path): no split plan and the index-range verdict is examined one batch later."""
    _lib.require_cuda(edge_index)
    if not deferred:
        return DeviceCSR.from_coo(edge_index[1], edge_index[0], n_dst, n_src)
    g = DeviceCSR.from_coo(edge_index[1], edge_index[0], n_dst, n_src, chunk=BATCH_GRAPH_CHUNK, validate=False)
    ws = getattr(g, "_build_ws", None)
    if ws is not None:
        _defer_check(ws[:4].view(torch.int32), "an edge_index of the previous batch holds node ids outside its node-feature "
                                               "matrices (lgb_csr_build: index outside sparse_sizes)")
        g._build_ws = None
    return g


class _Aggregate(torch.autograd.Function):
    """agg[t] = (+|mean|max)_{e: dst[e]=t} x_src[src[e]]  (SURVEY.md A8)."""

    @staticmethod
    def forward(ctx, x_src, g: DeviceCSR, aggr: str):
        x_src = _lib.f32c(x_src)
        ctx.g, ctx.aggr = g, aggr
        if aggr in ("add", "sum"):
            return g.spmm(x_src)
        if aggr == "mean":
            return g.spmm(x_src, mean=True)
        if aggr == "max":
            d = x_src.shape[1]
            out = torch.empty(g.n_rows, d, dtype=torch.float32, device=x_src.device)
            arg = torch.empty(g.n_rows, d, dtype=torch.int32, device=x_src.device)
            with torch.cuda.device(x_src.device):
                check(_lib.load().lgb_segment_max(C.byref(g.struct), ptr(x_src), d, ptr(out), ptr(arg), stream()), "segment_max")
            _lib.count_launch()
            ctx.arg = arg
            return out
        raise RuntimeError(f"SAGEConv aggr={aggr!r} is not supported (add, mean, max)")

    @staticmethod
    def backward(ctx, gout):
        g: DeviceCSR = ctx.g
        gout = _lib.f32c(gout)
        lib = _lib.load()
        d = gout.shape[1]
        if ctx.aggr == "max":
            gx = torch.empty(g.n_cols, d, dtype=torch.float32, device=gout.device)
            with torch.cuda.device(gout.device):
                check(lib.lgb_zero(ptr(gx), gx.numel() * 4, stream()), "zero")
                check(lib.lgb_segment_max_bwd(ptr(gout), ptr(ctx.arg), g.n_rows, d, ptr(gx), stream()), "segment_max_bwd")
            _lib.count_launch()
            return gx, None, None
        if ctx.aggr == "mean":
            scaled = torch.empty_like(gout)
            with torch.cuda.device(gout.device):
                check(lib.lgb_row_div_by_degree(ptr(gout), ptr(g.rowptr), g.n_rows, d, ptr(scaled), stream()),
                      "row_div_by_degree")
            _lib.count_launch()
            gout = scaled
        return g.transpose().spmm(gout), None, None


def aggregate(x_src: torch.Tensor, g: DeviceCSR, aggr: str) -> torch.Tensor:
    return _Aggregate.apply(x_src, g, aggr)


# --------------------------------------------------------------------------------------------
# Linear layers with the split-K weight / bias gradient kernel
# --------------------------------------------------------------------------------------------
# rows from which the split-K kernel takes over: below it the step is bound by host-side dispatch, not by the 66 us cuBLAS call,
# and one more Python-level autograd node costs more than the kernel saves (measured on the S / M batches, profiles/README.md)
WGRAD_MIN_ROWS = int(os.environ.get("LGB_WGRAD_MIN_ROWS", "65536"))


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b with torch's forward (cuBLAS) and input gradient, and lgb_linear_wgrad for dW / db: the [out x N] x [N x in]
    contraction over the 10^4..10^5 node rows of a batch, for which cuBLAS runs four CTAs (csrc/dense.cu)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return F.linear(x, weight, bias)

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        gy, x = _lib.f32c(gy), _lib.f32c(x)
        gx = gy @ weight if ctx.needs_input_grad[0] else None
        gw = gb = None
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            N, n_in = x.shape
            n_out = gy.shape[1]
            lib = _lib.load()
            need = C.c_size_t(0)
            check(lib.lgb_linear_wgrad_ws_bytes(N, n_in, n_out, C.byref(need)), "linear_wgrad_ws_bytes")
            ws = torch.empty(need.value, dtype=torch.uint8, device=x.device)
            gw = torch.empty(n_out, n_in, dtype=torch.float32, device=x.device)
            gb = torch.empty(n_out, dtype=torch.float32, device=x.device) if ctx.has_bias else None
            with torch.cuda.device(x.device):
                check(lib.lgb_linear_wgrad(ptr(x), ptr(gy), N, n_in, n_out, ptr(gw), ptr(gb), ptr(ws), ws.numel(), stream()), "linear_wgrad")
            _lib.count_launch(3 if gb is not None else 2)
        return gx, gw, gb


def linear(module: nn.Module, x: torch.Tensor) -> torch.Tensor:
    """module(x) for an nn.Linear (or a materialised LazyLinear) whose backward uses the split-K kernel when the batch has
    enough rows; anything else (lazy parameters, other dtypes, tiny batches) goes through the module itself."""
    lazy = isinstance(module, nn.LazyLinear) and module.has_uninitialized_params()
    if (not lazy and isinstance(module, nn.Linear) and x.dim() == 2 and x.dtype == torch.float32 and x.shape[0] >= WGRAD_MIN_ROWS
            and _lib.on_device(x) and torch.is_grad_enabled() and module.weight.requires_grad):
        return _LinearFn.apply(x, module.weight, module.bias)
    return module(x)


class _SageFn(torch.autograd.Function):
    """One SAGEConv (add / mean) as ONE autograd node: out = (aggr_j x_j) Wl^T + bl + x_dst Wr^T.  The S / M batches of the
    ranking step are bound by host dispatch (profiles/README.md r2c): the composition aggregate -> Linear -> Linear -> add
    costs seven autograd nodes and a dozen op dispatches per layer and edge type, this one costs one node, one library call
    and two GEMM calls forward; backward: the transposed aggregation, three GEMM calls and the two weight gradients (split-K
    kernel from WGRAD_MIN_ROWS rows, cuBLAS below).  Same arithmetic as the composition (the second GEMM accumulates onto the
    first one's fp32 result exactly like the separate add)."""

    @staticmethod
    def forward(ctx, x_src, x_dst, w_l, b_l, w_r, g: DeviceCSR, mean: bool):
        x_src = _lib.f32c(x_src)
        agg = g.spmm(x_src, mean=mean)
        out = torch.addmm(b_l, agg, w_l.t()) if b_l is not None else agg @ w_l.t()
        if w_r is not None:
            out.addmm_(x_dst, w_r.t())
        ctx.save_for_backward(agg, x_dst if w_r is not None else None, w_l, w_r)
        ctx.g, ctx.mean, ctx.has_bias = g, mean, b_l is not None
        return out

    @staticmethod
    def backward(ctx, gout):
        agg, x_dst, w_l, w_r = ctx.saved_tensors
        g: DeviceCSR = ctx.g
        gout = _lib.f32c(gout)
        need = ctx.needs_input_grad
        gx_src = gx_dst = gw_l = gb_l = gw_r = None
        if need[0]:
            d_agg = gout @ w_l
            if ctx.mean:
                scaled = torch.empty_like(d_agg)
                with torch.cuda.device(gout.device):
                    check(_lib.load().lgb_row_div_by_degree(ptr(d_agg), ptr(g.rowptr), g.n_rows, d_agg.shape[1], ptr(scaled), stream()),
                          "row_div_by_degree")
                _lib.count_launch()
                d_agg = scaled
            gx_src = g.transpose().spmm(d_agg)
        if need[2] or (ctx.has_bias and need[3]):
            gw_l, gb_l = _wgrad(agg, gout, ctx.has_bias)
        if w_r is not None:
            if need[1]:
                gx_dst = gout @ w_r
            if need[4]:
                gw_r, _ = _wgrad(x_dst, gout, False)
        return gx_src, gx_dst, gw_l, gb_l, gw_r, None, None


def _wgrad(x: torch.Tensor, gy: torch.Tensor, want_bias: bool):
    """(dW [out, in], db [out] or None) of y = x W^T + b: split-K kernel for many rows, cuBLAS + a column sum otherwise."""
    x = _lib.f32c(x)
    N, n_in = x.shape
    n_out = gy.shape[1]
    if N < WGRAD_MIN_ROWS:
        return gy.t() @ x, (gy.sum(0) if want_bias else None)
    lib = _lib.load()
    need = C.c_size_t(0)
    check(lib.lgb_linear_wgrad_ws_bytes(N, n_in, n_out, C.byref(need)), "linear_wgrad_ws_bytes")
    ws = torch.empty(need.value, dtype=torch.uint8, device=x.device)
    gw = torch.empty(n_out, n_in, dtype=torch.float32, device=x.device)
    gb = torch.empty(n_out, dtype=torch.float32, device=x.device) if want_bias else None
    with torch.cuda.device(x.device):
        check(lib.lgb_linear_wgrad(ptr(x), ptr(gy), N, n_in, n_out, ptr(gw), ptr(gb), ptr(ws), ws.numel(), stream()), "linear_wgrad")
    _lib.count_launch(3 if gb is not None else 2)
    return gw, gb


FUSED_SAGE = os.environ.get("LGB_SAGE_FUSED", "1") == "1"


class SAGEConv(nn.Module):
    """PyG ``SAGEConv(in_channels, out_channels, aggr, normalize=False, bias=True)`` as the reference builds it
    (model/layers.py:9-24): out = lin_l(aggr_j x_j) + lin_r(x_i); lazy input widths (``-1``)."""

    def __init__(self, in_channels, out_channels: int, aggr: str = "mean", normalize: bool = False,
                 root_weight: bool = True, bias: bool = True):
        super().__init__()
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.aggr, self.normalize, self.root_weight = aggr, normalize, root_weight
        self.lin_l = nn.LazyLinear(out_channels, bias=bias) if in_channels[0] <= 0 else Linear(in_channels[0], out_channels, bias=bias)
        if root_weight:
            self.lin_r = nn.LazyLinear(out_channels, bias=False) if in_channels[1] <= 0 else Linear(in_channels[1], out_channels, bias=False)

    # Opt-in (SAGEConv.project_first = True, or LGB_SAGE_PROJECT_FIRST=1): when the layer narrows (out < in, e.g. the last
    # encoder layer 128 -> 64) apply lin_l's weight BEFORE the neighbour aggregation -- sum/mean commute with a linear map,
    # lin_l(agg_j x_j) = agg_j(W x_j) + b -- so the gather/scatter kernel and its backward move out/in of the bytes.
    # Same function, different fp32 rounding order than the reference's (PyG's) evaluation, hence not the default.
    project_first = os.environ.get("LGB_SAGE_PROJECT_FIRST", "0") == "1"

    def forward(self, x, edge_index, size=None, graph: Optional[DeviceCSR] = None) -> torch.Tensor:
        x_src, x_dst = (x, x) if isinstance(x, torch.Tensor) else x
        if graph is None:
            graph = build_edge_csr(edge_index, x_src.shape[0], x_dst.shape[0])
        lazy = (isinstance(self.lin_l, nn.LazyLinear) and self.lin_l.has_uninitialized_params()) or \
            (self.root_weight and isinstance(self.lin_r, nn.LazyLinear) and self.lin_r.has_uninitialized_params())
        if (FUSED_SAGE and not lazy and not self.project_first and not self.normalize and self.aggr in ("add", "sum", "mean")
                and x_src.dtype == torch.float32 and _lib.on_device(x_src) and (not self.root_weight or x_dst is not None)):
            return _SageFn.apply(x_src, x_dst if self.root_weight else None, self.lin_l.weight, self.lin_l.bias,
                                 self.lin_r.weight if self.root_weight else None, graph, self.aggr == "mean")
        if (self.project_first and not lazy and self.aggr in ("add", "sum", "mean")
                and self.lin_l.out_features < self.lin_l.in_features):
            out = aggregate(F.linear(x_src, self.lin_l.weight), graph, self.aggr)
            if self.lin_l.bias is not None:
                out = out + self.lin_l.bias
        else:
            out = linear(self.lin_l, aggregate(x_src, graph, self.aggr))
        if self.root_weight and x_dst is not None:
            out = out + linear(self.lin_r, x_dst)
        if self.normalize:
            out = F.normalize(out, p=2.0, dim=-1)
        return out


def get_SAGEConv_layers(num_layers: int, hidden_channels: int, out_channels: int, agg_type: str) -> nn.ModuleList:
    """model/layers.py:6-32."""
    single = SAGEConv((-1, -1, -1), hidden_channels, aggr=agg_type, normalize=False, bias=True)
    last = SAGEConv((-1, -1, -1), out_channels, aggr=agg_type, normalize=False, bias=True)
    if num_layers == 1:
        return nn.ModuleList([last])
    return nn.ModuleList([copy.deepcopy(single) for _ in range(num_layers - 1)] + [last])


def get_linear_layers(num_layers: int, in_channels: int, hidden_channels: int, out_channels: int) -> nn.ModuleList:
    """model/layers.py:35-56.  The reference instantiates first / middle / last layer BEFORE it branches on num_layers (and
    builds fresh ones for 1 and 2 layers): the same three initialisations are drawn here, so that a seeded run initialises
    the decoder -- and everything constructed after it -- exactly like the reference."""
    first = Linear(in_channels, hidden_channels)
    middle = Linear(hidden_channels, hidden_channels)
    last = Linear(hidden_channels, out_channels)
    if num_layers == 1:
        return nn.ModuleList([Linear(in_channels, out_channels)])
    if num_layers == 2:
        return nn.ModuleList([Linear(in_channels, hidden_channels), Linear(hidden_channels, out_channels)])
    return nn.ModuleList([first] + [copy.deepcopy(middle) for _ in range(num_layers - 2)] + [last])


# --------------------------------------------------------------------------------------------
# encoder (GNNEncoder lifted over node / edge types)
# --------------------------------------------------------------------------------------------
class GNNEncoder(nn.Module):
    """model/encoder_decoder.py:17-46 (homogeneous form; ``to_hetero`` lifts it)."""

    def __init__(self, layers: ModuleList, p_dropout_edges: Optional[float], p_dropout_features: Optional[float]):
        super().__init__()
        self.layers = layers
        self.p_dropout_edges = p_dropout_edges
        self.p_dropout_features = p_dropout_features

    def forward(self, x, edge_index):
        for index, layer in enumerate(self.layers):
            if index == len(self.layers) - 1:
                x = layer(x, edge_index)
            else:
                if self.p_dropout_features is not None:
                    x = F.dropout(x, p=self.p_dropout_features, training=self.training)
                x = layer(x, edge_index).relu()
        return x


_FAN_IN = {"sum": torch.add, "mean": torch.add, "max": torch.max, "min": torch.min, "mul": torch.mul}


def _fan_in(outs: List[torch.Tensor], aggr: str) -> torch.Tensor:
    """Per-destination pairwise reduction of temporary_hetero.py:201-228 (pop two, combine, push back)."""
    n = len(outs)
    queue = deque(outs)
    while len(queue) >= 2:
        a, b = queue.popleft(), queue.popleft()
        queue.append(_FAN_IN[aggr](a, b))
    out = queue.popleft()
    return torch.div(out, n) if (aggr == "mean" and n > 1) else out


def _is_flip_of(a: torch.Tensor, b: torch.Tensor) -> bool:
    """a == b.flip(0), entry for entry (host-synchronising: used once per edge-type pair, see HeteroEncoder._flip_mate)."""
    return bool(torch.equal(a[0], b[1]) and torch.equal(a[1], b[0]))


class HeteroEncoder(nn.Module):
    """``to_hetero(GNNEncoder(...), metadata, aggr)``: one copy of every conv layer per edge type
    (sub-module names ``layers.<i>.<src>__<rel>__<dst>`` as PyG generates them), fan-in per destination."""

    def __init__(self, module: GNNEncoder, metadata, aggr: str = "sum", bake_dropout_training: bool = True):
        super().__init__()
        if aggr not in _FAN_IN:
            raise RuntimeError(f"to_hetero aggr={aggr!r} not supported")
        self.node_types, self.edge_types = list(metadata[0]), [tuple(e) for e in metadata[1]]
        self.aggr = aggr
        self.p_dropout_features = module.p_dropout_features
        self.bake_dropout_training = bake_dropout_training
        self.layers = nn.ModuleList([
            nn.ModuleDict({key2str(et): copy.deepcopy(layer) for et in self.edge_types}) for layer in module.layers])
        self._flip_of: Dict[tuple, Optional[tuple]] = {}      # edge type -> the edge type it is the flip of (decided on the first batch)

    def _flip_mate(self, et, graphs, edge_index_dict):
        """The already-built edge type that ``et`` is the exact flip of (``rev_buys`` of ``buys``, what ToUndirected produces), or
        None.  A property of the data pipeline, not of one batch: decided on the first batch with a host-side comparison and
        from then on re-verified on the device, the verdict examined one batch later (no host synchronisation per step)."""
        ei = edge_index_dict[et]
        if et not in self._flip_of:
            self._flip_of[et] = next((o for o in graphs if o[0] == et[2] and o[2] == et[0]
                                      and edge_index_dict[o].shape == ei.shape and _is_flip_of(ei, edge_index_dict[o])), None)
            return self._flip_of[et]
        mate = self._flip_of[et]
        if mate is None or mate not in graphs:
            return None
        other = edge_index_dict[mate]
        if other.shape != ei.shape:
            raise RuntimeError(f"edge type {et} was the flip of {mate} on the first batch but has a different number of edges now")
        differs = ((ei[0] != other[1]) | (ei[1] != other[0])).any().to(torch.int32)
        _defer_check(differs, f"edge type {et} was the exact flip of {mate} on the first batch but not on the previous one: "
                              "construct the model with reuse_flipped_edge_types=False")
        return mate

    reuse_flipped_edge_types = True

    def forward(self, x_dict: Dict[str, torch.Tensor], edge_index_dict, graphs: Optional[Dict[tuple, DeviceCSR]] = None
                ) -> Dict[str, torch.Tensor]:
        x = dict(x_dict)
        flush_deferred_checks()                    # verdicts of the previous batch (index range, flipped edge types)
        # one CSR (+ lazily its transpose) per edge type per batch, shared by all layers and the backward; an edge type
        # that is the exact flip of another one IS that one's transpose, so the pair costs two sorts instead of four.
        # ``graphs``: the batch's CSRs when the caller already has them (SubgraphSampler.sample()["graphs"]).
        graphs = dict(graphs) if graphs else {}
        for et in self.edge_types:
            if et in graphs:
                continue
            ei = edge_index_dict[et]
            mate = self._flip_mate(et, graphs, edge_index_dict) if self.reuse_flipped_edge_types else None
            graphs[et] = graphs[mate].transpose() if mate is not None else \
                build_edge_csr(ei, x[et[0]].shape[0], x[et[2]].shape[0], deferred=True)
        for li, convs in enumerate(self.layers):
            last = li == len(self.layers) - 1
            if not last and self.p_dropout_features is not None:
                training = True if self.bake_dropout_training else self.training
                x = {k: F.dropout(v, p=self.p_dropout_features, training=training) for k, v in x.items()}
            per_dst: Dict[str, List[torch.Tensor]] = defaultdict(list)
            for et in self.edge_types:
                s, _, t = et
                per_dst[t].append(convs[key2str(et)]((x[s], x[t]), edge_index_dict[et], graph=graphs[et]))
            x = {t: _fan_in(v, self.aggr) for t, v in per_dst.items()}
            if not last:
                x = {k: v.relu() for k, v in x.items()}
        return x


def to_hetero(module: nn.Module, metadata, aggr: str = "sum", **kwargs) -> HeteroEncoder:
    """The one use the reference makes of PyG ``to_hetero`` (model/encoder_decoder.py:93-95)."""
    if not hasattr(module, "layers") or not hasattr(module, "p_dropout_features"):
        raise RuntimeError("to_hetero here lifts GNNEncoder-shaped modules (a ModuleList `layers` of SAGEConv) only")
    return HeteroEncoder(module, metadata, aggr, **kwargs)


# --------------------------------------------------------------------------------------------
# edge decoder
# --------------------------------------------------------------------------------------------
class _EdgeConcat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, zu, zi, row, col):
        zu, zi, row, col = _lib.f32c(zu), _lib.f32c(zi), _lib.i64c(row), _lib.i64c(col)
        _lib.require_cuda(zu, zi, row, col)
        L, du, di = row.numel(), zu.shape[1], zi.shape[1]
        out = torch.empty(L, du + di, dtype=torch.float32, device=zu.device)
        with torch.cuda.device(zu.device):
            check(_lib.load().lgb_edge_concat_fwd(ptr(zu), ptr(zi), ptr(row), ptr(col), L, du, di, ptr(out), stream()),
                  "edge_concat_fwd")
        _lib.count_launch()
        ctx.save_for_backward(row, col)
        ctx.shapes = (zu.shape, zi.shape)
        return out

    @staticmethod
    def backward(ctx, gout):
        row, col = ctx.saved_tensors
        (nu, du), (ni, di) = ctx.shapes
        gout = _lib.f32c(gout)
        dzu = torch.empty(nu, du, dtype=torch.float32, device=gout.device) if ctx.needs_input_grad[0] else None
        dzi = torch.empty(ni, di, dtype=torch.float32, device=gout.device) if ctx.needs_input_grad[1] else None
        lib = _lib.load()
        with torch.cuda.device(gout.device):
            for t in (dzu, dzi):
                if t is not None:
                    check(lib.lgb_zero(ptr(t), t.numel() * 4, stream()), "zero")
            check(lib.lgb_edge_concat_bwd(ptr(gout), ptr(row), ptr(col), row.numel(), du, di, ptr(dzu), ptr(dzi), stream()),
                  "edge_concat_bwd")
        _lib.count_launch()
        return dzu, dzi, None, None


class _EdgeDot(torch.autograd.Function):
    @staticmethod
    def forward(ctx, zu, zi, row, col):
        zu, zi, row, col = _lib.f32c(zu), _lib.f32c(zi), _lib.i64c(row), _lib.i64c(col)
        _lib.require_cuda(zu, zi, row, col)
        if zu.shape[1] != zi.shape[1]:
            raise RuntimeError("dot decoder needs equal embedding widths")
        L, d = row.numel(), zu.shape[1]
        out = torch.empty(L, dtype=torch.float32, device=zu.device)
        with torch.cuda.device(zu.device):
            check(_lib.load().lgb_edge_dot_fwd(ptr(zu), ptr(zi), ptr(row), ptr(col), L, d, ptr(out), stream()), "edge_dot_fwd")
        _lib.count_launch()
        ctx.save_for_backward(zu, zi, row, col)
        return out

    @staticmethod
    def backward(ctx, gout):
        zu, zi, row, col = ctx.saved_tensors
        gout = _lib.f32c(gout)
        dzu = torch.empty_like(zu) if ctx.needs_input_grad[0] else None
        dzi = torch.empty_like(zi) if ctx.needs_input_grad[1] else None
        lib = _lib.load()
        with torch.cuda.device(gout.device):
            for t in (dzu, dzi):
                if t is not None:
                    check(lib.lgb_zero(ptr(t), t.numel() * 4, stream()), "zero")
            check(lib.lgb_edge_dot_bwd(ptr(zu), ptr(zi), ptr(row), ptr(col), ptr(gout), row.numel(), zu.shape[1],
                                       ptr(dzu), ptr(dzi), stream()), "edge_dot_bwd")
        _lib.count_launch()
        return dzu, dzi, None, None


def edge_concat(zu, zi, row, col):
    return _EdgeConcat.apply(zu, zi, row, col)


def edge_dot(zu, zi, row, col):
    return _EdgeDot.apply(zu, zi, row, col)


def edge_mlp2_inference(zu, zi, row, col, lin1, lin2) -> torch.Tensor:
    """Two-layer concat-MLP decoder without the per-edge concatenation (forward only, dropout off): the first Linear is applied
    to the NODES (``pu = zu W1[:, :du]^T + b1``, ``pi = zi W1[:, du:]^T``, two cuBLAS GEMMs), then lgb_edge_mlp2_fwd evaluates
    ``b2 + w2 . relu(pu[row] + pi[col])`` per label edge.  Same function as model/encoder_decoder.py:55-72 in eval mode, other
    fp32 summation order (rtol 1e-5 against the concat form, tests)."""
    zu, zi, row, col = _lib.f32c(zu), _lib.f32c(zi), _lib.i64c(row), _lib.i64c(col)
    _lib.require_cuda(zu, zi, row, col)
    du = zu.shape[1]
    W1, b1 = lin1.weight.detach(), lin1.bias
    pu = torch.addmm(b1.detach(), zu, W1[:, :du].t()) if b1 is not None else zu @ W1[:, :du].t()
    pi = zi @ W1[:, du:].t()
    L, H = row.numel(), W1.shape[0]
    out = torch.empty(L, dtype=torch.float32, device=zu.device)
    w2 = lin2.weight.detach().reshape(-1).contiguous()
    b2 = lin2.bias.detach() if lin2.bias is not None else None
    with torch.cuda.device(zu.device):
        check(_lib.load().lgb_edge_mlp2_fwd(ptr(pu), ptr(pi), ptr(row), ptr(col), L, H, ptr(w2), ptr(b2), ptr(out), stream()),
              "edge_mlp2_fwd")
    _lib.count_launch()
    return out


class EdgeDecoder(nn.Module):
    """model/encoder_decoder.py:49-72.  ``mode="mlp"`` is the reference (concat -> [dropout -> Linear -> relu]* ->
    Linear -> view(-1)); ``mode="dot"`` is the dot-product decoder of north_star (layers unused)."""

    def __init__(self, layers: ModuleList, p_dropout_features: Optional[float], mode: str = "mlp",
                 node_projection: str = "auto"):
        super().__init__()
        self.layers = layers
        self.p_dropout_features = p_dropout_features
        self.mode = mode
        # "auto": without autograd and without active dropout, a two-layer decoder whose label edges outnumber the nodes they
        # touch is evaluated through per-node projections (edge_mlp2_inference); "off": always the per-edge concat form
        self.node_projection = node_projection

    def _use_node_projection(self, zu, zi, L: int) -> bool:
        if self.node_projection == "off" or len(self.layers) != 2 or self.layers[1].out_features != 1:
            return False
        if any(isinstance(m, nn.LazyLinear) and m.has_uninitialized_params() for m in self.layers):
            return False
        if self.layers[0].in_features != zu.shape[1] + zi.shape[1] or not _lib.on_device(zu):
            return False
        if torch.is_grad_enabled() and (zu.requires_grad or zi.requires_grad or self.layers[0].weight.requires_grad):
            return False
        if self.training and self.p_dropout_features:
            return False
        if self.node_projection == "on":
            return True
        du, di = zu.shape[1], zi.shape[1]
        return L * (du + di) > zu.shape[0] * du + zi.shape[0] * di      # rows through the first GEMM: edges vs nodes

    def forward(self, z_dict: dict, edge_label_index) -> torch.Tensor:
        customer_index, article_index = edge_label_index
        zu, zi = z_dict[NODE_USER], z_dict[NODE_ITEM]
        if self.mode == "dot":
            return edge_dot(zu, zi, customer_index, article_index)
        if self._use_node_projection(zu, zi, customer_index.numel()):
            return edge_mlp2_inference(zu, zi, customer_index, article_index, self.layers[0], self.layers[1])
        z = edge_concat(zu, zi, customer_index, article_index)
        for index, layer in enumerate(self.layers):
            if index == len(self.layers) - 1:
                z = linear(layer, z)
            else:
                if self.p_dropout_features is not None:
                    z = F.dropout(z, p=self.p_dropout_features, training=self.training)
                z = linear(layer, z).relu()
        return z.view(-1)


# --------------------------------------------------------------------------------------------
# the model
# --------------------------------------------------------------------------------------------
def padded_stack(tensors: List[torch.Tensor], value=0) -> torch.Tensor:
    """utils/tensor.py:24-61 (right padding, constant)."""
    full = max(x.size(-1) for x in tensors)
    return torch.stack([F.pad(x, (0, full - x.size(-1)), value=value) if full > x.size(-1) else x for x in tensors], dim=0)


def rebatch_by_user(out: torch.Tensor, users: torch.Tensor, value) -> torch.Tensor:
    """The re-batching of ``infer`` (model/encoder_decoder.py:161-164: ``unique`` + one boolean mask per user + padded_stack)
    as one segmented scatter: row r holds the scores of the r-th smallest user id in their original edge order, right-padded
    with ``value``.  Same [users_in_batch, max_candidates] result, no per-user Python loop."""
    uniq, inv, counts = torch.unique(users, sorted=True, return_inverse=True, return_counts=True)
    if uniq.numel() == 0:
        return out.new_empty(0, 0)
    order = torch.sort(inv, stable=True).indices                  # edges grouped by user, original order kept inside a group
    starts = torch.cumsum(counts, 0) - counts
    pos = torch.arange(users.numel(), device=users.device) - starts[inv[order]]
    res = out.new_full((uniq.numel(), int(counts.max())), float(value))
    res[inv[order], pos] = out[order]
    return res


class Encoder_Decoder_Model(nn.Module):
    """Drop-in for model/encoder_decoder.py:75-164 (same ctor kwargs, forward / infer /
    initialize_encoder_input_size, same state_dict keys)."""

    def __init__(self, encoder_layers: ModuleList, decoder_layers: ModuleList, feature_info: dict, metadata,
                 embedding: bool, heterogeneous_prop_agg_type: str, batch_normalize: bool,
                 p_dropout_edges: Optional[float], p_dropout_features: Optional[float],
                 decoder_mode: str = "mlp", bake_dropout_training: bool = True):
        super().__init__()
        self.embedding = embedding
        self.batch_normalize = batch_normalize
        self.encoder = to_hetero(GNNEncoder(encoder_layers, p_dropout_edges, p_dropout_features), metadata,
                                 aggr=heterogeneous_prop_agg_type, bake_dropout_training=bake_dropout_training)
        self.decoder = EdgeDecoder(decoder_layers, p_dropout_features, mode=decoder_mode)
        self.encoder_layer_norm_customer = BatchNorm1d(encoder_layers[-1].out_channels)
        self.encoder_layer_norm_article = BatchNorm1d(encoder_layers[-1].out_channels)
        # reference quirk: a plain dict -> not registered, not trained, not in state_dict (model/encoder_decoder.py:101-114)
        self.embedding_layers = dict()
        if self.embedding:
            for key, item in feature_info.items():
                self.embedding_layers[key] = ModuleList([
                    Embedding(num_embeddings=int(item.num_cat[i] + 1), embedding_dim=int(item.embedding_size[i]), max_norm=1)
                    for i in range(item.num_feat)])

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        for ml in self.embedding_layers.values():   # follow the module across devices (still unregistered)
            ml._apply(fn)
        return out

    def _embed(self, x_dict: dict) -> dict:
        for key, item in self.embedding_layers.items():
            features = x_dict[key]
            x_dict[key] = torch.cat([layer(features[:, i]) for i, layer in enumerate(item)], dim=1)
        return x_dict

    def initialize_encoder_input_size(self, data) -> None:
        x_dict, edge_index_dict = data.x_dict, data.edge_index_dict
        if self.embedding:
            x_dict = self._embed(x_dict)
        self.encoder(x_dict, edge_index_dict)

    def forward(self, x_dict, edge_index_dict: dict, edge_label_index: torch.Tensor, graphs=None) -> torch.Tensor:
        if self.embedding:
            x_dict = self._embed(x_dict)
        z_dict = self.encoder(x_dict, edge_index_dict, graphs=graphs) if graphs else self.encoder(x_dict, edge_index_dict)
        if self.batch_normalize:
            z_dict[NODE_USER] = self.encoder_layer_norm_customer(z_dict[NODE_USER])
            z_dict[NODE_ITEM] = self.encoder_layer_norm_article(z_dict[NODE_ITEM])
        return self.decoder(z_dict, edge_label_index)

    def infer(self, x_dict, edge_index_dict: dict, edge_label_index: torch.Tensor) -> torch.Tensor:
        self.eval()
        with torch.no_grad():           # the reference detaches the scores; without a tape the decoder may use its inference form
            out = self.forward(x_dict, edge_index_dict, edge_label_index)
        return rebatch_by_user(out, edge_label_index[0], value=-(1 << 50))
