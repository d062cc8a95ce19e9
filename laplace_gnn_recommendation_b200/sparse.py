"""`torch_sparse.SparseTensor` / `matmul` and PyG `gcn_norm` call surface, backed by DeviceCSR.

The reference builds its adjacency with ``SparseTensor(row=, col=, sparse_sizes=)`` on the CPU and moves it
with ``.to(device)`` (data/lightgcn_loader.py:65-79, run_pipeline_lightgcn.py:108-111).  Here the object is a
thin COO holder until it reaches a CUDA device, where the CSR (and lazily the CSC) is built once by
liblaplace_b200 and cached; ``gcn_norm`` is cached per graph too (the reference recomputes it every
forward, model/lightgcn.py:56 -- the graph is immutable during training so caching is exact).
There is no CPU compute path: using a CPU SparseTensor in ``matmul`` raises.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from .csr import DEFAULT_CHUNK, DeviceCSR


class SparseTensor:
    def __init__(self, row: Optional[torch.Tensor] = None, col: Optional[torch.Tensor] = None,
                 value: Optional[torch.Tensor] = None, sparse_sizes: Optional[Tuple[int, int]] = None,
                 rowptr: Optional[torch.Tensor] = None, is_sorted: bool = False, trust_data: bool = False):
        if row is None or col is None:
            raise RuntimeError("SparseTensor needs row= and col= (rowptr-only construction is not supported)")
        if sparse_sizes is None:
            sparse_sizes = (int(row.max()) + 1 if row.numel() else 0, int(col.max()) + 1 if col.numel() else 0)
        self._sizes = (int(sparse_sizes[0]), int(sparse_sizes[1]))
        self._row, self._col, self._value = row, col, value
        self._csr: Optional[DeviceCSR] = None
        self._norm: Optional["SparseTensor"] = None  # cached gcn_norm(add_self_loops=False)
        self.chunk = DEFAULT_CHUNK

    # -- torch_sparse surface used by the reference ------------------------------------------
    def sparse_sizes(self) -> Tuple[int, int]:
        return self._sizes

    def size(self, dim: int) -> int:
        return self._sizes[dim]

    def nnz(self) -> int:
        return int(self._row.numel()) if self._csr is None else self._csr.nnz

    def has_value(self) -> bool:
        return self._value is not None or (self._csr is not None and self._csr.val is not None)

    @property
    def device(self):
        return self._row.device if self._csr is None else self._csr.device

    def is_cuda(self) -> bool:
        return self.device.type == "cuda"

    def to(self, device, *args, **kwargs) -> "SparseTensor":
        device = torch.device(device)
        here = self.device
        if device.type == here.type and (device.index is None or device.index == here.index):
            return self            # 'cuda' names the device the tensor already lives on: keep the cached CSR / gcn_norm
        if self._csr is not None:
            raise RuntimeError("moving a built SparseTensor to another device is not supported (its COO copy was freed "
                               "when the CSR was built): construct it again from row/col")
        out = SparseTensor(row=self._row.to(device), col=self._col.to(device),
                           value=None if self._value is None else self._value.to(device), sparse_sizes=self._sizes)
        out.chunk = self.chunk
        return out

    def cuda(self) -> "SparseTensor":
        return self.to("cuda")

    # -- device graph ---------------------------------------------------------------------------
    def csr(self) -> DeviceCSR:
        """Build (once) and return the device CSR.  Raises on a CPU tensor: there is no CPU fallback."""
        if self._csr is None:
            _lib.require_cuda(self._row)   # a CPU SparseTensor must be moved with .to('cuda') first: no CPU fallback
            g = DeviceCSR.from_coo(self._row, self._col, self._sizes[0], self._sizes[1], chunk=self.chunk,
                                   want_perm=self._value is not None)
            if self._value is not None:
                g.val = _lib.f32c(self._value.to(torch.float32))[g.perm].contiguous()
            self._csr = g
            self._row = self._col = None  # the COO copy is no longer needed (saves 16 B / nnz of HBM)
            self._value = None
        return self._csr

    @classmethod
    def from_csr(cls, g: DeviceCSR) -> "SparseTensor":
        out = cls.__new__(cls)
        out._sizes = (g.n_rows, g.n_cols)
        out._row = out._col = out._value = None
        out._csr, out._norm, out.chunk = g, None, g.chunk
        return out

    def t(self) -> "SparseTensor":
        return SparseTensor.from_csr(self.csr().transpose())


def gcn_norm(edge_index: SparseTensor, edge_weight=None, num_nodes=None, improved: bool = False,
             add_self_loops: bool = True, dtype=None) -> SparseTensor:
    """PyG ``gcn_norm`` for a SparseTensor (reference model/lightgcn.py:56).  Cached on the input."""
    if not isinstance(edge_index, SparseTensor):
        raise RuntimeError("gcn_norm: only the SparseTensor form used by the reference is supported")
    if edge_index.has_value():
        raise RuntimeError("gcn_norm: the reference always passes a value-less adjacency")
    if add_self_loops:
        raise RuntimeError("gcn_norm(add_self_loops=True) is not on the reference path (LightGCN passes False)")
    if edge_index._norm is None:
        g = edge_index.csr()
        _, val = g.gcn_norm()
        edge_index._norm = SparseTensor.from_csr(g.with_values(val))
    return edge_index._norm


def matmul(src: SparseTensor, other: torch.Tensor, reduce: str = "sum") -> torch.Tensor:
    """``torch_sparse.matmul(adj_t, x)`` (reference model/lightgcn.py:85-87), differentiable w.r.t. ``other``."""
    if reduce not in ("sum", "add", "mean"):
        raise RuntimeError(f"matmul: reduce={reduce!r} not supported")
    return _SpMM.apply(other, src.csr(), reduce == "mean")


class _SpMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, g: DeviceCSR, mean: bool):
        ctx.g, ctx.mean = g, mean
        return g.spmm(X, mean=mean)

    @staticmethod
    def backward(ctx, gY):
        g: DeviceCSR = ctx.g
        gY = _lib.f32c(gY)
        if ctx.mean:
            lib = _lib.load()
            scaled = torch.empty_like(gY)
            _lib.check(lib.lgb_row_div_by_degree(_lib.ptr(gY), _lib.ptr(g.rowptr), g.n_rows, gY.shape[1],
                                                 _lib.ptr(scaled), _lib.stream()), "row_div_by_degree")
            _lib.count_launch()
            gY = scaled
        return g.transpose().spmm(gY), None, None
