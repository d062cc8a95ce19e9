"""LightGCN with the reference's constructor / forward / attributes (model/lightgcn.py:11-87) on the
sm_100a kernels of liblaplace_b200.

HBM layout: the two embedding tables live back to back in ONE [U+I, d] fp32 buffer; ``users_emb.weight``
and ``items_emb.weight`` are Parameters whose storage are the two row ranges of it, so E^0 = cat(Wu, Wi)
(model/lightgcn.py:58) costs nothing and the backward hands the two gradient row-ranges of one buffer
straight to the optimizer.  Forward keeps two ping-pong layer buffers and one accumulator:

    layer 1:   Y1 = A E0        acc  = E0 + Y1            (one fused kernel)
    layer k:   Yk = A Y(k-1)    acc += Yk
    layer K:                    E_f  = (acc + A Y(K-1)) / (K+1)      (Y_K is never written)

which replaces K SpMMs + stack + mean (+ the [N, K+1, d] tensor) of the reference.  Backward is the Horner
form g <- A^T g + r with r = dE_f/(K+1), K launches of the same kernel on the CSC arrays.
"""
from __future__ import annotations

import os

import ctypes as C
from typing import Optional, Tuple

import torch
from torch import nn

from . import _lib
from ._lib import check, ptr, stream
from .bpr import bpr_indexed
from .csr import DeviceCSR, rows_bitmap
from .sparse import SparseTensor, gcn_norm, matmul


_AUTOTUNE = os.environ.get("LGB_SPMM_AUTOTUNE", "0") == "1"   # tune the SpMM variant per graph on first use


# plan space of LightGCN.autotune(thorough=True) and bench.py: slice sizes (the CTA-wide-slice variants walk a slice with four warps,
# so 2048 / 4096 keep their per-warp chain at 512 / 1024 while quartering the partial rows of stage 2), row order, slice order
AUTOTUNE_SPACE = dict(chunks=(4096, 2048, 1024, 512), degree_orders=(False, True), sweeps=(False, True))


# first backward layer as lgb_spmm_rowsparse while the batch rows are at most 1 / ROWSPARSE_MAX_SHARE of the table
ROWSPARSE_BACKWARD = os.environ.get("LGB_ROWSPARSE_BACKWARD", "1") != "0"
ROWSPARSE_MAX_SHARE = 8


def _scale_concat(a: Optional[torch.Tensor], b: Optional[torch.Tensor], na: int, nb: int, d: int, scale: float,
                  out: torch.Tensor) -> torch.Tensor:
    with torch.cuda.device(out.device):
        check(_lib.load().lgb_scale_concat(ptr(a), na, ptr(b), nb, d, float(scale), ptr(out), stream()), "scale_concat")
    _lib.count_launch()
    return out


def propagate_forward(g: DeviceCSR, E0: torch.Tensor, K: int) -> torch.Tensor:
    """E_f = mean_k (A^k E0), k = 0..K, with the fused accumulate epilogue (no stack, no Y_K)."""
    if K == 0:
        return E0.clone()
    N, d = E0.shape
    E_f = torch.empty_like(E0)
    if K == 1:
        g.spmm(E0, want_y=False, acc_in=E0, acc_out=E_f, acc_div=2.0)
        return E_f
    ya = torch.empty_like(E0)
    yb = torch.empty_like(E0) if K > 2 else None
    g.spmm(E0, Y=ya, acc_in=E0, acc_out=E_f)
    x, y = ya, yb
    for _ in range(K - 2):
        g.spmm(x, Y=y, acc_in=E_f, acc_out=E_f)
        x, y = y, x
    g.spmm(x, want_y=False, acc_in=E_f, acc_out=E_f, acc_div=float(K + 1))
    return E_f


def propagate_backward(gt: DeviceCSR, r: torch.Tensor, K: int, out: Optional[torch.Tensor] = None,
                       r_rows: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dE0 = sum_{k=0..K} (A^T)^k r  via Horner: g <- A^T g + r, K times (r already holds dE_f/(K+1)).
    ``r_rows``: bitmap of the rows of r that may be non-zero (the BPR batch): the first layer gathers from r itself and runs as
    lgb_spmm_rowsparse -- of the nnz entries only those whose column is a batch row are fetched."""
    if K == 0:
        return r
    g = r
    bufs = [None, None]
    for k in range(K):
        last = k == K - 1
        dst = out if (last and out is not None) else bufs[k % 2]
        if dst is None:
            dst = torch.empty_like(r)
            if not last:
                bufs[k % 2] = dst
        gt.spmm(g, Y=dst, resid=r, x_rows=r_rows if k == 0 else None, resid_rows=r_rows if k == 0 else None)
        g = dst
    return g


class _Propagate(torch.autograd.Function):
    """E_f = mean_k A^k [Wu; Wi], returned as its user and item blocks (two views of one buffer): the backward then receives
    the two gradient blocks as they are and scales / concatenates them in ONE pass, instead of autograd's cat for a split."""

    @staticmethod
    def forward(ctx, Wu, Wi, model: "LightGCN", g: DeviceCSR, K: int):
        E0 = model._table_for(Wu, Wi)
        ctx.g, ctx.K, ctx.U, ctx.model = g, K, Wu.shape[0], model
        ctx.shape = tuple(E0.shape)
        E_f = propagate_forward(g, E0, K)
        return E_f[: ctx.U], E_f[ctx.U:]

    @staticmethod
    def backward(ctx, gU, gI):
        K, U = ctx.K, ctx.U
        N, d = ctx.shape
        gU = _lib.f32c(gU) if gU is not None else None
        gI = _lib.f32c(gI) if gI is not None else None
        dev = (gU if gU is not None else gI).device
        r = torch.empty(N, d, dtype=torch.float32, device=dev)
        r_rows = None
        if ROWSPARSE_BACKWARD and K > 0 and d in (16, 32, 64):
            # dE_f arrives dense, but under mini-batch BPR only the batch rows are non-zero: the scaling pass notes which
            r_rows = ctx.model._scale_and_find_rows(gU, gI, U, N - U, d, 1.0 / (K + 1), r)
        else:
            _scale_concat(gU, gI, U, N - U, d, 1.0 / (K + 1), r)
        G = propagate_backward(ctx.g.transpose(), r, K, r_rows=r_rows)
        return G[:U], G[U:], None, None, None


class LightGCN(nn.Module):
    """Drop-in for the reference ``model.lightgcn.LightGCN`` (same ctor, forward, attributes, state_dict keys)."""

    def __init__(self, num_users, num_items, embedding_dim: int, num_iterations: int, add_self_loops=False):
        super().__init__()
        self.num_users, self.num_items = int(num_users), int(num_items)
        self.embedding_dim, self.num_iterations = int(embedding_dim), int(num_iterations)
        self.add_self_loops = add_self_loops
        # same construction + init order as the reference (model/lightgcn.py:36-44) so seeded runs draw the same values
        self.users_emb = nn.Embedding(num_embeddings=self.num_users, embedding_dim=self.embedding_dim)
        self.items_emb = nn.Embedding(num_embeddings=self.num_items, embedding_dim=self.embedding_dim)
        nn.init.normal_(self.users_emb.weight, std=0.1)
        nn.init.normal_(self.items_emb.weight, std=0.1)
        self._table: Optional[torch.Tensor] = None
        self._fuse_tables()

    # ---- one [U+I, d] buffer behind both Parameters ------------------------------------------
    def _fuse_tables(self) -> None:
        Wu, Wi = self.users_emb.weight, self.items_emb.weight
        table = torch.empty(self.num_users + self.num_items, self.embedding_dim, dtype=Wu.dtype, device=Wu.device)
        table[: self.num_users].copy_(Wu.data)
        table[self.num_users:].copy_(Wi.data)
        Wu.data = table[: self.num_users]
        Wi.data = table[self.num_users:]
        self._table = table

    def _is_fused(self, Wu: torch.Tensor, Wi: torch.Tensor) -> bool:
        t = self._table
        return (t is not None and t.device == Wu.device and Wu.data_ptr() == t.data_ptr()
                and Wi.data_ptr() == t.data_ptr() + self.num_users * self.embedding_dim * t.element_size()
                and Wu.is_contiguous() and Wi.is_contiguous())

    def _table_for(self, Wu: torch.Tensor, Wi: torch.Tensor) -> torch.Tensor:
        """E^0 = cat(Wu, Wi) (model/lightgcn.py:58): zero-copy when the tables are fused, else one copy kernel."""
        if self._is_fused(Wu, Wi):
            return self._table
        d = self.embedding_dim
        out = torch.empty(self.num_users + self.num_items, d, dtype=torch.float32, device=Wu.device)
        return _scale_concat(_lib.f32c(Wu.detach()), _lib.f32c(Wi.detach()), self.num_users, self.num_items, d, 1.0, out)

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if not self._is_fused(self.users_emb.weight, self.items_emb.weight):
            self._fuse_tables()
        return out

    def _rows_bitmap_buf(self, n: int, device) -> torch.Tensor:
        buf = getattr(self, "_bitmap", None)
        if buf is None or buf.numel() * 32 < n or buf.device != device:
            buf = self._bitmap = torch.empty((n + 31) // 32, dtype=torch.int32, device=device)
        return buf

    def _scale_and_find_rows(self, gU, gI, U: int, I: int, d: int, scale: float, r: torch.Tensor):
        """r = scale * [gU; gI] (None = zeros) plus the bitmap of its non-zero rows (lgb_scale_rows_nonzero, one pass per block).
        Whether the bitmap is worth using (few rows flagged) is a host decision: the row count is read back on the first call and
        every 64th after it (one small synchronisation each), in between the last answer stands -- a stale answer costs speed,
        never correctness.  Returns the bitmap, or None when the gradient is (taken to be) dense."""
        N = U + I
        dev = r.device
        bm = self._rows_bitmap_buf(N, dev)
        cnt = getattr(self, "_rows_count", None)
        if cnt is None or cnt.device != dev:
            cnt = self._rows_count = torch.zeros(1, dtype=torch.int32, device=dev)
        lib = _lib.load()
        with torch.cuda.device(dev):
            check(lib.lgb_zero(ptr(bm), bm.numel() * 4, stream()), "zero")
            check(lib.lgb_zero(ptr(cnt), 4, stream()), "zero")
            for g_blk, lo, n in ((gU, 0, U), (gI, U, I)):
                if n == 0:
                    continue
                if g_blk is None:
                    check(lib.lgb_zero(ptr(r[lo:lo + n]), n * d * 4, stream()), "zero")
                else:
                    check(lib.lgb_scale_rows_nonzero(ptr(g_blk), n, d, float(scale), ptr(r[lo:lo + n]), lo, ptr(bm), ptr(cnt), stream()),
                          "scale_rows_nonzero")
                    _lib.count_launch()
        st = self.__dict__.setdefault("_rows_state", {"calls": 0, "sparse": False})
        capturing = dev.type == "cuda" and torch.cuda.is_current_stream_capturing()
        if st["calls"] % 64 == 0 and not capturing:
            st["sparse"] = int(cnt.item()) * ROWSPARSE_MAX_SHARE <= N
        st["calls"] += 1
        return bm if st["sparse"] else None

    # ---- plan-time kernel selection (optional) ------------------------------------------------
    def autotune(self, edge_index: SparseTensor, thorough: Optional[bool] = None):
        """Time the SpMM kernel variants on THIS graph (and its transpose, for the backward) at this embedding width and
        keep the fastest for all later calls (``DeviceCSR.autotune``).  One-off per graph; also run automatically on first use
        when LGB_SPMM_AUTOTUNE=1.  ``thorough`` (default: graphs of a million entries and more) also walks the slice size of
        the long-row plan, the degree-bucketed row order and the column-sweep order of the slices -- on the H&M-shaped graph
        that search is what takes the epoch from 8.1 to 6.9 ms (profiles/README.md r2l).  Returns (forward, backward variant)."""
        g = gcn_norm(edge_index, add_self_loops=self.add_self_loops).csr()
        d = self.users_emb.weight.shape[1]
        if thorough is None:
            thorough = g.nnz >= 1_000_000
        space = AUTOTUNE_SPACE if thorough else {}
        return g.autotune(d, **space), g.transpose().autotune(d, **space)

    # ---- reference API ------------------------------------------------------------------------
    def forward(self, edge_index: SparseTensor):
        """-> (e_u^K, e_u^0, e_i^K, e_i^0); elements 1 and 3 are the Parameters themselves, like the reference."""
        Wu, Wi = self.users_emb.weight, self.items_emb.weight
        _lib.require_cuda(Wu, Wi)
        adj = gcn_norm(edge_index, add_self_loops=self.add_self_loops)   # cached per graph
        g = adj.csr()
        if _AUTOTUNE and g.variant is None:
            self.autotune(edge_index)
        if g.n_rows != self.num_users + self.num_items or g.n_cols != g.n_rows:
            raise RuntimeError(f"adjacency is {g.n_rows}x{g.n_cols}, expected a square matrix of "
                               f"{self.num_users + self.num_items} nodes")
        users_emb_final, items_emb_final = _Propagate.apply(Wu, Wi, self, g, self.num_iterations)
        return users_emb_final, Wu, items_emb_final, Wi

    def propagate(self, edge_index: SparseTensor, x: torch.Tensor, **kwargs) -> torch.Tensor:
        return self.message_and_aggregate(edge_index, x)

    def message(self, x_j: torch.Tensor) -> torch.Tensor:
        return x_j

    def message_and_aggregate(self, adj_t: SparseTensor, x: torch.Tensor) -> torch.Tensor:
        return matmul(adj_t, x)

    # ---- CUDA-graph replay of the fused step (launch-bound graphs: ML-1M-sized and smaller) -------------------
    def capture_step(self, edge_index: SparseTensor, batch_size: int, lambda_val: float):
        """Capture ``fused_step`` for a fixed batch size into ONE CUDA graph and return ``step(u, p, n) -> loss`` that copies
        the three index vectors into static buffers and replays it: the 2K + 3 kernel launches (+ memset) of an iteration
        become a single graph launch, which is what bounds an iteration on graphs whose SpMM takes tens of microseconds.
        The gradients land in ``users_emb.weight.grad`` / ``items_emb.weight.grad`` (same buffers on every replay), the loss
        in the returned 0-dim tensor.  The embedding tables are read in place, so optimizer steps between replays are seen."""
        Wu = self.users_emb.weight
        _lib.require_cuda(Wu)
        dev = Wu.device
        g = gcn_norm(edge_index, add_self_loops=self.add_self_loops).csr()
        g.transpose()                                                                   # build everything outside the capture
        if _AUTOTUNE and g.variant is None:
            self.autotune(edge_index)                                                   # ... including the plan-time tuning
        su, sp, sn = (torch.zeros(int(batch_size), dtype=torch.int64, device=dev) for _ in range(3))
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                     # warm-up on a side stream, as torch.cuda.graph requires
            for _ in range(2):
                self.fused_step(edge_index, su, sp, sn, lambda_val)
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss = self.fused_step(edge_index, su, sp, sn, lambda_val)
        grads = (self.users_emb.weight.grad, self.items_emb.weight.grad)

        def step(user_indices, pos_item_indices, neg_item_indices):
            su.copy_(user_indices, non_blocking=True)
            sp.copy_(pos_item_indices, non_blocking=True)
            sn.copy_(neg_item_indices, non_blocking=True)
            graph.replay()
            self.users_emb.weight.grad, self.items_emb.weight.grad = grads      # zero_grad(set_to_none) may have dropped them
            return loss
        step.graph = graph
        return step

    # ---- fused training step (no autograd graph, no gathered copies) ------------------------------
    @torch.no_grad()
    def fused_step(self, edge_index: SparseTensor, user_indices: torch.Tensor, pos_item_indices: torch.Tensor,
                   neg_item_indices: torch.Tensor, lambda_val: float) -> torch.Tensor:
        """One iteration of run_pipeline_lightgcn.py:120-158 (forward, six gathers, bpr_loss, backward) as
        2K + 3 kernels: returns the loss (0-dim device tensor) and leaves the gradients in
        ``users_emb.weight.grad`` / ``items_emb.weight.grad`` (overwritten, i.e. zero_grad + backward)."""
        Wu, Wi = self.users_emb.weight, self.items_emb.weight
        _lib.require_cuda(Wu, Wi)
        K, U = self.num_iterations, self.num_users
        g = gcn_norm(edge_index, add_self_loops=self.add_self_loops).csr()
        if _AUTOTUNE and g.variant is None:
            self.autotune(edge_index)
        E0 = self._table_for(Wu, Wi)
        N, d = E0.shape
        E_f = propagate_forward(g, E0, K)
        loss = torch.empty((), dtype=torch.float32, device=E0.device)
        r = torch.empty_like(E0)
        with torch.cuda.device(E0.device):
            check(_lib.load().lgb_zero(ptr(r), r.numel() * 4, stream()), "zero")
        # loss + d/dE_f (already divided by K+1), scattered with vector atomics
        bpr_indexed(E_f, E0, U, user_indices, pos_item_indices, neg_item_indices, lambda_val, loss=loss, dE_f=r,
                    gscale=1.0 / (K + 1))
        if K == 0:
            G = r
        else:
            # dE_f is non-zero on the batch rows only: while they are a small part of the table, say so to the first layer
            r_rows = None
            if ROWSPARSE_BACKWARD and 3 * user_indices.numel() * ROWSPARSE_MAX_SHARE <= N and d % 4 == 0 and d <= 64:
                r_rows = rows_bitmap(N, ((user_indices, 0), (pos_item_indices, U), (neg_item_indices, U)), E0.device,
                                     out=self._rows_bitmap_buf(N, E0.device))
            G = propagate_backward(g.transpose(), r, K, r_rows=r_rows)
        # + 2*lambda*E0[rows] on the layer-0 rows of the batch
        bpr_indexed(E_f, E0, U, user_indices, pos_item_indices, neg_item_indices, lambda_val, dE_0=G)
        Wu.grad, Wi.grad = G[:U], G[U:]
        self.last_final = E_f
        return loss
