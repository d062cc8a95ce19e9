"""Generic multi-GPU LightGCN for an ARBITRARY [N, N] adjacency (SURVEY.md 8e "generic row-range sharding"): the layout
north_star spells out -- embedding table and adjacency rows range-partitioned over the GPUs, the rows of X a layer needs
from the other ranks moved with an NCCL all-gather, gradients flowing back the same way on the transposed shards.

``dist.ShardedLightGCN`` (users sharded, items replicated) is the fast path for the symmetric bipartite wiring; this
engine is the fallback that also covers the reference's own wiring (R) (rows = user ids, cols = item ids in [0, I), a
non-symmetric matrix, ``data/lightgcn_loader.py:65-69``) and anything else that is not bipartite-symmetric.

Rank g owns the nodes [b_g, b_{g+1}) (split so that the non-zeros of its forward AND backward shards balance):
    table     [n_g, d]            its rows of E0 = cat(Wu, Wi)                      (model/lightgcn.py:58)
    G_fwd     [n_g, G*P] CSR      its rows of A_hat                                  values dinv[row]*dinv[col] (global degrees)
    G_bwd     [n_g, G*P] CSR      its rows of A_hat^T (= its columns of A_hat)
Column ids are remapped once to the layout of the all-gather output: node v owned by rank r sits in slot r*P + (v - b_r),
P = max_g n_g, so `all_gather_into_tensor` of the (padded) local blocks IS the operand of the SpMM -- no compaction pass.
A layer = all-gather(X) + one lgb_spmm with the fused accumulate / mean / residual epilogue on the local rows.
The B-triple BPR is evaluated redundantly by every rank on the gathered final embeddings (B*3 rows; the kernel's loss
reduction is deterministic, so the loss is bit-identical on all ranks without a collective); each rank keeps the gradient
rows of the nodes it owns.  Per layer and rank N*d*4*(G-1)/G bytes arrive over NVLink (331 MB at the H&M shape, d=64):
the all-gather is exposed in this first version, which is why the bipartite engine is the default where it applies.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib
from .bpr import LgbBprArgs, _launch as _bpr_launch, _ws as _bpr_ws
from .csr import DEFAULT_CHUNK, DeviceCSR
from .dist import CudaOps, balanced_user_bounds


def all_gather_blocks(ops, full: torch.Tensor, block: torch.Tensor) -> None:
    """full[r*P:(r+1)*P] = rank r's block (all blocks [P, d]); stream-ordered like every other call of the step."""
    group = getattr(ops, "group", None)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_gather_into_tensor(full, block, group=group)
    else:
        full[: block.shape[0]].copy_(block)


class RowShardedLightGCN:
    """fused_step(u, p, n, lambda) on the global batch -> global loss; ``grad`` holds dLoss/dE0 for the owned node rows
    (node order = users then items, like the reference's ``t.cat([users_emb.weight, items_emb.weight])``)."""

    def __init__(self, num_users: int, num_items: int, embedding_dim: int, num_iterations: int, row: torch.Tensor,
                 col: torch.Tensor, device, group=None, ops=None, rank: Optional[int] = None, world: Optional[int] = None,
                 init_tables=None, chunk: int = DEFAULT_CHUNK):
        self.U, self.I, self.d, self.K = int(num_users), int(num_items), int(embedding_dim), int(num_iterations)
        self.N = self.U + self.I
        self.device = torch.device(device)
        inited = dist.is_available() and dist.is_initialized()
        self.rank = rank if rank is not None else (dist.get_rank(group) if inited else 0)
        self.world = world if world is not None else (dist.get_world_size(group) if inited else 1)
        self.ops = ops if ops is not None else CudaOps(self.device, group)
        row, col = _lib.i64c(row.to(self.device)), _lib.i64c(col.to(self.device))
        N, G = self.N, self.world

        # gcn_norm(add_self_loops=False) on the GLOBAL matrix (model/lightgcn.py:56): dinv from the row sums, both factors
        deg = torch.bincount(row, minlength=N).to(torch.float32)
        dinv = torch.where(deg > 0, 1.0 / torch.sqrt(deg), torch.zeros_like(deg))
        val = dinv[row] * dinv[col]
        # node ranges that balance forward + backward non-zeros
        work = torch.bincount(row, minlength=N) + torch.bincount(col, minlength=N)
        self.bounds: List[int] = balanced_user_bounds(work, G)
        self.lo, self.hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        self.n_loc = self.hi - self.lo
        self.P = max(max(b - a for a, b in zip(self.bounds, self.bounds[1:])), 1)
        starts = torch.tensor(self.bounds[:-1], dtype=torch.int64, device=self.device)
        ends = torch.tensor(self.bounds[1:], dtype=torch.int64, device=self.device)
        self._starts, self._ends = starts, ends

        def shard(r_own: torch.Tensor, c_other: torch.Tensor) -> DeviceCSR:
            m = (r_own >= self.lo) & (r_own < self.hi)
            g = DeviceCSR.from_coo(r_own[m] - self.lo, self.slot(c_other[m]), self.n_loc, G * self.P, chunk=chunk, want_perm=True)
            g.val = val[m][g.perm].contiguous()
            g.perm = None
            return g
        self.g_fwd = shard(row, col)          # rows of A_hat
        self.g_bwd = shard(col, row)          # rows of A_hat^T
        self.local_nnz = (self.g_fwd.nnz, self.g_bwd.nnz)

        f32 = dict(dtype=torch.float32, device=self.device)
        P, d = self.P, self.d
        self._table_pad = torch.zeros(P, d, **f32)
        self.table = self._table_pad[: self.n_loc]
        if init_tables is not None:
            Wu, Wi = init_tables
            full = torch.cat([Wu.to(self.device), Wi.to(self.device)])
            self.table.copy_(full[self.lo:self.hi])
        else:
            gen = torch.Generator(device=self.device).manual_seed(1000 + self.rank)
            self.table.normal_(0, 0.1, generator=gen)
        self._ef_pad = torch.zeros(P, d, **f32)
        self.E_f = self._ef_pad[: self.n_loc]
        self._y = [torch.zeros(P, d, **f32), torch.zeros(P, d, **f32)]
        self._grad_pad = torch.zeros(P, d, **f32)
        self.grad = self._grad_pad[: self.n_loc]
        self._xg = torch.empty(G * P, d, **f32)      # gathered operand of a layer
        self._e0g = torch.empty(G * P, d, **f32)     # gathered E0 (layer-1 operand, BPR regulariser rows)
        self._efg = torch.empty(G * P, d, **f32)     # gathered E_f (BPR rows)
        self._rg = torch.empty(G * P, d, **f32)      # full-size gradient scatter target
        self.loss = torch.zeros((), **f32)
        self._ws = None

    def slot(self, nodes: torch.Tensor) -> torch.Tensor:
        """Row of node ``v`` in an all-gathered [G*P, d] buffer."""
        owner = torch.bucketize(nodes, self._ends, right=True)
        return owner * self.P + (nodes - self._starts[owner])

    # ------------------------------------------------------------------------------------------------
    def forward(self) -> torch.Tensor:
        """E_f = mean_k A_hat^k E0 on the owned rows (model/lightgcn.py:58-68)."""
        K, ops = self.K, self.ops
        all_gather_blocks(ops, self._e0g, self._table_pad)
        if K == 0:
            self.E_f.copy_(self.table)
            return self.E_f
        x_full = self._e0g
        for k in range(K):
            last = k == K - 1
            y = self._y[k % 2]
            self.g_fwd.spmm(x_full, Y=None if last else y[: self.n_loc], acc_in=self.table if k == 0 else self.E_f,
                            acc_out=self.E_f, acc_div=float(K + 1) if last else 1.0, want_y=not last)
            if not last:
                all_gather_blocks(ops, self._xg, y)
                x_full = self._xg
        return self.E_f

    def backward(self, r_pad: torch.Tensor) -> torch.Tensor:
        """grad = sum_k (A_hat^T)^k r, r = dE_f/(K+1) on the owned rows (autograd of model/lightgcn.py:58-68)."""
        K, ops, n = self.K, self.ops, self.n_loc
        r = r_pad[:n]
        if K == 0:
            self.grad.copy_(r)
            return self.grad
        g_pad = r_pad
        for k in range(K):
            last = k == K - 1
            all_gather_blocks(ops, self._xg, g_pad)
            dst = self._grad_pad if last else self._y[k % 2]
            self.g_bwd.spmm(self._xg, Y=dst[:n], resid=r)
            g_pad = dst
        return self.grad

    def adam_step(self, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8) -> None:
        """Fused Adam (lgb_adam_step, torch.optim.Adam arithmetic) on the owned rows with the gradients of the last fused_step."""
        if not hasattr(self, "_adam"):
            self._adam = dict(step=0, m=torch.zeros_like(self.table), v=torch.zeros_like(self.table))
        st = self._adam
        st["step"] += 1
        self.ops.adam(self.table, self.grad, st["m"], st["v"], float(lr), float(betas[0]), float(betas[1]), float(eps), st["step"])

    def _bpr(self, iu, ip, in_, lam: float, B: int, loss=None, dEf=None, dE0=None, gscale: float = 1.0) -> None:
        a = LgbBprArgs()
        ef, e0 = self._efg.data_ptr(), self._e0g.data_ptr()
        a.uf = a.pf = a.nf = ef
        a.u0 = a.p0 = a.n0 = e0
        a.iu, a.ip, a.in_ = iu.data_ptr(), ip.data_ptr(), in_.data_ptr()
        a.B, a.B_norm, a.d, a.lambda_, a.gscale = B, B, self.d, float(lam), float(gscale)
        if dEf is not None:
            a.duf = a.dpf = a.dnf = dEf.data_ptr()
        if dE0 is not None:
            a.du0 = a.dp0 = a.dn0 = dE0.data_ptr()
        if loss is not None:
            if self._ws is None or self._ws.numel() < 2 * int(_lib.load().lgb_bpr_blocks(max(B, 1))):
                self._ws = _bpr_ws(max(B, 1), self.device)
            a.loss, a.ws = loss.data_ptr(), self._ws.data_ptr()
        _bpr_launch(a, self.device)

    @torch.no_grad()
    def fused_step(self, user_indices: torch.Tensor, pos_item_indices: torch.Tensor, neg_item_indices: torch.Tensor,
                   lambda_val: float) -> torch.Tensor:
        """One training iteration of run_pipeline_lightgcn.py:120-158 (forward, six gathers, bpr_loss, backward)."""
        ops, K, P = self.ops, self.K, self.P
        u, p, n = (_lib.i64c(t.to(self.device)) for t in (user_indices, pos_item_indices, neg_item_indices))
        B = u.numel()
        iu, ip, in_ = self.slot(u), self.slot(p + self.U), self.slot(n + self.U)
        self.forward()
        all_gather_blocks(ops, self._efg, self._ef_pad)
        mine = slice(self.rank * P, (self.rank + 1) * P)
        ops.zero(self._rg)
        if B > 0:
            self._bpr(iu, ip, in_, lambda_val, B, loss=self.loss, dEf=self._rg, gscale=1.0 / (K + 1))
        else:
            self.loss.fill_(float("nan"))         # mean over an empty batch, like the reference
        G = self.backward(self._rg[mine])
        if B > 0:                                 # + 2*lambda*E0 on the batch rows (layer-0 regulariser of bpr_loss)
            ops.zero(self._rg)
            self._bpr(iu, ip, in_, lambda_val, B, dE0=self._rg)
            ops.accumulate(self._rg[mine][: self.n_loc], G, None, 1.0, G)
        return self.loss
