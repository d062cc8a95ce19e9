"""`bpr_loss` with the reference's signature (utils/metrics_lightgcn.py:9-45) on the fused CUDA kernel.

Two entry points:
  * ``bpr_loss(u_f, u_0, p_f, p_0, n_f, n_0, lambda_val)`` -- drop-in for the reference function: six
    already-gathered [B, d] tensors, differentiable (autograd.Function around lgb_bpr, rows mode);
  * ``bpr_indexed(...)`` -- the fully fused form used by LightGCN.fused_step: row indices instead of
    gathered copies, gradients scattered with vector atomics straight into the [N, d] buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import LgbBprArgs, check, ptr, stream


def _ws(B: int, device) -> torch.Tensor:
    return torch.empty(2 * int(_lib.load().lgb_bpr_blocks(B)), dtype=torch.float32, device=device)


def _launch(args: LgbBprArgs, device) -> None:
    with torch.cuda.device(device):
        check(_lib.load().lgb_bpr(C.byref(args), stream()), "bpr")
    _lib.count_launch(2 if args.loss else 1)


class _BprRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, uf, u0, pf, p0, nf, n0, lambda_val: float):
        ts = [_lib.f32c(t) for t in (uf, u0, pf, p0, nf, n0)]
        _lib.require_cuda(*ts)
        B, d = ts[0].shape
        for t in ts:
            if tuple(t.shape) != (B, d):
                raise RuntimeError("bpr_loss: all six operands must have the same [B, d] shape")
        loss = torch.empty((), dtype=torch.float32, device=ts[0].device)
        ctx.save_for_backward(*ts)
        ctx.lambda_val = float(lambda_val)
        if B == 0:  # the reference's mean over an empty batch is nan
            return loss.fill_(float("nan"))
        a = LgbBprArgs()
        a.uf, a.u0, a.pf, a.p0, a.nf, a.n0 = (ptr(t) for t in ts)
        a.B, a.d, a.lambda_, a.gscale = B, d, float(lambda_val), 1.0
        ws = _ws(B, loss.device)
        a.loss, a.ws = ptr(loss) or loss.data_ptr(), ptr(ws)
        _launch(a, loss.device)
        return loss

    @staticmethod
    def backward(ctx, gout):
        ts = ctx.saved_tensors
        B, d = ts[0].shape
        gout = gout.to(torch.float32).contiguous()
        grads = [torch.empty_like(t) if ctx.needs_input_grad[i] else None for i, t in enumerate(ts)]
        a = LgbBprArgs()
        a.uf, a.u0, a.pf, a.p0, a.nf, a.n0 = (ptr(t) for t in ts)
        a.B, a.d, a.lambda_, a.gscale = B, d, ctx.lambda_val, 1.0
        a.gout = gout.data_ptr()
        a.duf, a.du0, a.dpf, a.dp0, a.dnf, a.dn0 = (ptr(g) for g in grads)
        if B > 0:
            _launch(a, ts[0].device)
        return (*grads, None)


def bpr_loss(users_emb_final, users_emb_0, pos_items_emb_final, pos_items_emb_0, neg_items_emb_final,
             neg_items_emb_0, lambda_val: float) -> torch.Tensor:
    """-mean(softplus(<u,p> - <u,n>)) + lambda*(|u0|^2 + |p0|^2 + |n0|^2) -- the reference's sign and scaling."""
    return _BprRows.apply(users_emb_final, users_emb_0, pos_items_emb_final, pos_items_emb_0,
                          neg_items_emb_final, neg_items_emb_0, float(lambda_val))


def bpr_indexed(E_f: torch.Tensor, E_0: torch.Tensor, num_users: int, u: torch.Tensor, p: torch.Tensor,
                n: torch.Tensor, lambda_val: float, loss: Optional[torch.Tensor] = None,
                dE_f: Optional[torch.Tensor] = None, dE_0: Optional[torch.Tensor] = None, gscale: float = 1.0,
                ws: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """Fused gather + BPR (+ scatter backward) on the full [N, d] tables.

    E_f/E_0: final / layer-0 embeddings of all N = U + I nodes (items start at row ``num_users``).
    dE_f / dE_0 (optional, [N, d], caller-zeroed or to be accumulated into) receive the gradients.
    """
    _lib.require_cuda(E_f, E_0, u, p, n)
    N, d = E_f.shape
    u, p, n = _lib.i64c(u), _lib.i64c(p), _lib.i64c(n)
    B = u.numel()
    off = num_users * d * 4
    a = LgbBprArgs()
    a.uf, a.pf, a.nf = E_f.data_ptr(), E_f.data_ptr() + off, E_f.data_ptr() + off
    a.u0, a.p0, a.n0 = E_0.data_ptr(), E_0.data_ptr() + off, E_0.data_ptr() + off
    a.iu, a.ip, a.in_ = ptr(u), ptr(p), ptr(n)
    a.B, a.d, a.lambda_, a.gscale = B, d, float(lambda_val), float(gscale)
    if dE_f is not None:
        a.duf, a.dpf, a.dnf = dE_f.data_ptr(), dE_f.data_ptr() + off, dE_f.data_ptr() + off
    if dE_0 is not None:
        a.du0, a.dp0, a.dn0 = dE_0.data_ptr(), dE_0.data_ptr() + off, dE_0.data_ptr() + off
    if loss is not None:
        if ws is None:
            ws = _ws(B, E_f.device)
        a.loss, a.ws = loss.data_ptr(), ptr(ws)
    if B == 0 and loss is None:
        return loss
    _launch(a, E_f.device)
    return loss
