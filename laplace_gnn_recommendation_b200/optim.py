"""Fused Adam for the embedding tables (the optimizer of reference run_pipeline_lightgcn.py:103-104,157-159,178-179).

``FusedAdam`` is a ``torch.optim.Optimizer`` (so ``ExponentialLR`` and ``zero_grad`` work unchanged) whose ``step()`` is
one lgb_adam_step launch per parameter instead of torch's multi-kernel foreach path; same defaults and the same
arithmetic order as ``torch.optim.Adam`` (no weight decay / amsgrad / maximize -- the reference uses none of them)."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ptr, stream


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                _lib.require_cuda(p, p.grad)
                if not p.is_contiguous() or p.dtype != torch.float32:
                    raise RuntimeError("FusedAdam needs contiguous float32 parameters")
                g = _lib.f32c(p.grad)
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                with torch.cuda.device(p.device):
                    check(lib.lgb_adam_step(ptr(p), ptr(g), ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]), p.numel(),
                                            float(group["lr"]), float(beta1), float(beta2), float(group["eps"]), st["step"],
                                            stream()), "adam_step")
                _lib.count_launch()
        return loss
