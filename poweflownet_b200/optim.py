"""`torch.optim.AdamW` (train.py:123) as ONE kernel launch per step over every parameter tensor (SURVEY.md section 8
f4).  Same constructor defaults, same `param_groups` (so `OneCycleLR`, train.py:129/145, drives `group['lr']` as
usual) and the same per-parameter state keys (`step`, `exp_avg`, `exp_avg_sq`) as torch's optimizer, so a
`state_dict()` saved by either loads into the other.  CUDA fp32 parameters only; `amsgrad` / `maximize` are the
reference's defaults (False) and not offered.
"""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import check, lib


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        if lr < 0.0 or eps < 0.0 or weight_decay < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"invalid AdamW hyper-parameters: lr={lr} betas={betas} eps={eps} weight_decay={weight_decay}")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            ps, gs, ms, vs = [], [], [], []
            step = None
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                    raise RuntimeError("FusedAdamW: CUDA float32 parameters and gradients only (no CPU fallback)")
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdamW does not support sparse gradients")
                if not p.is_contiguous():
                    raise RuntimeError("FusedAdamW: parameters must be contiguous")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)  # host scalar tensor, as torch keeps it
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                k = int(st["step"].item())
                if step is None:
                    step = k
                elif step != k:  # parameters that joined later: their own launch keeps the bias corrections right
                    self._launch([p], [p.grad.contiguous()], [st["exp_avg"]], [st["exp_avg_sq"]], group, k)
                    continue
                ps.append(p)
                gs.append(p.grad if p.grad.is_contiguous() else p.grad.contiguous())
                ms.append(st["exp_avg"])
                vs.append(st["exp_avg_sq"])
            if ps:
                self._launch(ps, gs, ms, vs, group, step)
        return loss

    @staticmethod
    def _launch(ps, gs, ms, vs, group, step):
        n = len(ps)
        ptrs = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])  # noqa: E731
        numel = (C.c_int64 * n)(*[t.numel() for t in ps])
        beta1, beta2 = group["betas"]
        with torch.cuda.device(ps[0].device):
            check(lib().pfn_adamw_step(n, ptrs(ps), ptrs(gs), ptrs(ms), ptrs(vs), numel, float(group["lr"]), float(beta1),
                                       float(beta2), float(group["eps"]), float(group["weight_decay"]), int(step),
                                       torch.cuda.current_stream().cuda_stream), "pfn_adamw_step")
