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
        self._tables = {}  # param-group index -> (data_ptr key, ctypes pointer tables, element counts)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            by_step = {}
            for p in group["params"]:
                g = p.grad
                if g is None:
                    continue
                st = self.state[p]
                if len(st) == 0:
                    self._check(p, g)
                    st["step"] = 0.0  # a Python number: torch's AdamW converts it on load (`__setstate__`)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                k = st["step"] + 1  # stays a tensor when the state came from torch.optim.AdamW
                st["step"] = k
                if not g.is_contiguous():
                    g = g.contiguous()
                by_step.setdefault(int(k), []).append((p, g, st["exp_avg"], st["exp_avg_sq"]))
            # parameters that joined later carry their own step count: one launch per distinct count keeps the bias
            # corrections right (normally there is exactly one)
            for k, items in by_step.items():
                self._launch(gi, items, group, k)
        return loss

    @staticmethod
    def _check(p, g):
        if not p.is_cuda or p.dtype != torch.float32 or g.dtype != torch.float32:
            raise RuntimeError("FusedAdamW: CUDA float32 parameters and gradients only (no CPU fallback)")
        if g.is_sparse:
            raise RuntimeError("FusedAdamW does not support sparse gradients")
        if not p.is_contiguous():
            raise RuntimeError("FusedAdamW: parameters must be contiguous")

    def _launch(self, gi, items, group, step):
        n = len(items)
        key = tuple(t.data_ptr() for it in items for t in it)
        cached = self._tables.get(gi)
        if cached is None or cached[0] != key:
            # pointer tables are rebuilt only when a tensor moved (with static gradient buffers -- CUDA-graph replays,
            # `zero_grad(set_to_none=False)` -- never)
            for p, g, _, _ in items:
                self._check(p, g)
            cols = list(zip(*items))
            tabs = [(C.c_void_p * n)(*[t.data_ptr() for t in col]) for col in cols]
            numel = (C.c_int64 * n)(*[t.numel() for t in cols[0]])
            cached = (key, tabs, numel)
            self._tables[gi] = cached
        _, tabs, numel = cached
        beta1, beta2 = group["betas"]
        with torch.cuda.device(items[0][0].device):
            check(lib().pfn_adamw_step(n, tabs[0], tabs[1], tabs[2], tabs[3], numel, float(group["lr"]), float(beta1),
                                       float(beta2), float(group["eps"]), float(group["weight_decay"]), int(step),
                                       torch.cuda.current_stream().cuda_stream), "pfn_adamw_step")
