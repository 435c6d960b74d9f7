"""Loss modules of the reference's `utils/custom_loss_functions.py` on the library's kernels, under the
reference's class names and call signatures (so `train.py:95-103` and `utils/training.py:61-72` dispatch on
them unchanged):

    Masked_L2_loss(regularize, regcoeff)(output, target, mask)              custom_loss_functions.py:10-46
    PowerImbalance(xymean, xystd, edgemean, edgestd)(x, edge_index, edge_attr)          :99-286
    MixedMSEPoweImbalance(xymean, xystd, edgemean, edgestd, alpha)(x, edge_index, edge_attr, y)   :289-306

Each forward runs the fused loss+gradient kernels once and hands the stored gradient to autograd, so
`loss.backward()` costs one scale.  CUDA tensors only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import nn

from . import ops
from ._lib import check, lib
from .training import masked_l2_loss_and_grad, mse_loss_and_grad


class _StoredGradLoss(torch.autograd.Function):
    """loss value + precomputed d loss / d x; backward scales the stored gradient by the incoming one."""

    @staticmethod
    def forward(ctx, x, loss, dx):
        ctx.save_for_backward(dx)
        return loss.view(())

    @staticmethod
    def backward(ctx, gout):
        (dx,) = ctx.saved_tensors
        return dx * gout, None, None


class Masked_L2_loss(nn.Module):  # noqa: N801 -- the reference's class name
    """custom_loss_functions.py:10-46: MSE over the entries where `mask == 1`, plus `regcoeff` x MSE over the rest
    when `regularize`.  Element counts stay on the device (the reference's `masked_select` synchronises twice)."""

    def __init__(self, regularize=True, regcoeff=1):
        super().__init__()
        self.regularize, self.regcoeff = regularize, regcoeff

    def forward(self, output, target, mask):
        loss, dout = masked_l2_loss_and_grad(output.detach(), target, mask, self.regularize, float(self.regcoeff))
        return _StoredGradLoss.apply(output, loss, dout)


def power_imbalance_loss_and_grad(x: torch.Tensor, edge_index: torch.Tensor, edge_attr: torch.Tensor, stats_host,
                                  need_grad: bool = True, graph: Optional[ops.PreparedGraph] = None):
    """(loss [1], d loss / d x [N, 4] or None) of the reference's `PowerImbalance.forward` (:254-286).
    `stats_host`: 12 floats = xymean[4], xystd[4], edgemean[2], edgestd[2].  `graph`: a `PreparedGraph(mode=1)` of the
    same `edge_index` / `edge_attr` to reuse (built here otherwise)."""
    dev = ops.require_cuda(x, edge_index, edge_attr)
    if x.dim() != 2 or x.size(1) < 4:
        raise ValueError(f"x must be [N, >=4] (Vm, Va, P, Q); got {tuple(x.shape)}")
    if edge_index.size(1) == 0:
        # the reference reads edge_index[0, 0] unguarded (:133)
        raise IndexError("index 0 is out of bounds for dimension 1 with size 0")
    with torch.cuda.device(dev):
        n = int(x.size(0))
        x4 = x.detach()
        if x4.dtype != torch.float32 or x4.stride(1) != 1 or x4.stride(0) % 4 != 0 or x4.data_ptr() % 16 != 0:
            x4 = x4[:, :4].float().contiguous()
        if graph is None:
            graph = ops.PreparedGraph(edge_index, edge_attr.detach().float(), n, mode=1)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        dx = torch.empty((n, 4), dtype=torch.float32, device=dev) if need_grad else None
        scratch = torch.empty(int(lib().pfn_power_imbalance_scratch_bytes(n)), dtype=torch.uint8, device=dev)
        stats = (C.c_float * 12)(*[float(v) for v in stats_host])
        check(lib().pfn_power_imbalance_fwd_bwd(
            x4.data_ptr(), int(x4.stride(0)), graph.ws.data_ptr(), n, graph.e_raw, stats, loss.data_ptr(),
            None if dx is None else dx.data_ptr(), 0 if dx is None else int(dx.stride(0)), scratch.data_ptr(),
            torch.cuda.current_stream().cuda_stream), "pfn_power_imbalance_fwd_bwd")
        if dx is not None and dx.size(1) != x.size(1):  # wider-than-4 inputs: the extra columns carry no gradient
            full = torch.zeros((n, int(x.size(1))), dtype=torch.float32, device=dev)
            full[:, :4] = dx
            dx = full
    return loss, dx


class PowerImbalance(nn.Module):
    """custom_loss_functions.py:99-286.  mean over buses of dP^2 + dQ^2, where (dP, dQ) is the mismatch between the
    predicted injections (P, Q) and the branch flows implied by the predicted voltages (Vm, Va).  The reference derives
    from PyG `MessagePassing(aggr='add', flow='target_to_source')`; here loss and gradient are three kernels on the CSR
    built by `pfn_graph_prep`."""
    base_sn = 100
    base_voltage = 345
    base_ohm = 1190.25

    def __init__(self, xymean, xystd, edgemean, edgestd, reduction='mean'):
        super().__init__()
        if xymean.shape[0] > 1:  # :119-122
            xymean = xymean[0:1]
        if xystd.shape[0] > 1:
            xystd = xystd[0:1]
        self.xymean, self.xystd, self.edgemean, self.edgestd = xymean, xystd, edgemean, edgestd
        self._stats = None

    def _stats_host(self):
        if self._stats is None:
            parts = (self.xymean.reshape(-1)[:4], self.xystd.reshape(-1)[:4], self.edgemean.reshape(-1)[:2],
                     self.edgestd.reshape(-1)[:2])
            self._stats = [float(v) for p in parts for v in p.detach().float().cpu().tolist()]
        return self._stats

    def forward(self, x, edge_index, edge_attr, graph: Optional[ops.PreparedGraph] = None):
        need_grad = torch.is_grad_enabled() and x.requires_grad
        loss, dx = power_imbalance_loss_and_grad(x, edge_index, edge_attr, self._stats_host(), need_grad, graph)
        if not need_grad:
            return loss.view(())
        return _StoredGradLoss.apply(x, loss, dx.to(x.dtype))


class MixedMSEPoweImbalance(nn.Module):
    """custom_loss_functions.py:289-306: `alpha * MSE(x, y) + (1 - alpha) * 0.020 * PowerImbalance(x, ...)`."""

    def __init__(self, xymean, xystd, edgemean, edgestd, alpha=0.5, reduction='mean'):
        super().__init__()
        assert alpha <= 1. and alpha >= 0
        self.power_imbalance = PowerImbalance(xymean, xystd, edgemean, edgestd, reduction)
        self.alpha = alpha

    def forward(self, x, edge_index, edge_attr, y, graph: Optional[ops.PreparedGraph] = None):
        power_imb_loss = self.power_imbalance(x, edge_index, edge_attr, graph)
        mse, dmse = mse_loss_and_grad(x.detach(), y)
        mse_loss = _StoredGradLoss.apply(x, mse, dmse) if (torch.is_grad_enabled() and x.requires_grad) else mse.view(())
        return self.alpha * mse_loss + (1 - self.alpha) * 0.020 * power_imb_loss
