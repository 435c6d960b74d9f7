"""The model's share of one optimisation step as the reference's `train_epoch` drives it
(utils/training.py:55-77): `data.to(device)`, `out = model(data)`, `loss = MSELoss(out, data.y)`,
`loss.backward()`, and the `loss.item()` read-back.  `optimizer.step()` stays in torch (SURVEY.md
section 8 f4) and is not part of this module.

`fused_mse_step` is the fast route for the configuration the reference's README/runs.sh use
(`--train_loss_fn mse_loss`, train.py:103): loss value and d loss/d out come from one fused kernel
(`pfn_mse_fwd_bwd`) and the backward is entered directly, so the step needs no autograd graph.
Any other loss goes through `model(data)` + torch autograd as in the reference.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops
from ._lib import check, lib
from .networks.MPN import MaskEmbdMultiMPN


def mse_loss_and_grad(out: torch.Tensor, y: torch.Tensor, total_count: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(loss, d loss / d out) for `torch.nn.MSELoss()(out, y)`; `total_count` = global element count under
    data parallelism (defaults to `out.numel()`)."""
    dev = ops.require_cuda(out, y)
    with torch.cuda.device(dev):
        out, y = out.contiguous(), y.contiguous().float()
        count = out.numel()
        inv = 1.0 / float(total_count if total_count is not None else count)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        dout = torch.empty_like(out)
        scratch = torch.empty(int(lib().pfn_mse_scratch_bytes(count)), dtype=torch.uint8, device=dev)
        check(lib().pfn_mse_fwd_bwd(out.data_ptr(), y.data_ptr(), count, inv, loss.data_ptr(), dout.data_ptr(),
                                    scratch.data_ptr(), torch.cuda.current_stream().cuda_stream), "pfn_mse_fwd_bwd")
    return loss, dout


def fused_mse_step(model: MaskEmbdMultiMPN, data, total_count: Optional[int] = None) -> torch.Tensor:
    """forward + MSE + backward; parameter `.grad`s are SET (as after `zero_grad(); loss.backward()`).
    Returns the loss as a 1-element device tensor (call `.item()` for the reference's read-back)."""
    with torch.enable_grad():
        out = model(data)
    loss, dout = mse_loss_and_grad(out.detach(), data.y, total_count)
    params = model._engine_params()
    grads = torch.autograd.grad(out, params, grad_outputs=dout, allow_unused=False)
    for p, g in zip(params, grads):
        p.grad = g
    return loss


def train_step(model: MaskEmbdMultiMPN, host_batch, device, loss: str = "mse", total_count: Optional[int] = None):
    """End-to-end step from HOST memory: H2D of the batch (pinned -> non_blocking), forward, loss, backward,
    and the device->host read of the loss (utils/training.py:56-77 minus optimizer.step)."""
    data = host_batch.to(device, non_blocking=True)
    if loss == "mse":
        val = fused_mse_step(model, data, total_count)
    else:
        raise ValueError(f"unknown loss {loss!r}")
    return float(val.item())
