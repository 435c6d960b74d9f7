"""Data parallelism over graphs: one process per GPU, graphs of a mini-batch sharded across ranks,
ONE all-reduce per step over the flat gradient buffer the backward writes.

The reference has no multi-GPU mode; graphs in a PyG batch are independent (block-diagonal
adjacency, `data.batch` unused at networks/MPN.py:532), so the only exchange the path needs is the
sum of parameter gradients.  With `MSELoss(mean)` the per-rank loss must be normalised by the
GLOBAL element count so that the reduced gradient equals the single-process gradient of the whole
batch even when ranks hold different numbers of nodes (`global_count`).
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> tuple:
    """(rank, world, local_rank); initialises torch.distributed from torchrun's environment if world > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local_rank


def allreduce_flat_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of one contiguous fp32 buffer (NCCL over NVLink on the GPU box, gloo in CPU tests)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def attach_gradient_allreduce(model, group=None) -> None:
    """Make `model`'s backward all-reduce its flat gradient buffer (all parameter gradients are views of it)
    before autograd hands them to the optimizer: one collective per step, stream-ordered after the last
    weight-gradient kernel."""
    model._grad_reducer = lambda flat: allreduce_flat_(flat, group)


def global_count(local_numel: int, device=None, group=None) -> int:
    """Sum over ranks of the number of output elements (for the mean in MSELoss)."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return int(local_numel)
    t = torch.tensor([local_numel], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def world_size(group=None) -> int:
    return dist.get_world_size(group) if dist.is_initialized() else 1


def global_mask_counts(pred_mask: torch.Tensor, group=None) -> torch.Tensor:
    """float32 [2] on `pred_mask`'s device: (number of mask != 0 entries, number of mask == 0 entries) summed over all
    ranks -- the denominators of the two means of `Masked_L2_loss` (utils/custom_loss_functions.py:29-46) under data
    parallelism.  No host synchronisation: the counts stay on the device and are read by the loss kernel."""
    n1 = (pred_mask != 0).sum()
    counts = torch.stack([n1, pred_mask.numel() - n1]).to(torch.float32)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def broadcast_parameters(model, src: int = 0, group=None) -> None:
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        for p in model.parameters():
            dist.broadcast(p.data, src=src, group=group)
