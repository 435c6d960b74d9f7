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


class OneShotAllReduce:
    """In-place SUM all-reduce of one flat fp32 CUDA buffer as ONE kernel of libpfn_b200.so over NVLink peer memory
    (`pfn_allreduce_peer`, csrc/allreduce.cu).  Up to `TWO_SHOT_FROM - 1` ranks: every rank pushes its buffer into a
    per-source slot of every peer's symmetric receive buffer, one flag exchange, local sum in rank order.  More ranks:
    reduce-scatter + all-gather through the same symmetric buffers (two flag exchanges, (world-1)/world x 2n floats out per
    rank instead of (world-1) x n).  No host synchronisation and no host-side state per call, so the kernel can be CAPTURED
    inside the CUDA graph of a training step (`graph_safe`), unlike an NCCL call issued after the replay.  Meant for the
    latency-bound case (the 1.4 MB gradient of configs/standard.json); large buffers stay on NCCL
    (`attach_gradient_allreduce` picks by size).  Construction is collective (symmetric-memory rendezvous)."""

    graph_safe = True
    CTAS = 64
    TWO_SHOT_FROM = 5

    def __init__(self, n_floats: int, device, group=None, two_shot: Optional[bool] = None):
        import torch.distributed._symmetric_memory as symm
        from ._lib import lib
        lib()  # fail early if the library is missing
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.two_shot = bool(two_shot) if two_shot is not None else self.world >= self.TWO_SHOT_FROM
        self.pad_multiple = 4 * self.world
        self.n = (int(n_floats) + self.pad_multiple - 1) // self.pad_multiple * self.pad_multiple
        self.device = torch.device(device)
        recv_floats = 2 * self.n if self.two_shot else 2 * self.world * self.n
        self.recv = symm.empty(recv_floats, dtype=torch.float32, device=self.device)
        self.res = symm.empty(2 * self.n, dtype=torch.float32, device=self.device)
        self.sig = symm.empty(2 * self.CTAS * self.world, dtype=torch.int32, device=self.device)
        self.recv.zero_()
        self.res.zero_()
        self.sig.zero_()
        handles = [symm.rendezvous(t, self.group) for t in (self.recv, self.res, self.sig)]
        self._handles = handles  # keep the mappings alive
        ptrs = lambda h: torch.tensor([int(p) for p in h.buffer_ptrs], dtype=torch.int64, device=self.device)  # noqa: E731
        self.peer_recv, self.peer_res, self.peer_sig = (ptrs(h) for h in handles)
        self.epochs = torch.zeros(self.CTAS, dtype=torch.int32, device=self.device)
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)  # every rank has zeroed its flags before anyone signals

    def __call__(self, flat: torch.Tensor) -> torch.Tensor:
        from ._lib import check, lib
        if not (flat.is_cuda and flat.dtype == torch.float32 and flat.is_contiguous() and flat.data_ptr() % 16 == 0):
            raise ValueError("OneShotAllReduce: contiguous 16-byte-aligned float32 CUDA buffer expected")
        n = flat.numel()
        if n != self.n:
            raise ValueError(f"OneShotAllReduce: buffer of {n} floats, built for {self.n} (pad to a multiple of {self.pad_multiple})")
        check(lib().pfn_allreduce_peer(flat.data_ptr(), self.peer_recv.data_ptr(), self.peer_res.data_ptr(), self.peer_sig.data_ptr(),
                                       self.epochs.data_ptr(), self.rank, self.world, n, self.CTAS, int(self.two_shot),
                                       torch.cuda.current_stream(flat.device).cuda_stream), "pfn_allreduce_peer")
        return flat


ONE_SHOT_MAX_BYTES = 4 << 20  # beyond this the (world - 1) x traffic of the one-shot exchange loses to NCCL's ring


def attach_gradient_allreduce(model, group=None, one_shot: Optional[bool] = None) -> None:
    """Make `model`'s backward all-reduce its flat gradient buffer (all parameter gradients are views of it)
    before autograd hands them to the optimizer: one collective per step, stream-ordered after the last
    weight-gradient kernel.  Small gradients (<= ONE_SHOT_MAX_BYTES, CUDA, NCCL world > 1) take the library's own
    one-shot NVLink kernel (`OneShotAllReduce`; `one_shot=False` forces NCCL, `True` insists); anything else NCCL / gloo.
    If the symmetric-memory rendezvous is not available on the system the NCCL path is kept (a warning says so)."""
    params = list(model.parameters())
    n = sum(p.numel() for p in params)
    dev = params[0].device if params else torch.device("cpu")
    want = one_shot if one_shot is not None else (4 * n <= ONE_SHOT_MAX_BYTES and os.environ.get("PFN_ONE_SHOT_ALLREDUCE", "1") != "0")
    if want and dev.type == "cuda" and dist.is_initialized() and dist.get_world_size(group) > 1 and dist.get_backend(group) == "nccl":
        try:
            reducer = OneShotAllReduce(n, dev, group)
            # every rank must take the same path: agree on success
            ok = torch.ones(1, dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 1:
                model._grad_reducer = reducer
                return
        except Exception as exc:  # noqa: BLE001 -- symmetric memory unsupported here: fall back, loudly
            import warnings
            warnings.warn(f"one-shot NVLink all-reduce unavailable ({type(exc).__name__}: {exc}); using NCCL")
            if one_shot:
                raise
            bad = torch.zeros(1, dtype=torch.int32, device=dev)
            dist.all_reduce(bad, op=dist.ReduceOp.MIN, group=group)
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
