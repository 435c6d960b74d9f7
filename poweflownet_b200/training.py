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


def masked_l2_loss_and_grad(out: torch.Tensor, y: torch.Tensor, mask: torch.Tensor, regularize: bool = True,
                            regcoeff: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """(loss, d loss / d out) for the reference's `Masked_L2_loss(regularize, regcoeff)(out, y, mask)`
    (utils/custom_loss_functions.py:10-46): mean over the masked entries plus `regcoeff` x mean over the others.
    Element counts are taken on the device (no masked_select, no host sync).  Single process only: under data
    parallelism the two means need global counts -- use `model(data)` + torch autograd there."""
    dev = ops.require_cuda(out, y, mask)
    with torch.cuda.device(dev):
        out, y = out.contiguous(), y.contiguous().float()
        mask = mask.contiguous()
        if mask.dtype != torch.int64:
            mask = mask.long()
        if mask.shape != out.shape:
            raise ValueError(f"mask shape {tuple(mask.shape)} != output shape {tuple(out.shape)}")
        count = out.numel()
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        dout = torch.empty_like(out)
        scratch = torch.empty(int(lib().pfn_masked_l2_scratch_bytes(count)), dtype=torch.uint8, device=dev)
        check(lib().pfn_masked_l2_fwd_bwd(out.data_ptr(), y.data_ptr(), mask.data_ptr(), count, int(bool(regularize)),
                                          float(regcoeff), loss.data_ptr(), dout.data_ptr(), scratch.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream), "pfn_masked_l2_fwd_bwd")
    return loss, dout


def fused_masked_l2_step(model: MaskEmbdMultiMPN, data, regularize: bool = True, regcoeff: float = 1.0) -> torch.Tensor:
    """forward + Masked_L2_loss + backward (the parser-default training loss, utils/training.py:61-62); parameter
    `.grad`s are SET.  Returns the loss as a 1-element device tensor."""
    with torch.enable_grad():
        out = model(data)
    loss, dout = masked_l2_loss_and_grad(out.detach(), data.y, data.pred_mask, regularize, regcoeff)
    params = model._engine_params()
    grads = torch.autograd.grad(out, params, grad_outputs=dout, allow_unused=False)
    for p, g in zip(params, grads):
        p.grad = g
    return loss


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


class GraphedMSEStep:
    """forward + MSE + backward of one mini-batch SHAPE captured once as a CUDA graph and replayed per step.

    The kernel launches of a step (13 on the graph-resident route, ~90 on the layer-wise one) and the host-side
    tensor-map encodes of the tensor-core GEMMs cost more
    CPU time than the kernels take on a B200, so the step is recorded once on static buffers; every call copies
    the new batch into those buffers, refreshes the device-resident dropout seed and replays.  Parameter `.grad`s
    are static tensors owned by the graph (as after `zero_grad(); loss.backward()`), so an optimizer step can
    follow directly.  The data-parallel gradient all-reduce runs right after the replay, outside the graph.
    Mini-batches of a different shape (other N / E_raw) need their own instance.
    """

    FIELDS = ("x", "y", "bus_type", "pred_mask", "edge_index", "edge_attr", "batch", "ptr")

    def __init__(self, model: MaskEmbdMultiMPN, example_batch, total_count: Optional[int] = None, warmup: int = 3):
        dev = ops.require_cuda(*[p for p in model.parameters()])
        self.model, self.device, self.total_count = model, dev, total_count
        self.static = example_batch.to(dev)
        self.static = type(self.static)(*[getattr(self.static, f).clone() for f in self.FIELDS])
        self.seed_host = torch.zeros(1, dtype=torch.int64).pin_memory()
        self.seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.reducer, model._grad_reducer = model._grad_reducer, None  # the collective stays outside the graph
        model._seed_device = self.seed_dev
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(max(warmup, 1)):  # also runs every one-time cudaFuncSetAttribute / workspace allocation
                    self._refresh_seed()
                    fused_mse_step(model, self.static, total_count)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.loss = fused_mse_step(model, self.static, total_count)
            self.params = model._engine_params()
            self.grads = [p.grad for p in self.params]
        finally:
            model._seed_device = None
            model._grad_reducer = self.reducer

    def _refresh_seed(self):
        self.seed_host.random_()
        self.seed_dev.copy_(self.seed_host, non_blocking=True)

    def load(self, batch) -> None:
        """Copy a batch of the captured shape (host pinned or device) into the static input buffers."""
        for f in self.FIELDS:
            dst, src = getattr(self.static, f), getattr(batch, f)
            if dst.shape != src.shape:
                raise ValueError(f"batch.{f} has shape {tuple(src.shape)}, the captured graph expects {tuple(dst.shape)}")
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)

    def __call__(self, batch=None) -> torch.Tensor:
        if batch is not None:
            self.load(batch)
        self._refresh_seed()
        self.graph.replay()
        for p, g in zip(self.params, self.grads):
            p.grad = g
        if self.reducer is not None:
            flat = self.grads[0]._base if self.grads[0]._base is not None else None
            if flat is not None and flat.numel() == sum(g.numel() for g in self.grads):
                self.reducer(flat)
            else:
                for g in self.grads:
                    self.reducer(g)
        return self.loss


class PipelinedMSESteps:
    """Two `GraphedMSEStep`s over two sets of static input buffers: the host->device copy of batch i+1 runs on a side
    stream while batch i computes (what a prefetching loader with `pin_memory=True` does for the reference's
    `data.to(device)` at utils/training.py:56).  Usage::

        pipe.prefetch(batch_0)
        for i in range(n):
            pipe.prefetch(batch_{i+1})        # asynchronous, overlaps the step below
            loss = float(pipe.step().item())  # forward + MSE + backward of batch_i, then the device->host read
    """

    def __init__(self, model: MaskEmbdMultiMPN, example_batch, total_count: Optional[int] = None):
        self.steps = [GraphedMSEStep(model, example_batch, total_count) for _ in range(2)]
        dev = self.steps[0].device
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.done = [torch.cuda.Event() for _ in range(2)]
        self._in, self._run, self._pending = 0, 0, 0

    def prefetch(self, batch) -> None:
        if self._pending >= 2:
            raise RuntimeError("both buffer sets hold batches that have not been stepped yet")
        k = self._in
        self._in ^= 1
        self._pending += 1
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.done[k])  # the replay that last read this buffer set has finished
            self.steps[k].load(batch)
            self.ready[k].record(self.copy_stream)

    def step(self) -> torch.Tensor:
        if self._pending <= 0:
            raise RuntimeError("step() without a prefetched batch")
        k = self._run
        self._run ^= 1
        self._pending -= 1
        cur = torch.cuda.current_stream(self.steps[k].device)
        cur.wait_event(self.ready[k])
        loss = self.steps[k](None)
        self.done[k].record(cur)
        return loss


class GraphedEpochs:
    """Whole optimisation epochs at a fixed mini-batch shape with (almost) no host work per step: the sample ids of an
    epoch go to the GPU once, `pfn_batch_assemble` writes every mini-batch straight INTO the static input buffers of a
    captured forward + MSE + backward graph (`GraphedMSEStep`), the graph is replayed, and the optimizer (meant for
    `optim.FusedAdamW`: one launch, pointer tables cached because the gradient buffers are static) follows.  The loss
    sum stays on the device; one read-back per epoch.  The `train_epoch` loop of utils/training.py:30-80 for
    `--train_loss_fn mse_loss` on a single-case dataset (all graphs of one size; incomplete last batches are dropped)."""

    def __init__(self, model: MaskEmbdMultiMPN, dataset, batch_size: int, optimizer, total_count: Optional[int] = None,
                 rank: int = 0, world: int = 1):
        """Data parallel: pass `rank` / `world` (every rank must seed `run_epoch`'s generator alike) and
        `total_count = world * batch_size * nodes_per_graph * output_dim`; the model's gradient all-reduce
        (`parallel.attach_gradient_allreduce`) runs after every replay."""
        if batch_size > len(dataset):
            raise ValueError("batch_size exceeds the dataset")
        self.model, self.dataset, self.batch_size, self.optimizer = model, dataset, int(batch_size), optimizer
        self.rank, self.world = int(rank), int(world)
        model.train()
        self.step = GraphedMSEStep(model, dataset.batch(list(range(batch_size))), total_count)
        self.n_attr = len(self.step.static)

    def run_epoch(self, shuffle: bool = True, generator: Optional[torch.Generator] = None) -> float:
        from .datasets import epoch_batches
        ds, bs = self.dataset, self.batch_size
        batches = epoch_batches(len(ds), bs, shuffle, generator, True, self.rank, self.world)
        steps = len(batches)
        total = None
        if steps:
            order = torch.cat(batches)
            order_dev = order.to(self.step.device, non_blocking=True)  # the ids of the whole epoch travel once
            order_host = order.numpy()
        self.model.train()
        for k in range(steps):
            ds.batch(order_host[k * bs:(k + 1) * bs], ids_device=order_dev[k * bs:(k + 1) * bs], out=self.step.static)
            loss = self.step(None)  # replay; parameter .grads are the graph's static buffers
            self.optimizer.step()
            total = loss.clone() if total is None else total + loss  # `loss` is a static tensor of the graph
        if total is None:
            raise ZeroDivisionError("GraphedEpochs.run_epoch: no full batch")
        return float(total.item()) / steps


def train_step(model: MaskEmbdMultiMPN, host_batch, device, loss: str = "mse", total_count: Optional[int] = None):
    """End-to-end step from HOST memory: H2D of the batch (pinned -> non_blocking), forward, loss, backward,
    and the device->host read of the loss (utils/training.py:56-77 minus optimizer.step)."""
    data = host_batch.to(device, non_blocking=True)
    if loss == "mse":
        val = fused_mse_step(model, data, total_count)
    else:
        raise ValueError(f"unknown loss {loss!r}")
    return float(val.item())


def train_epoch(model: MaskEmbdMultiMPN, loader, loss_fn, optimizer, device) -> float:
    """`utils/training.py:30-80` with the same arguments and the same loss dispatch (:61-72), on the library's kernels:

    * `torch.nn.MSELoss()` (train.py:103) and `Masked_L2_loss` (the parser default) take the fused
      forward + loss + backward steps above (no autograd tape);
    * `PowerImbalance` gets `out * pred_mask + x * (1 - pred_mask)` (:63-68), `MixedMSEPoweImbalance` gets `out` and `y`
      (:69-70) -- both through `model(data)` + `loss.backward()`;
    * any other callable is applied as `loss_fn(out, data.y)` (:72).

    The reference reads `loss.item()` every step (:77, one stream synchronisation per batch); here the running sum stays
    on the device and is read once per epoch.  The return value is the reference's: sum(loss * len(data)) /
    sum(len(data)), where `len(data)` is the number of stored attributes of the batch (PyG `BaseData.__len__`), i.e. an
    unweighted mean over batches.  `loader`: any iterable of batches (`datasets.PowerFlowData.loader`, a PyG DataLoader)."""
    from .losses import Masked_L2_loss, MixedMSEPoweImbalance, PowerImbalance
    model = model.to(device)
    total, num_samples = None, 0
    model.train()
    for data in loader:
        data = data.to(device)
        optimizer.zero_grad()
        if isinstance(loss_fn, torch.nn.MSELoss) and loss_fn.reduction == "mean":
            loss = fused_mse_step(model, data).view(())
        elif isinstance(loss_fn, Masked_L2_loss):
            loss = fused_masked_l2_step(model, data, loss_fn.regularize, float(loss_fn.regcoeff)).view(())
        else:
            out = model(data)
            if isinstance(loss_fn, PowerImbalance):
                # "have to mask out the non-predicted values, otherwise the network can learn to predict full-zeros"
                masked_out = out * data.pred_mask + data.x * (1 - data.pred_mask)
                loss = loss_fn(masked_out, data.edge_index, data.edge_attr)
            elif isinstance(loss_fn, MixedMSEPoweImbalance):
                loss = loss_fn(out, data.edge_index, data.edge_attr, data.y)
            else:
                loss = loss_fn(out, data.y)
            loss.backward()
        optimizer.step()
        num_samples += len(data)
        weighted = loss.detach() * len(data)
        total = weighted if total is None else total + weighted
    if total is None:
        raise ZeroDivisionError("train_epoch: empty loader")  # the reference divides by num_samples == 0
    return float(total.item()) / num_samples


@torch.no_grad()
def evaluate_epoch(model: MaskEmbdMultiMPN, loader, loss_fn, device="cuda", pre_loss_fn=None) -> float:
    """`utils/evaluation.py:53-104` with the same arguments and dispatch: `model.eval()`, `out = model(data)`, the loss
    picked by the type of `loss_fn` (including the reference's `out*pred_mask + pred_mask*(1-pred_mask)` for
    `PowerImbalance`, :85-90 -- the second term is identically zero, kept as written), mean weighted by `len(data)`.
    The running sum stays on the device; one read-back per epoch instead of one per batch (:101)."""
    from .losses import Masked_L2_loss, MixedMSEPoweImbalance, PowerImbalance
    pre_loss_fn = pre_loss_fn or (lambda x: x)
    model.eval()
    total, num_samples = None, 0
    for data in loader:
        data = data.to(device)
        out = model(data)
        if isinstance(loss_fn, Masked_L2_loss):
            out = pre_loss_fn(out)
            target = pre_loss_fn(data.y)
            loss = loss_fn(out, target, data.pred_mask)
        elif isinstance(loss_fn, PowerImbalance):
            masked_out = out * data.pred_mask + data.pred_mask * (1 - data.pred_mask)
            masked_out = pre_loss_fn(masked_out)
            loss = loss_fn(masked_out, data.edge_index, data.edge_attr)
        elif isinstance(loss_fn, MixedMSEPoweImbalance):
            out = pre_loss_fn(out)
            loss = loss_fn(out, data.edge_index, data.edge_attr, data.y)
        else:
            out, target = pre_loss_fn(out), pre_loss_fn(data.y)
            loss = loss_fn(out, target)
        num_samples += len(data)
        weighted = loss.detach().reshape(()) * len(data)
        total = weighted if total is None else total + weighted
    if total is None:
        raise ZeroDivisionError("evaluate_epoch: empty loader")
    return float(total.item()) / num_samples
