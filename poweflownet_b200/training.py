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

from typing import Callable, Optional, Tuple

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
                            regcoeff: float = 1.0, global_counts: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(loss, d loss / d out) for the reference's `Masked_L2_loss(regularize, regcoeff)(out, y, mask)`
    (utils/custom_loss_functions.py:10-46): mean over the masked entries plus `regcoeff` x mean over the others.
    Element counts are taken on the device (no masked_select, no host sync).  Data parallel: `global_counts` = device
    float32 [2] = (masked, unmasked) element counts over ALL ranks (`parallel.global_mask_counts`); the two means are
    then global ones, the SUM over ranks of the returned loss is the whole batch's loss, and the SUM all-reduce of the
    parameter gradients equals the single-process gradient."""
    dev = ops.require_cuda(out, y, mask)
    with torch.cuda.device(dev):
        out, y = out.contiguous(), y.contiguous().float()
        mask = mask.contiguous()
        if mask.dtype != torch.int64:
            mask = mask.long()
        if mask.shape != out.shape:
            raise ValueError(f"mask shape {tuple(mask.shape)} != output shape {tuple(out.shape)}")
        if global_counts is not None:
            ops.require_cuda(global_counts)
            if global_counts.dtype != torch.float32 or global_counts.numel() != 2 or not global_counts.is_contiguous():
                raise ValueError("global_counts must be a contiguous float32 tensor of 2 elements (masked, unmasked)")
        count = out.numel()
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        dout = torch.empty_like(out)
        scratch = torch.empty(int(lib().pfn_masked_l2_scratch_bytes(count)), dtype=torch.uint8, device=dev)
        check(lib().pfn_masked_l2_fwd_bwd(out.data_ptr(), y.data_ptr(), mask.data_ptr(), count, int(bool(regularize)),
                                          float(regcoeff), None if global_counts is None else global_counts.data_ptr(),
                                          loss.data_ptr(), dout.data_ptr(), scratch.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream), "pfn_masked_l2_fwd_bwd")
    return loss, dout


def _set_grads(model: MaskEmbdMultiMPN, out: torch.Tensor, dout: torch.Tensor) -> None:
    params = model._engine_params()
    grads = torch.autograd.grad(out, params, grad_outputs=dout, allow_unused=False)
    for p, g in zip(params, grads):
        p.grad = g


def fused_masked_l2_step(model: MaskEmbdMultiMPN, data, regularize: bool = True, regcoeff: float = 1.0,
                         global_counts: Optional[torch.Tensor] = None) -> torch.Tensor:
    """forward + Masked_L2_loss + backward (the parser-default training loss, utils/training.py:61-62); parameter
    `.grad`s are SET.  Returns the loss as a 1-element device tensor (this rank's share when `global_counts` is given)."""
    with torch.enable_grad():
        out = model(data)
    loss, dout = masked_l2_loss_and_grad(out.detach(), data.y, data.pred_mask, regularize, regcoeff, global_counts)
    _set_grads(model, out, dout)
    return loss


def fused_mse_step(model: MaskEmbdMultiMPN, data, total_count: Optional[int] = None) -> torch.Tensor:
    """forward + MSE + backward; parameter `.grad`s are SET (as after `zero_grad(); loss.backward()`).
    Returns the loss as a 1-element device tensor (call `.item()` for the reference's read-back)."""
    with torch.enable_grad():
        out = model(data)
    loss, dout = mse_loss_and_grad(out.detach(), data.y, total_count)
    _set_grads(model, out, dout)
    return loss


class _SeedRing:
    """Dropout seeds for CUDA-graph replays: the captured kernels read ONE device-resident int64; every replay is
    preceded by an asynchronous copy of a fresh host-drawn seed into it.  The host runs many steps ahead of the device
    (`GraphedEpochs`, `PipelinedMSESteps`), so each queued copy gets its own pinned slot -- a single slot would be
    overwritten by a later `random_()` before the copy engine reads it, and consecutive replays would share masks.  A
    slot is reused only after the copy that read it has completed (event per slot).  Seeds come from torch's global CPU
    generator, so runs are reproducible under `torch.manual_seed`; nothing is drawn when the model's dropout rate is 0
    (as `nn.Dropout(p=0)` draws nothing in the reference)."""

    SLOTS = 64

    def __init__(self, device, model=None):
        self.model = model
        self.active = model is None or float(model.dropout.p) > 0.0
        self.host = torch.zeros(self.SLOTS, dtype=torch.int64).pin_memory()
        self.dev = torch.zeros(1, dtype=torch.int64, device=device)
        self.events = [None] * self.SLOTS
        self.k = 0

    def refresh(self) -> None:
        if not self.active:
            return
        k = self.k
        self.k = (k + 1) % self.SLOTS
        if self.events[k] is not None:
            self.events[k].synchronize()  # the copy issued SLOTS refreshes ago (long done in practice)
        else:
            self.events[k] = torch.cuda.Event()
        self.host[k:k + 1].random_()
        self.dev.copy_(self.host[k:k + 1], non_blocking=True)
        self.events[k].record(torch.cuda.current_stream(self.dev.device))


class GraphedStep:
    """forward + loss + backward of one mini-batch SHAPE captured once as a CUDA graph and replayed per step.

    `loss`: "mse" (`torch.nn.MSELoss()`, train.py:103) or "masked_l2" (`Masked_L2_loss`, the parser default,
    utils/argument_parser.py:36-37 -- `regularize` / `regcoeff` as its constructor takes them).

    The kernel launches of a step (13 on the graph-resident route, ~90 on the layer-wise one) and the host-side
    tensor-map encodes of the tensor-core GEMMs cost more CPU time than the kernels take on a B200, so the step is
    recorded once on static buffers; every call copies the new batch into those buffers, refreshes the device-resident
    dropout seed and replays.  Parameter `.grad`s are static tensors owned by the graph (as after `zero_grad();
    loss.backward()`), so an optimizer step can follow directly.  Data parallel: `total_count` = global number of output
    elements for "mse"; for "masked_l2" the global (masked, unmasked) counts are all-reduced right before every replay
    (2 floats) into a static buffer the captured loss kernel reads; the gradient all-reduce runs right after the
    replay, outside the graph.  Mini-batches of a different shape (other N / E_raw) need their own instance.

    The activation / scratch workspaces the captured kernels address are held in a pool PRIVATE to this object: no other
    forward of the same shape (an `evaluate_epoch` between replays) can take, overwrite-and-free or drop them.
    """

    FIELDS = ("x", "y", "bus_type", "pred_mask", "edge_index", "edge_attr", "batch", "ptr")

    def __init__(self, model: MaskEmbdMultiMPN, example_batch, total_count: Optional[int] = None, warmup: int = 3,
                 loss: str = "mse", regularize: bool = True, regcoeff: float = 1.0, group=None):
        if loss not in ("mse", "masked_l2"):
            raise ValueError(f"GraphedStep captures 'mse' or 'masked_l2' steps, not {loss!r}")
        dev = ops.require_cuda(*[p for p in model.parameters()])
        self.model, self.device, self.total_count = model, dev, total_count
        self.loss_kind, self.regularize, self.regcoeff, self.group = loss, bool(regularize), float(regcoeff), group
        self.static = example_batch.to(dev)
        self.static = type(self.static)(*[getattr(self.static, f).clone() for f in self.FIELDS])
        self.seeds = _SeedRing(dev, model)
        self.seed_dev = self.seeds.dev
        # an NCCL collective stays outside the graph (issued after every replay); the library's own one-shot NVLink
        # all-reduce is a plain kernel without host-side state and is captured with the step
        self.reducer = model._grad_reducer
        self.reducer_in_graph = bool(getattr(self.reducer, "graph_safe", False))
        if not self.reducer_in_graph:
            model._grad_reducer = None
        self.counts = None
        if loss == "masked_l2" and self.reducer is not None:
            self.counts = torch.zeros(2, dtype=torch.float32, device=dev)
            self._refresh_counts()
        self._pool = {}  # private workspace pool (see the class docstring)
        outer_pool = model._swap_pool(self._pool)
        model._seed_device = self.seed_dev
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(max(warmup, 1)):  # also runs every one-time cudaFuncSetAttribute / workspace allocation
                    self._refresh_seed()
                    self._step()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.loss = self._step()
            self.params = model._engine_params()
            self.grads = [p.grad for p in self.params]
        finally:
            model._seed_device = None
            model._grad_reducer = self.reducer
            model._swap_pool(outer_pool)

    def _step(self) -> torch.Tensor:
        if self.loss_kind == "mse":
            return fused_mse_step(self.model, self.static, self.total_count)
        return fused_masked_l2_step(self.model, self.static, self.regularize, self.regcoeff, self.counts)

    def _refresh_seed(self):
        self.seeds.refresh()

    def _refresh_counts(self):
        from .parallel import global_mask_counts
        self.counts.copy_(global_mask_counts(self.static.pred_mask, self.group))

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
        if self.counts is not None:
            self._refresh_counts()
        self.graph.replay()
        for p, g in zip(self.params, self.grads):
            p.grad = g
        if self.reducer is not None and not self.reducer_in_graph:
            flat = self.grads[0]._base if self.grads[0]._base is not None else None
            if flat is not None and flat.numel() >= sum(g.numel() for g in self.grads):
                self.reducer(flat)
            else:
                for g in self.grads:
                    self.reducer(g)
        return self.loss


class GraphedMSEStep(GraphedStep):
    """`GraphedStep(loss="mse")` under its round-1 name."""

    def __init__(self, model: MaskEmbdMultiMPN, example_batch, total_count: Optional[int] = None, warmup: int = 3):
        super().__init__(model, example_batch, total_count, warmup, loss="mse")


class PipelinedMSESteps:
    """Two `GraphedMSEStep`s over two sets of static input buffers: the host->device copy of batch i+1 runs on a side
    stream while batch i computes (what a prefetching loader with `pin_memory=True` does for the reference's
    `data.to(device)` at utils/training.py:56).  Usage::

        pipe.prefetch(batch_0)
        for i in range(n):
            loss = pipe.step()                # queue forward + MSE + backward of batch_i (already on the device)
            pipe.prefetch(batch_{i+1})        # asynchronous copy on the side stream, overlaps the step
            value = float(loss.item())        # the device->host read of batch_i's loss
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

    def step_async(self) -> "Callable[[], float]":
        """`step()` whose loss travels to the host asynchronously: the 4-byte device->host copy is queued behind the replay
        into a pinned slot of its own, and the returned callable waits for THAT copy only.  Reading the loss of step i
        after step i + 1 has been queued keeps the GPU busy across the read-back (an asynchronous logger); the values are
        the ones `step().item()` returns."""
        loss = self.step()
        if not hasattr(self, "_loss_host"):
            self._loss_host = torch.zeros(8, dtype=torch.float32).pin_memory()
            self._loss_events = [torch.cuda.Event() for _ in range(8)]
            self._loss_k = 0
        k = self._loss_k
        self._loss_k = (k + 1) % 8
        slot = self._loss_host[k:k + 1]
        slot.copy_(loss.reshape(1), non_blocking=True)
        ev = self._loss_events[k]
        ev.record(torch.cuda.current_stream(loss.device))

        def read() -> float:
            ev.synchronize()
            return float(slot[0])
        return read

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
    """Whole optimisation epochs with (almost) no host work per step: the sample ids of an epoch go to the GPU once,
    `pfn_batch_assemble` writes every mini-batch straight INTO the static input buffers of a captured forward + loss +
    backward graph (`GraphedStep`), the graph is replayed, and the optimizer (meant for `optim.FusedAdamW`: one launch,
    pointer tables cached because the gradient buffers are static) follows.  The loss sum stays on the device; one
    read-back per epoch.  The `train_epoch` loop of utils/training.py:30-80 for `--train_loss_fn mse_loss` or the parser
    default `masked_l2` (incomplete last batches are dropped).

    Datasets that mix graph sizes (`--case mixed`, datasets/PowerFlowData.py:67-70): the shape of a mini-batch is set by
    how many samples of each case it drew, so steps are BUCKETED BY BATCH SHAPE -- the first batch of a shape captures
    its own graph (at most `max_graphs` are kept; further shapes run the same step eagerly), later ones replay it.
    """

    def __init__(self, model: MaskEmbdMultiMPN, dataset, batch_size: int, optimizer, total_count: Optional[int] = None,
                 rank: int = 0, world: int = 1, loss: str = "mse", regularize: bool = True, regcoeff: float = 1.0,
                 max_graphs: int = 64):
        """Data parallel: pass `rank` / `world` (every rank must seed `run_epoch`'s generator alike); single-case
        datasets with loss "mse" also need `total_count = world * batch_size * nodes_per_graph * output_dim` (for mixed
        datasets the global count is all-reduced per step).  The model's gradient all-reduce
        (`parallel.attach_gradient_allreduce`) runs after every replay."""
        if batch_size > len(dataset):
            raise ValueError("batch_size exceeds the dataset")
        self.model, self.dataset, self.batch_size, self.optimizer = model, dataset, int(batch_size), optimizer
        self.rank, self.world, self.total_count = int(rank), int(world), total_count
        self.loss_kind, self.regularize, self.regcoeff, self.max_graphs = loss, regularize, regcoeff, int(max_graphs)
        self.uniform = len(set(int(v) for v in dataset._n)) == 1 and len(set(int(v) for v in dataset._e)) == 1
        model.train()
        self.steps = {}
        self.eager_steps = 0
        if self.uniform:
            self.step = self._step_for(dataset.batch(list(range(batch_size))))
            self.n_attr = len(self.step.static)

    def _shape_key(self, ids_host):
        ds = self.dataset
        import numpy as np
        case_of = np.searchsorted(ds._first, ids_host, side="right") - 1
        return int(ds._n[case_of].sum()), int(ds._e[case_of].sum())

    def _step_for(self, example) -> "GraphedStep":
        key = (int(example.x.size(0)), int(example.edge_index.size(1)))
        step = self.steps.get(key)
        if step is None:
            total = self.total_count
            if not self.uniform and self.loss_kind == "mse" and self.world > 1:
                raise NotImplementedError("mixed-size datasets under data parallelism: use loss='masked_l2' (global counts "
                                          "are all-reduced per step) or the eager `train_epoch`")
            step = GraphedStep(self.model, example, total, loss=self.loss_kind, regularize=self.regularize,
                               regcoeff=self.regcoeff)
            self.steps[key] = step
        return step

    def run_epoch(self, shuffle: bool = True, generator: Optional[torch.Generator] = None) -> float:
        from .datasets import epoch_batches
        ds, bs = self.dataset, self.batch_size
        batches = epoch_batches(len(ds), bs, shuffle, generator, True, self.rank, self.world)
        steps = len(batches)
        total = None
        if steps:
            order = torch.cat(batches)
            dev = next(self.model.parameters()).device
            order_dev = order.to(dev, non_blocking=True)  # the ids of the whole epoch travel once
            order_host = order.numpy()
        self.model.train()
        for k in range(steps):
            ids_h, ids_d = order_host[k * bs:(k + 1) * bs], order_dev[k * bs:(k + 1) * bs]
            if self.uniform:
                step = self.step
            else:
                key = self._shape_key(ids_h)
                step = self.steps.get(key)
                if step is None and len(self.steps) < self.max_graphs:
                    step = self._step_for(ds.batch(ids_h, ids_device=ids_d))
            if step is not None:
                ds.batch(ids_h, ids_device=ids_d, out=step.static)
                loss = step(None)  # replay; parameter .grads are the graph's static buffers
            else:  # more distinct shapes than `max_graphs`: the same step, launched eagerly
                self.eager_steps += 1
                data = ds.batch(ids_h, ids_device=ids_d)
                loss = (fused_mse_step(self.model, data, self.total_count) if self.loss_kind == "mse" else
                        fused_masked_l2_step(self.model, data, self.regularize, self.regcoeff))
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
    from . import parallel
    model = model.to(device)
    total, num_samples = None, 0
    model.train()
    # data parallel (`parallel.attach_gradient_allreduce` + `loader(rank=, world=)`): the means of the fused losses are
    # taken over the GLOBAL element counts, so the SUM all-reduce of the gradients equals the single-process gradient
    # (also when ranks hold unequal last batches) and the returned loss is this rank's share of the global one.  Losses
    # that go through autograd below keep per-rank means: their reduced gradient is divided by the world size instead.
    dp = model._grad_reducer is not None and parallel.world_size() > 1
    for data in loader:
        data = data.to(device)
        optimizer.zero_grad()
        if isinstance(loss_fn, torch.nn.MSELoss) and loss_fn.reduction == "mean":
            count = parallel.global_count(data.y.numel(), device) if dp else None
            loss = fused_mse_step(model, data, count).view(())
        elif isinstance(loss_fn, Masked_L2_loss):
            counts = parallel.global_mask_counts(data.pred_mask) if dp else None
            loss = fused_masked_l2_step(model, data, loss_fn.regularize, float(loss_fn.regcoeff), counts).view(())
        else:
            if dp:
                raise NotImplementedError(
                    "data-parallel train_epoch supports MSELoss and Masked_L2_loss (global element counts); other losses "
                    "normalise per rank -- divide the reduced gradient by the world size in your own loop")
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
