"""Device-resident counterpart of the reference's `datasets/PowerFlowData.py` + PyG `DataLoader` (SURVEY.md
section 8 f2): the raw arrays of a split live in HBM in the reference's file format and every mini-batch is assembled
on the GPU by one kernel (`pfn_batch_assemble`) from a list of sample ids -- no per-sample `Data` objects, no host
collation, no per-step H2D copy of features (only the ids travel: 8 bytes per graph).

    ds = PowerFlowData(root, case='118v2', split=[.5, .2, .3], task='train', device='cuda')   # same arguments
    for data in ds.loader(batch_size=128, shuffle=True):      # GraphBatch in the PyG `Batch` layout
        out = model(data)

What is kept from the reference: raw file names and layout (:58-61,142-147,178-179), the `torch.split` cut by
`[int(S * f) for f in split]` (:183-187), per-task concatenation of the cases of `--case mixed` (:151-155,214),
statistics over the split unless handed in (:99-108,126-138), `get_data_dimensions()`, `get_data_means_stds()`,
`len()`.  Statistics are computed once on the host with the reference's torch expressions.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from ._lib import DatasetCase, check, lib
from .data import GraphBatch

SPLIT_ORDER = {"train": 0, "val": 1, "test": 2}
MIXED_CASES = ["118v2", "14v2"]  # datasets/PowerFlowData.py:67-70


def epoch_batches(n_samples: int, batch_size: int, shuffle: bool = False, generator: Optional[torch.Generator] = None,
                  drop_last: bool = False, rank: int = 0, world: int = 1) -> List[torch.Tensor]:
    """Sample ids of one pass, cut into mini-batches.  With `world > 1` the permutation is cut into groups of `world`
    consecutive mini-batches and rank r takes the r-th of each group, so the ranks see disjoint samples, the same number
    of steps (a trailing group that cannot serve every rank is dropped) and, when `drop_last`, the same batch shape --
    what the per-step gradient all-reduce needs."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside [0, {world})")
    order = torch.randperm(n_samples, generator=generator) if shuffle else torch.arange(n_samples)
    batches = [order[lo:lo + batch_size] for lo in range(0, n_samples, batch_size)]
    if drop_last and batches and batches[-1].numel() < batch_size:
        batches.pop()
    if world > 1:
        batches = batches[:len(batches) // world * world][rank::world]
    return batches


class PowerFlowData:
    def __init__(self, root: Optional[str] = None, case: str = "14", split: Optional[Sequence[float]] = None,
                 task: str = "train", normalize: bool = True, xymean=None, xystd=None, edgemean=None, edgestd=None,
                 device="cuda", raw: Optional[Sequence[Tuple[torch.Tensor, torch.Tensor]]] = None,
                 random_bus_type: bool = False):
        """`raw`: [(edge_features [S, E, 4], node_features [S, n, 6]), ...] instead of the files under `root/raw`."""
        assert split is not None and len(split) == 3
        assert task in SPLIT_ORDER
        self.case, self.split, self.task, self.normalize = case, list(split), task, normalize
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("poweflownet_b200.datasets.PowerFlowData keeps the dataset in GPU memory (no CPU path)")
        self.random_bus_type = random_bus_type
        if raw is None:
            raw = [(torch.from_numpy(np.load(e)), torch.from_numpy(np.load(n))) for e, n in self._raw_paths(root)]
        idx = SPLIT_ORDER[task]
        self._host: List[Tuple[torch.Tensor, torch.Tensor]] = []
        for edge_features, node_features in raw:
            edge_features, node_features = torch.as_tensor(edge_features).float(), torch.as_tensor(node_features).float()
            if edge_features.dim() != 3 or edge_features.size(2) != 4 or node_features.dim() != 3 or node_features.size(2) != 6:
                raise ValueError("expected edge_features [S, E, 4] and node_features [S, n, 6]")
            split_len = [int(len(node_features) * f) for f in self.split]
            e = torch.split(edge_features, split_len, dim=0)[idx].contiguous()  # raises like the reference if the cut is ragged
            nf = torch.split(node_features, split_len, dim=0)[idx].contiguous()
            self._host.append((e, nf))
        self._counts = [int(nf.size(0)) for _, nf in self._host]
        self._first = np.concatenate([[0], np.cumsum(self._counts)]).astype(np.int64)
        self._n = np.array([int(nf.size(1)) for _, nf in self._host], dtype=np.int64)
        self._e = np.array([int(e.size(1)) for e, _ in self._host], dtype=np.int64)
        # statistics: handed in (:99-108) or taken over this split with the reference's expressions (:126-138)
        self.xymean, self.xystd = (xymean, xystd) if (xymean is not None and xystd is not None) else (None, None)
        self.edgemean, self.edgestd = (edgemean, edgestd) if (edgemean is not None and edgestd is not None) else (None, None)
        if normalize:
            if self.xymean is None:
                xy = torch.cat([nf[:, :, 2:].reshape(-1, 4) for _, nf in self._host], dim=0)
                self.xymean, self.xystd = torch.mean(xy, dim=0, keepdim=True), torch.std(xy, dim=0, keepdim=True)
            if self.edgemean is None:
                ea = torch.cat([e[:, :, 2:].reshape(-1, 2) for e, _ in self._host], dim=0)
                self.edgemean, self.edgestd = torch.mean(ea, dim=0, keepdim=True), torch.std(ea, dim=0, keepdim=True)
        self._norm = None
        if normalize:
            f = lambda t: t.detach().float().cpu().reshape(-1)  # noqa: E731
            den_xy, den_e = f(self.xystd)[:4] + 0.0000001, f(self.edgestd)[:2] + 0.0000001
            vals = torch.cat([f(self.xymean)[:4], den_xy, f(self.edgemean)[:2], den_e]).tolist()
            self._norm = (C.c_float * 12)(*vals)
        # the raw arrays go to the device once and stay there
        self._dev = [(e.to(self.device), nf.to(self.device)) for e, nf in self._host]
        self._cases = (DatasetCase * len(self._dev))(*[
            DatasetCase(nf.data_ptr(), e.data_ptr() if e.numel() else None, int(nf.size(0)), int(nf.size(1)), int(e.size(1)))
            for e, nf in self._dev])
        self._host = None  # host copies are not needed any more

    # -- reference surface ------------------------------------------------------------------------
    def _raw_paths(self, root):
        names = [self.case] if self.case != "mixed" else MIXED_CASES
        return [(os.path.join(root, "raw", f"case{c}_edge_features.npy"), os.path.join(root, "raw", f"case{c}_node_features.npy"))
                for c in names]

    def len(self) -> int:
        return int(self._first[-1])

    __len__ = len

    def get_data_dimensions(self):
        return 4, 4, 2

    def get_data_means_stds(self):
        assert self.normalize == True  # noqa: E712 -- as the reference (:116)
        return self.xymean[:1, :], self.xystd[:1, :], self.edgemean[:1, :], self.edgestd[:1, :]

    # -- batches ----------------------------------------------------------------------------------
    def batch(self, ids, ids_device: Optional[torch.Tensor] = None, seed: Optional[int] = None,
              out: Optional[GraphBatch] = None) -> GraphBatch:
        """The PyG `Batch` of samples `ids` (host sequence / tensor), assembled on the device.  `ids_device`: the same
        ids already on the GPU (int64), to skip the copy.  `out`: a batch of the same shape to overwrite (the static
        input buffers of a captured CUDA graph, `training.GraphedEpochs`)."""
        ids_host = np.asarray(torch.as_tensor(ids).cpu().numpy() if torch.is_tensor(ids) else ids, dtype=np.int64).reshape(-1)
        b = int(ids_host.size)
        if b == 0:
            raise ValueError("empty batch")
        if ids_host.min() < 0 or ids_host.max() >= self.len():
            raise IndexError(f"sample id out of range [0, {self.len()})")
        case_of = np.searchsorted(self._first, ids_host, side="right") - 1
        n_total, e_total = int(self._n[case_of].sum()), int(self._e[case_of].sum())
        dev = self.device
        with torch.cuda.device(dev):
            if ids_device is None:
                ids_device = torch.from_numpy(ids_host).to(dev, non_blocking=True)
            f32, i64 = dict(dtype=torch.float32, device=dev), dict(dtype=torch.int64, device=dev)
            if out is None:
                out = GraphBatch(x=torch.empty((n_total, 4), **f32), y=torch.empty((n_total, 4), **f32),
                                 bus_type=torch.empty((n_total,), **i64), pred_mask=torch.empty((n_total, 4), **i64),
                                 edge_index=torch.empty((2, e_total), **i64), edge_attr=torch.empty((e_total, 2), **f32),
                                 batch=torch.empty((n_total,), **i64), ptr=torch.empty((b + 1,), **i64))
            elif (tuple(out.x.shape) != (n_total, 4) or tuple(out.edge_index.shape) != (2, e_total) or out.ptr.numel() != b + 1
                  or out.x.device != dev or not all(getattr(out, f).is_contiguous() for f in GraphBatch.__dataclass_fields__)):
                raise ValueError(f"`out` does not have the shape of this batch (N={n_total}, E={e_total}, graphs={b})")
            scratch = torch.empty(int(lib().pfn_batch_assemble_scratch_bytes(b)), dtype=torch.uint8, device=dev)
            bus_seed = 0
            if self.random_bus_type:
                bus_seed = (int(seed) if seed is not None else int(torch.empty((), dtype=torch.int64).random_().item())) | 1
            check(lib().pfn_batch_assemble(
                self._cases, len(self._dev), ids_device.data_ptr(), b, n_total, e_total, self._norm, bus_seed,
                out.x.data_ptr(), out.y.data_ptr(), out.bus_type.data_ptr(), out.pred_mask.data_ptr(),
                out.edge_index.data_ptr(), out.edge_attr.data_ptr(), out.batch.data_ptr(), out.ptr.data_ptr(),
                scratch.data_ptr(), torch.cuda.current_stream().cuda_stream), "pfn_batch_assemble")
        return out

    def loader(self, batch_size: int = 1, shuffle: bool = False, generator: Optional[torch.Generator] = None,
               drop_last: bool = False, rank: int = 0, world: int = 1) -> Iterator[GraphBatch]:
        """`DataLoader(dataset, batch_size, shuffle)` (train.py:90-92): one permutation per pass, drawn on the host.
        Data parallel (`world > 1`): every rank draws the SAME permutation (seed the generators alike) and takes every
        `world`-th mini-batch of it, see `epoch_batches`."""
        for ids in epoch_batches(self.len(), batch_size, shuffle, generator, drop_last, rank, world):
            yield self.batch(ids)

    def num_batches(self, batch_size: int, drop_last: bool = False) -> int:
        return self.len() // batch_size if drop_last else -(-self.len() // batch_size)
