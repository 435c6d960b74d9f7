"""Mini-batch container and synthetic `PowerFlowData`-shaped batches.

`GraphBatch` carries exactly the tensors the reference model reads from a PyG `Batch`
(datasets/PowerFlowData.py:191-205 + PyG collation): `x [N,4] f32`, `y [N,4] f32`,
`bus_type [N] i64`, `pred_mask [N,4] i64`, `edge_index [2,E_raw] i64` (one entry per branch, node
offsets accumulated across graphs), `edge_attr [E_raw,2] f32`, `batch [N] i64`, `ptr [B+1] i64`.
Any object exposing those attributes (a real PyG `Batch` included) is accepted by the model.

No dataset can be downloaded here, so `synthetic_batch` builds batches of the reference's
shapes: one fixed connected topology per case, stored one direction per
branch so that `is_directed` is True and the undirect path (networks/MPN.py:506-523) is exercised.
"""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Dict, List, Optional, Sequence, Tuple

import torch

# (buses, branches) of the grids the reference trains on (runs.sh:2, dataset_generator.py:154-164)
CASES: Dict[str, Tuple[int, int]] = {
    "14": (14, 20),
    "118v2": (118, 186),
    "6470rte": (6470, 9005),
}

# datasets/PowerFlowData.py:71-74 -- which of (Vm, Va, P, Q) must be predicted for slack / PV / PQ buses
BUS_TYPE_MASK = ((0, 0, 1, 1), (0, 1, 0, 1), (1, 1, 0, 0))


@dataclass
class GraphBatch:
    x: torch.Tensor
    y: torch.Tensor
    bus_type: torch.Tensor
    pred_mask: torch.Tensor
    edge_index: torch.Tensor
    edge_attr: torch.Tensor
    batch: torch.Tensor
    ptr: torch.Tensor

    @property
    def num_graphs(self) -> int:
        return int(self.ptr.numel()) - 1

    @property
    def num_nodes(self) -> int:
        return int(self.x.size(0))

    def __len__(self) -> int:
        # PyG `BaseData.__len__` counts stored attributes; utils/training.py:76-77 relies on it.
        return len(fields(self))

    def to(self, device, non_blocking: bool = False) -> "GraphBatch":
        return GraphBatch(*[getattr(self, f.name).to(device, non_blocking=non_blocking) for f in fields(self)])

    def pin_memory(self) -> "GraphBatch":
        return GraphBatch(*[getattr(self, f.name).pin_memory() for f in fields(self)])

    def nbytes(self) -> int:
        return sum(getattr(self, f.name).numel() * getattr(self, f.name).element_size() for f in fields(self))


def synthetic_topology(n: int, e_raw: int, seed: Optional[int] = None) -> torch.Tensor:
    """`[2, e_raw]` int64 branch list of one connected grid: a random spanning tree over a node
    permutation plus `e_raw-(n-1)` extra random non-self branches, one direction each (parallel
    branches allowed, as in real cases with double circuits)."""
    if e_raw < n - 1:
        raise ValueError("need at least n-1 branches for a connected grid")
    g = torch.Generator().manual_seed(1000 + n if seed is None else seed)
    perm = torch.randperm(n, generator=g)
    parent_pos = (torch.rand(n - 1, generator=g) * torch.arange(1, n)).floor().long()
    tree = torch.stack([perm[parent_pos], perm[1:]], dim=0)
    extra = e_raw - (n - 1)
    a = torch.randint(0, n, (extra,), generator=g)
    b = (a + torch.randint(1, n, (extra,), generator=g)) % n if n > 1 else a
    edges = torch.cat([tree, torch.stack([a, b], dim=0)], dim=1)
    # keep edge 0's reverse out of the list so that the first-edge test reports "directed"
    a0, b0 = int(edges[0, 0]), int(edges[1, 0])
    rev = (edges[0] == b0) & (edges[1] == a0)
    edges[:, rev] = torch.tensor([[a0], [b0]])
    return edges.contiguous()


def synthetic_batch(case: str = "118v2", batch_size: int = 128, seed: int = 1234,
                    cases: Optional[Sequence[str]] = None) -> GraphBatch:
    """A `batch_size`-graph mini-batch of `case` (or of the per-graph list `cases` -- names from
    `CASES` or `(buses, branches)` tuples -- for variable-N batches) in PyG `Batch` layout.  `y ~ N(0,1)` (z-scored targets, PowerFlowData.py:133),
    `x = y * (1 - mask)` (:194), `edge_attr ~ N(0,1)` (:139); node 0 of each graph is the slack bus,
    ~45 % of the others are PV, the rest PQ."""
    names = list(cases) if cases is not None else [case] * batch_size
    g = torch.Generator().manual_seed(seed)
    mask_table = torch.tensor(BUS_TYPE_MASK, dtype=torch.long)
    dims = {c: (CASES[c] if isinstance(c, str) else (int(c[0]), int(c[1]))) for c in set(names)}
    topo = {c: synthetic_topology(*dims[c]) for c in dims}
    xs, ys, bts, pms, eis, eas, bs, ptr = [], [], [], [], [], [], [], [0]
    off = 0
    for gi, c in enumerate(names):
        n, e_raw = dims[c]
        bt = torch.where(torch.rand(n, generator=g) < 0.45, 1, 2).long()
        bt[0] = 0
        pm = mask_table[bt]
        y = torch.randn(n, 4, generator=g)
        xs.append(y * (1.0 - pm.float()))
        ys.append(y)
        bts.append(bt)
        pms.append(pm)
        eis.append(topo[c] + off)
        eas.append(torch.randn(e_raw, 2, generator=g))
        bs.append(torch.full((n,), gi, dtype=torch.long))
        off += n
        ptr.append(off)
    return GraphBatch(
        x=torch.cat(xs).contiguous(), y=torch.cat(ys).contiguous(), bus_type=torch.cat(bts),
        pred_mask=torch.cat(pms).contiguous(), edge_index=torch.cat(eis, dim=1).contiguous(),
        edge_attr=torch.cat(eas).contiguous(), batch=torch.cat(bs), ptr=torch.tensor(ptr, dtype=torch.long))


def synthetic_raw_case(case: str = "118v2", samples: int = 64, seed: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """(edge_features [S, E, 4] = (from, to, r, x), node_features [S, n, 6] = (index, type, Vm, Va, P, Q)), fp32, in the
    layout of the reference's raw files (datasets/PowerFlowData.py:58-61,178-204): one fixed topology, per-sample
    branch parameters and bus states.  Input for `datasets.PowerFlowData(raw=[...])` when no dataset can be downloaded."""
    n, e_raw = CASES[case]
    g = torch.Generator().manual_seed(seed)
    edges = torch.zeros((samples, e_raw, 4))
    edges[:, :, :2] = synthetic_topology(n, e_raw).T.float()
    edges[:, :, 2:] = (0.05 + 0.02 * torch.randn((samples, e_raw, 2), generator=g)).abs() + 1e-3
    nodes = torch.zeros((samples, n, 6))
    nodes[:, :, 0] = torch.arange(n).float()
    bt = torch.where(torch.rand(n, generator=g) < 0.45, 1, 2)
    bt[0] = 0
    nodes[:, :, 1] = bt.float()
    nodes[:, :, 2] = 1.0 + 0.02 * torch.randn((samples, n), generator=g)
    nodes[:, :, 3] = 5.0 * torch.randn((samples, n), generator=g)
    nodes[:, :, 4:] = 0.5 * torch.randn((samples, n, 2), generator=g)
    return edges, nodes


def shard_batch(batch: GraphBatch, rank: int, world: int) -> GraphBatch:
    """Contiguous block of whole graphs for `rank` (split at `ptr` boundaries, balanced by branch
    count so that a 6470-bus graph is not weighed like a 14-bus one)."""
    b = batch.num_graphs
    ptr = batch.ptr
    if world == 1:
        return batch
    src_graph = torch.bucketize(batch.edge_index[0], ptr[1:], right=True)
    e_per_graph = torch.bincount(src_graph, minlength=b).double()
    csum = torch.cat([torch.zeros(1, dtype=torch.double), e_per_graph.cumsum(0)])
    total = float(csum[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        lo, hi = cuts[-1] + 1, b - (world - r)  # leave at least one graph for every rank when b >= world
        if hi < lo:
            g_cut = min(max(cuts[-1], 0), b)
        else:
            g_cut = lo + int(torch.argmin((csum[lo:hi + 1] - target).abs()))
        cuts.append(g_cut)
    cuts.append(b)
    g0, g1 = cuts[rank], cuts[rank + 1]
    n0, n1 = int(ptr[g0]), int(ptr[g1])
    emask = (batch.edge_index[0] >= n0) & (batch.edge_index[0] < n1)
    return GraphBatch(
        x=batch.x[n0:n1].contiguous(), y=batch.y[n0:n1].contiguous(), bus_type=batch.bus_type[n0:n1].contiguous(),
        pred_mask=batch.pred_mask[n0:n1].contiguous(),
        edge_index=(batch.edge_index[:, emask] - n0).contiguous(), edge_attr=batch.edge_attr[emask].contiguous(),
        batch=(batch.batch[n0:n1] - g0).contiguous(), ptr=(ptr[g0:g1 + 1] - n0).contiguous())
