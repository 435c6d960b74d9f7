"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (plain PyTorch, op for op) of the reference hot path
`MaskEmbdMultiMPN` forward (+ autograd backward) of StavrosOrf/PoweFlowNet, used only as the
checker by `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs.  Nothing under `poweflownet_b200/` may import this file.

PARITY STATUS: **partially pinned**.  The reference publishes no golden vectors, tests or
checkpoints, and its graph arithmetic lives in `torch_geometric`, an
un-vendored, un-pinned dependency that is absent from this image.  What IS
pinned: `tests/golden/make_golden.py` imports the reference's real `networks/MPN.py` on top of
the minimal PyG stand-in in `oracle/pyg_shim/` and stores its outputs and gradients; this file is
checked against those fixtures (`tests/test_oracle.py`).  What is NOT pinned: the PyG internals
themselves (`MessagePassing.propagate`, `TAGConv`, `gcn_norm`, `degree`), which both this file and
the shim restate from PyG's published 2.3-2.5 semantics => "parity unpinned" for those pieces.

Every function cites the reference lines it follows (paths relative to /root/reference).
The same code runs in fp64 (`model.double()`) as the error-budget twin.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
from torch import nn
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# PyG primitives restated
# --------------------------------------------------------------------------------------------
def scatter_sum(msg: torch.Tensor, index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """PyG `aggr='add'`: `zeros(N, F).scatter_add_(0, index[:, None].expand_as(msg), msg)`."""
    out = torch.zeros((num_nodes, msg.size(1)), dtype=msg.dtype, device=msg.device)
    return out.scatter_add_(0, index.view(-1, 1).expand_as(msg), msg)


def in_degree(index: torch.Tensor, num_nodes: int, dtype) -> torch.Tensor:
    """PyG `utils.degree` (call site networks/MPN.py:44)."""
    return torch.zeros((num_nodes,), dtype=dtype, device=index.device).scatter_add_(
        0, index, torch.ones((index.numel(),), dtype=dtype, device=index.device))


def gcn_norm_weights(edge_index: torch.Tensor, num_nodes: int, dtype) -> torch.Tensor:
    """PyG `gcn_norm(add_self_loops=False)`: `w_e = d^-1/2[row_e] * 1 * d^-1/2[col_e]`, `d` = in-degree over
    `col = edge_index[1]`, `inf -> 0`."""
    row, col = edge_index[0], edge_index[1]
    dis = in_degree(col, num_nodes, dtype).pow(-0.5)
    dis = torch.where(torch.isinf(dis), torch.zeros_like(dis), dis)
    return dis[row] * torch.ones_like(dis[row]) * dis[col]


# --------------------------------------------------------------------------------------------
# networks/MPN.py:6-56  EdgeAggregation
# --------------------------------------------------------------------------------------------
def edge_aggregation(x, edge_index, edge_attr, w1, b1, w2, b2):
    """networks/MPN.py:23-28 (message) + :30-56 (forward).

    Default flow 'source_to_target': `x_j = x[edge_index[0]]` (source), `x_i = x[edge_index[1]]`
    (target); message = Linear(ReLU(Linear(cat[x_i, x_j, edge_attr]))) (:17-21,:28); messages are summed
    onto the target node (`aggr='add'`, :11).  The degree norm computed at :43-47 is handed to
    `propagate` but `message` has no `norm` parameter, so it never touches the result; it is omitted.
    """
    src, tgt = edge_index[0], edge_index[1]
    feats = torch.cat([x.index_select(0, tgt), x.index_select(0, src), edge_attr], dim=-1)
    hidden = torch.relu(F.linear(feats, w1, b1))
    msg = F.linear(hidden, w2, b2)
    return scatter_sum(msg, tgt, x.size(0))


class EdgeAggregation(nn.Module):
    """State-dict layout of networks/MPN.py:10-21: `edge_aggr.0.{weight,bias}`, `edge_aggr.2.{weight,bias}`."""

    def __init__(self, nfeature_dim, efeature_dim, hidden_dim, output_dim):
        super().__init__()
        self.nfeature_dim, self.efeature_dim, self.output_dim = nfeature_dim, efeature_dim, output_dim
        self.edge_aggr = nn.Sequential(
            nn.Linear(nfeature_dim * 2 + efeature_dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, output_dim))

    def forward(self, x, edge_index, edge_attr):
        l0, l2 = self.edge_aggr[0], self.edge_aggr[2]
        return edge_aggregation(x, edge_index, edge_attr, l0.weight, l0.bias, l2.weight, l2.bias)


# --------------------------------------------------------------------------------------------
# torch_geometric.nn.TAGConv (call sites networks/MPN.py:477,480,484 ctor; :545 forward)
# --------------------------------------------------------------------------------------------
def tag_conv(x, edge_index, lin_weights: Sequence[torch.Tensor], bias: Optional[torch.Tensor]):
    """`out = sum_{k=0..K} lins[k](A_hat^k x) + bias` with the hop `x <- scatter_sum(w_e * x[row_e], col_e)`."""
    n = x.size(0)
    w = gcn_norm_weights(edge_index, n, x.dtype)
    row, col = edge_index[0], edge_index[1]
    out = F.linear(x, lin_weights[0])
    for wk in lin_weights[1:]:
        x = scatter_sum(w.view(-1, 1) * x.index_select(0, row), col, n)
        out = out + F.linear(x, wk)
    if bias is not None:
        out = out + bias
    return out


class _BiasFreeLinear(nn.Module):
    def __init__(self, fin, fout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(fout, fin))
        bound = 1.0 / math.sqrt(fin)
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)


class TAGConv(nn.Module):
    """State-dict layout of PyG TAGConv: `lins.{k}.weight` [out, in] (k = 0..K), `bias` [out] (zero init)."""

    def __init__(self, in_channels, out_channels, K=3):
        super().__init__()
        self.in_channels, self.out_channels, self.K = in_channels, out_channels, K
        self.lins = nn.ModuleList([_BiasFreeLinear(in_channels, out_channels) for _ in range(K + 1)])
        self.bias = nn.Parameter(torch.zeros(out_channels))

    def forward(self, x, edge_index):
        return tag_conv(x, edge_index, [l.weight for l in self.lins], self.bias)


# --------------------------------------------------------------------------------------------
# networks/MPN.py:498-523  is_directed / undirect_graph  (integer work, bit-exact)
# --------------------------------------------------------------------------------------------
def is_directed(edge_index: torch.Tensor) -> bool:
    """networks/MPN.py:498-504: no edges -> False; otherwise look ONLY at the first edge (a -> b) and
    report True iff no edge (b -> a) exists anywhere in the list."""
    if edge_index.shape[1] == 0:
        return False
    a, b = edge_index[0, 0], edge_index[1, 0]
    targets_of_b = edge_index[1, edge_index[0, :] == b]
    return not bool((targets_of_b == a).any())


def undirect_graph(edge_index: torch.Tensor, edge_attr: torch.Tensor):
    """networks/MPN.py:506-523: if directed, append every edge reversed (all originals first, then all
    reversed) and duplicate `edge_attr`; otherwise pass both through untouched."""
    if not is_directed(edge_index):
        return edge_index, edge_attr
    flipped = torch.stack([edge_index[1, :], edge_index[0, :]], dim=0)
    return torch.cat([edge_index, flipped], dim=1), torch.cat([edge_attr, edge_attr], dim=0)


# --------------------------------------------------------------------------------------------
# networks/MPN.py:456-559  MaskEmbdMultiMPN
# --------------------------------------------------------------------------------------------
class MaskEmbdMultiMPN(nn.Module):
    """Constructor mirrors networks/MPN.py:462-496 (same attribute names => same state_dict keys)."""

    def __init__(self, nfeature_dim, efeature_dim, output_dim, hidden_dim, n_gnn_layers, K, dropout_rate):
        super().__init__()
        self.nfeature_dim, self.efeature_dim, self.output_dim = nfeature_dim, efeature_dim, output_dim
        self.hidden_dim, self.n_gnn_layers, self.K, self.dropout_rate = hidden_dim, n_gnn_layers, K, dropout_rate
        layers: List[nn.Module] = [EdgeAggregation(nfeature_dim, efeature_dim, hidden_dim, hidden_dim)]
        # :475-480 -- with a single GNN layer the TAGConv emits `output_dim`, otherwise `hidden_dim`
        layers.append(TAGConv(hidden_dim, output_dim if n_gnn_layers == 1 else hidden_dim, K=K))
        for _ in range(n_gnn_layers - 2):  # :482-484
            layers.append(EdgeAggregation(hidden_dim, efeature_dim, hidden_dim, hidden_dim))
            layers.append(TAGConv(hidden_dim, hidden_dim, K=K))
        layers.append(EdgeAggregation(hidden_dim, efeature_dim, hidden_dim, output_dim))  # :489
        self.layers = nn.ModuleList(layers)
        self.mask_embd = nn.Sequential(  # :491-495
            nn.Linear(nfeature_dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, nfeature_dim))
        self.dropout = nn.Dropout(dropout_rate, inplace=False)  # :496

    is_directed = staticmethod(is_directed)
    undirect_graph = staticmethod(undirect_graph)

    def forward(self, data, dropout_masks: Optional[Sequence[torch.Tensor]] = None):
        """networks/MPN.py:525-559.  `dropout_masks` (test hook, not in the reference): one 0/1 tensor per
        dropout application (len(layers)-1 of them); when given, `x * mask / (1-p)` replaces `nn.Dropout`
        so that train-mode numerics can be compared across implementations with different RNGs."""
        assert data.x.shape[-1] == 4  # :528
        x = data.x
        mask = data.pred_mask.to(x.dtype)  # :533 (`.float()`; dtype-generic so the fp64 twin works)
        edge_index, edge_attr = data.edge_index, data.edge_attr
        x = self.mask_embd(mask) + x  # :537
        edge_index, edge_attr = undirect_graph(edge_index, edge_attr)  # :539
        for i in range(len(self.layers) - 1):  # :541-547
            layer = self.layers[i]
            if isinstance(layer, EdgeAggregation):
                x = layer(x, edge_index, edge_attr)
            else:
                x = layer(x, edge_index)
            if dropout_masks is not None:
                x = x * dropout_masks[i].to(x.dtype) / (1.0 - self.dropout_rate)
            else:
                x = self.dropout(x)
            x = torch.relu(x)
        last = self.layers[-1]  # :554-557
        if isinstance(last, EdgeAggregation):
            x = last(x, edge_index, edge_attr)
        else:
            x = last(x, edge_index)
        return x


# --------------------------------------------------------------------------------------------
# utils/custom_loss_functions.py:10-46  Masked_L2_loss  (boundary of the path)
# --------------------------------------------------------------------------------------------
def masked_l2_loss(output, target, mask, regularize=True, regcoeff=1.0):
    """MSE over the entries where mask==1, plus (if `regularize`) `regcoeff` x MSE over mask==0."""
    sel = mask.to(torch.bool)
    loss = F.mse_loss(torch.masked_select(output, sel), torch.masked_select(target, sel))
    if regularize:
        inv = (1 - mask).to(torch.bool)
        loss = loss + regcoeff * F.mse_loss(torch.masked_select(output, inv), torch.masked_select(target, inv))
    return loss


# --------------------------------------------------------------------------------------------
# utils/custom_loss_functions.py:99-306  PowerImbalance / MixedMSEPoweImbalance  
# --------------------------------------------------------------------------------------------
def power_imbalance_message(x_i, x_j, edge_attr):
    """utils/custom_loss_functions.py:159-227 (`message`), the "another mine" formula that is live at :216-217.
    `x_*`: de-normalised (Vm, Va[deg], P, Q) of the aggregating / the neighbouring bus, `edge_attr`: de-normalised
    (r, x) of the branch.  Returns `[E, 2]` = (Pji, Qji)."""
    r_x = edge_attr[:, 0:2]
    r, x = r_x[:, 0:1], r_x[:, 1:2]
    g_ij = r / (r ** 2 + x ** 2)
    b_ij = -x / (r ** 2 + x ** 2)
    vm_i = x_i[:, 0:1]
    va_i = 1 / 180. * math.pi * x_i[:, 1:2]
    vm_j = x_j[:, 0:1]
    va_j = 1 / 180. * math.pi * x_j[:, 1:2]
    e_i = vm_i * torch.cos(va_i)
    f_i = vm_i * torch.sin(va_i)
    e_j = vm_j * torch.cos(va_j)
    f_j = vm_j * torch.sin(va_j)
    Pji = g_ij * (e_i * e_j - e_i ** 2 + f_i * f_j - f_i ** 2) + b_ij * (f_i * e_j - e_i * f_j)
    Qji = g_ij * (f_i * e_j - e_i * f_j) + b_ij * (-e_i * e_j + e_i ** 2 - f_i * f_j + f_i ** 2)
    return torch.cat([Pji, Qji], dim=-1)


def power_imbalance(x, edge_index, edge_attr, xymean, xystd, edgemean, edgestd):
    """utils/custom_loss_functions.py:254-286 (`forward`) with :229-252 (`update`) and :124-129 (`de_normalize`).

    `MessagePassing(aggr='add', flow='target_to_source')` (:118): `i = edge_index[0]` is the aggregating bus and
    `j = edge_index[1]` its neighbour.  The branch list is doubled when `is_directed` (:131-133: the same
    first-edge test as the model's, without the empty-list guard).  `dPQ_i = -sum_e msg_e + (P_i, Q_i)`;
    loss = mean over buses of `dP^2 + dQ^2`.  `xymean/xystd`: `[1, 4]` (rows beyond the first are dropped, :119-122),
    `edgemean/edgestd`: `[1, 2]`."""
    xymean, xystd = xymean[0:1].to(x.dtype), xystd[0:1].to(x.dtype)
    edgemean, edgestd = edgemean.to(x.dtype), edgestd.to(x.dtype)
    if is_directed(edge_index):
        edge_index, edge_attr = undirect_graph(edge_index, edge_attr)
    x = x * xystd + xymean
    edge_attr = edge_attr * edgestd + edgemean
    i, j = edge_index[0], edge_index[1]
    msg = power_imbalance_message(x.index_select(0, i), x.index_select(0, j), edge_attr)
    agg = scatter_sum(msg, i, x.size(0))
    dPi = -agg[:, 0:1] + x[:, 2:3]
    dQi = -agg[:, 1:2] + x[:, 3:4]
    dPQ = torch.cat([dPi, dQi], dim=-1)
    return dPQ.square().sum(dim=-1).mean()


def mixed_mse_power_imbalance(x, edge_index, edge_attr, y, stats, alpha=0.5):
    """utils/custom_loss_functions.py:289-306: `alpha * MSE(x, y) + (1 - alpha) * 0.020 * power_imbalance`."""
    return alpha * F.mse_loss(x, y) + (1 - alpha) * 0.020 * power_imbalance(x, edge_index, edge_attr, *stats)


def adamw_step(params, grads, exp_avgs, exp_avg_sqs, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-2):
    """One `torch.optim.AdamW` update (train.py:123; torch/optim/adamw.py `_single_tensor_adamw`, no amsgrad / maximize),
    in place on lists of tensors: decoupled decay, moment updates, bias corrections taken in Python doubles."""
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    for p, g, m, v in zip(params, grads, exp_avgs, exp_avg_sqs):
        p.mul_(1 - lr * weight_decay)
        m.lerp_(g, 1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = (v.sqrt() / bc2_sqrt).add_(eps)
        p.addcdiv_(m, denom, value=-step_size)


# --------------------------------------------------------------------------------------------
# datasets/PowerFlowData.py:119-140,171-217 + PyG collation  
# --------------------------------------------------------------------------------------------
BUS_TYPE_MASK = ((0, 0, 1, 1), (0, 1, 0, 1), (1, 1, 0, 0))  # datasets/PowerFlowData.py:71-74


def process_split(raw_cases, split, task):
    """datasets/PowerFlowData.py:171-217 (`process`) for one task: `raw_cases` = [(edge_features [S, E, 4],
    node_features [S, n, 6]), ...] in `raw_file_names` order; every case is cut with `torch.split` by
    `[int(S * f) for f in split]` and the per-case sample lists of the task are concatenated (:214).  Returns a list
    of dicts with the `Data` fields."""
    idx = {"train": 0, "val": 1, "test": 2}[task]
    table = torch.tensor(BUS_TYPE_MASK)
    samples = []
    for edge_features, node_features in raw_cases:
        edge_features, node_features = edge_features.float(), node_features.float()
        split_len = [int(len(node_features) * f) for f in split]
        e = torch.split(edge_features, split_len, dim=0)[idx]
        nf = torch.split(node_features, split_len, dim=0)[idx]
        y = nf[:, :, 2:]
        bus_type = nf[:, :, 1].type(torch.long)
        mask = table[bus_type]
        x = y.clone() * (1. - mask)
        for i in range(len(x)):
            samples.append(dict(x=x[i], y=y[i], bus_type=bus_type[i], pred_mask=mask[i],
                                edge_index=e[i, :, 0:2].T.to(torch.long), edge_attr=e[i, :, 2:]))
    return samples


def dataset_stats(samples):
    """datasets/PowerFlowData.py:126-138: mean / std (unbiased) over every bus / branch of the split."""
    y = torch.cat([s["y"] for s in samples], dim=0)
    ea = torch.cat([s["edge_attr"] for s in samples], dim=0)
    return (torch.mean(y, dim=0, keepdim=True), torch.std(y, dim=0, keepdim=True),
            torch.mean(ea, dim=0, keepdim=True), torch.std(ea, dim=0, keepdim=True))


def collate_batch(samples, ids, stats=None):
    """`_normalize_dataset` (:132-139, when `stats` is given) + PyG `Batch.from_data_list` of `samples[ids]`: tensors
    concatenated on dim 0, `edge_index` on dim 1 with cumulative node offsets, plus `batch` and `ptr`."""
    out = {k: [] for k in ("x", "y", "bus_type", "pred_mask", "edge_index", "edge_attr", "batch")}
    ptr = [0]
    for b, i in enumerate(int(v) for v in ids):
        s = samples[i]
        x, y, ea = s["x"], s["y"], s["edge_attr"]
        if stats is not None:
            xymean, xystd, edgemean, edgestd = stats
            x = (x - xymean) / (xystd + 0.0000001)
            y = (y - xymean) / (xystd + 0.0000001)
            ea = (ea - edgemean) / (edgestd + 0.0000001)
        out["x"].append(x)
        out["y"].append(y)
        out["edge_attr"].append(ea)
        out["bus_type"].append(s["bus_type"])
        out["pred_mask"].append(s["pred_mask"])
        out["edge_index"].append(s["edge_index"] + ptr[-1])
        out["batch"].append(torch.full((x.size(0),), b, dtype=torch.long))
        ptr.append(ptr[-1] + x.size(0))
    res = {k: torch.cat(v, dim=1 if k == "edge_index" else 0) for k, v in out.items()}
    res["ptr"] = torch.tensor(ptr, dtype=torch.long)
    return res


# --------------------------------------------------------------------------------------------
# utils/training.py:55-77  one optimisation step's model work (forward + loss + backward)
# --------------------------------------------------------------------------------------------
def forward_loss_backward(model: MaskEmbdMultiMPN, data, loss: str = "mse", dropout_masks=None):
    """`out = model(data)` (:58), loss dispatch (:61-72; 'mse' is `torch.nn.MSELoss`, train.py:103;
    'masked_l2' is the parser default), `loss.backward()` (:74).  Returns (loss, out)."""
    out = model(data, dropout_masks=dropout_masks) if dropout_masks is not None else model(data)
    if loss == "mse":
        val = F.mse_loss(out, data.y)
    elif loss == "masked_l2":
        val = masked_l2_loss(out, data.y, data.pred_mask)
    else:
        raise ValueError(loss)
    val.backward()
    return val.detach(), out.detach()
