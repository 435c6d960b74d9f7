"""PyG `nn` subset (test infrastructure): `MessagePassing`, `TAGConv` and import-only stubs.

Semantics restated from the published PyG 2.3-2.5 behaviour:

* `MessagePassing.propagate(edge_index, **kw)`: with `flow='source_to_target'`, `j = edge_index[0]`
  is the source and `i = edge_index[1]` the target; `<name>_i` / `<name>_j` arguments of `message`
  are `kw[name].index_select(0, i / j)`; other `message` arguments are taken from `kw` by name and
  everything else is dropped (that is why the `norm=` passed at networks/MPN.py:53 is dead);
  `aggr='add'` is `zeros(N, F).scatter_add_(0, i, msg)`; `update` defaults to identity and, like
  `message`, receives the `kw` entries its signature names (utils/custom_loss_functions.py:229).
* `TAGConv(in, out, K, bias=True, normalize=True)`: `out = sum_k lins[k](A_hat^k x) + bias`,
  `A_hat = D^-1/2 A D^-1/2` from `gcn_norm(add_self_loops=False)` with the degree taken over
  `edge_index[1]`.
"""
import inspect
import math

import torch
from torch import nn as _nn

from ..utils import degree


class MessagePassing(_nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2, **kwargs):
        super().__init__()
        if aggr not in ("add", "sum", "mean"):
            raise NotImplementedError(f"shim supports aggr add/sum/mean, got {aggr}")
        if flow not in ("source_to_target", "target_to_source"):
            raise ValueError(flow)
        self.aggr = aggr
        self.flow = flow
        self.node_dim = node_dim

    # -- hooks -----------------------------------------------------------------------------
    def message(self, x_j):
        return x_j

    def update(self, inputs):
        return inputs

    # -- driver ----------------------------------------------------------------------------
    def propagate(self, edge_index, size=None, **kwargs):
        i, j = (1, 0) if self.flow == "source_to_target" else (0, 1)
        num_nodes = None
        if size is not None:
            num_nodes = size[i] if isinstance(size, (tuple, list)) else size
        msg_args = {}
        for name in inspect.signature(self.message).parameters:
            if name.endswith("_i") or name.endswith("_j"):
                base = kwargs[name[:-2]]
                sel = edge_index[i] if name.endswith("_i") else edge_index[j]
                if num_nodes is None:
                    num_nodes = base.size(0)
                msg_args[name] = base.index_select(0, sel)
            elif name == "edge_index":
                msg_args[name] = edge_index
            elif name == "index":
                msg_args[name] = edge_index[i]
            elif name in kwargs:
                msg_args[name] = kwargs[name]
        msg = self.message(**msg_args)
        index = edge_index[i]
        if num_nodes is None:
            num_nodes = int(index.max()) + 1 if index.numel() else 0
        out = torch.zeros((num_nodes,) + tuple(msg.shape[1:]), dtype=msg.dtype, device=msg.device)
        expand = index.view((-1,) + (1,) * (msg.dim() - 1)).expand_as(msg)
        out = out.scatter_add_(0, expand, msg)
        if self.aggr == "mean":
            cnt = degree(index, num_nodes, dtype=msg.dtype).clamp_(min=1)
            out = out / cnt.view((-1,) + (1,) * (msg.dim() - 1))
        upd_args = {}
        for k, name in enumerate(inspect.signature(self.update).parameters):
            if k == 0:
                continue
            if name in kwargs:
                upd_args[name] = kwargs[name]
        return self.update(out, **upd_args)


class Linear(_nn.Module):
    """PyG `nn.dense.linear.Linear`: weight `[out, in]`, default init uniform(+-1/sqrt(in))."""

    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = _nn.Parameter(torch.empty(out_channels, in_channels))
        if bias:
            self.bias = _nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        bound = 1.0 / math.sqrt(self.in_channels) if self.in_channels > 0 else 0.0
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            if self.bias is not None:
                self.bias.zero_()

    def forward(self, x):
        return torch.nn.functional.linear(x, self.weight, self.bias)


def gcn_norm(edge_index, num_nodes, dtype):
    """`gcn_norm(edge_index, None, N, improved=False, add_self_loops=False)` of PyG."""
    row, col = edge_index[0], edge_index[1]
    w = torch.ones((edge_index.size(1),), dtype=dtype, device=edge_index.device)
    deg = torch.zeros((num_nodes,), dtype=dtype, device=edge_index.device).scatter_add_(0, col, w)
    dis = deg.pow(-0.5)
    dis.masked_fill_(dis == float("inf"), 0)
    return dis[row] * w * dis[col]


class TAGConv(MessagePassing):
    def __init__(self, in_channels, out_channels, K=3, bias=True, normalize=True, **kwargs):
        super().__init__(aggr="add", **kwargs)
        self.in_channels, self.out_channels, self.K, self.normalize = in_channels, out_channels, K, normalize
        self.lins = _nn.ModuleList([Linear(in_channels, out_channels, bias=False) for _ in range(K + 1)])
        if bias:
            self.bias = _nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        for lin in self.lins:
            lin.reset_parameters()
        if self.bias is not None:
            with torch.no_grad():
                self.bias.zero_()

    def forward(self, x, edge_index, edge_weight=None):
        if self.normalize:
            assert edge_weight is None, "shim: weighted TAGConv is outside the hot path"
            edge_weight = gcn_norm(edge_index, x.size(0), x.dtype)
        out = self.lins[0](x)
        for lin in self.lins[1:]:
            x = self.propagate(edge_index, x=x, edge_weight=edge_weight)
            out = out + lin(x)
        if self.bias is not None:
            out = out + self.bias
        return out

    def message(self, x_j, edge_weight):
        return x_j if edge_weight is None else edge_weight.view(-1, 1) * x_j


class _ImportOnly(_nn.Module):
    """`GCNConv` / `ChebConv` are imported by networks/MPN.py:3 and networks/GCN.py but are used only by
    ablation models outside the hot path."""

    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError(f"{type(self).__name__} is outside the hot path; shim stub")


class GCNConv(_ImportOnly):
    pass


class ChebConv(_ImportOnly):
    pass
