"""PyG `utils` subset (test infrastructure).  Only `degree` is on the hot path
(reference call sites: networks/MPN.py:44,128)."""
import torch


def degree(index, num_nodes=None, dtype=None):
    """`zeros(N).scatter_add_(0, index, ones(E))` -- PyG `torch_geometric.utils.degree`."""
    if num_nodes is None:
        num_nodes = int(index.max()) + 1 if index.numel() > 0 else 0
    out = torch.zeros((num_nodes,), dtype=dtype, device=index.device)
    one = torch.ones((index.size(0),), dtype=out.dtype, device=out.device)
    return out.scatter_add_(0, index, one)


def _unsupported(name):
    def fn(*a, **k):
        raise NotImplementedError(f"torch_geometric.utils.{name} is outside the hot path; shim stub")
    fn.__name__ = name
    return fn


# imported (never called on the path) by datasets/PowerFlowData.py:12 and utils/explanation.py:13,18
from_scipy_sparse_matrix = _unsupported("from_scipy_sparse_matrix")
dense_to_sparse = _unsupported("dense_to_sparse")
k_hop_subgraph = _unsupported("k_hop_subgraph")
to_networkx = _unsupported("to_networkx")
