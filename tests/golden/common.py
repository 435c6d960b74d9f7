"""Shared definitions of the golden cases: model configs, seeded inputs and seeded weights.

Used by `make_golden.py` (which runs the reference's real `networks/MPN.py`) and by the tests
(which re-create the same inputs/weights and compare against the stored reference outputs).
Weights are re-derived from a seed instead of stored so the fixtures stay small.
"""
from __future__ import annotations

import math
import os
import sys
from collections import OrderedDict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from poweflownet_b200.data import GraphBatch, synthetic_batch  # noqa: E402

GOLDEN_DIR = os.path.dirname(os.path.abspath(__file__))

# name -> (model kwargs, batch spec).  nfeature_dim/efeature_dim/output_dim are the dataset's
# real dims (4, 2, 4) as train.py:106-117 passes them; hidden/L/K/dropout from configs/*.json.
CASES = OrderedDict(
    tiny=dict(model=dict(hidden_dim=8, n_gnn_layers=2, K=2, dropout_rate=0.2),
              batch=dict(cases=[(6, 7), (5, 4), (6, 7)], seed=11)),
    case14_small=dict(model=dict(hidden_dim=64, n_gnn_layers=2, K=3, dropout_rate=0.2),      # configs/small.json
                      batch=dict(case="14", batch_size=16, seed=1234)),
    case118_h33=dict(model=dict(hidden_dim=33, n_gnn_layers=4, K=3, dropout_rate=0.2),
                     batch=dict(case="118v2", batch_size=2, seed=21)),
    case118_standard=dict(model=dict(hidden_dim=129, n_gnn_layers=4, K=3, dropout_rate=0.2),  # configs/standard.json
                          batch=dict(case="118v2", batch_size=2, seed=1234), summary_only=True),
    mixed=dict(model=dict(hidden_dim=16, n_gnn_layers=3, K=2, dropout_rate=0.1),
               batch=dict(cases=["14", "118v2", "14", (9, 12)], seed=5)),
    already_undirected=dict(model=dict(hidden_dim=12, n_gnn_layers=2, K=3, dropout_rate=0.2),
                            batch=dict(special="already_undirected")),
    first_edge_reversed_only=dict(model=dict(hidden_dim=12, n_gnn_layers=2, K=1, dropout_rate=0.0),
                                  batch=dict(special="first_edge_reversed_only")),
    no_edges=dict(model=dict(hidden_dim=8, n_gnn_layers=2, K=2, dropout_rate=0.2),
                  batch=dict(special="no_edges")),
    isolated_and_parallel=dict(model=dict(hidden_dim=20, n_gnn_layers=3, K=3, dropout_rate=0.5),
                               batch=dict(special="isolated_and_parallel")),
)

MODEL_DIMS = dict(nfeature_dim=4, efeature_dim=2, output_dim=4)


def model_kwargs(name):
    kw = dict(MODEL_DIMS)
    kw.update(CASES[name]["model"])
    return kw


def _node_payload(n, seed):
    g = torch.Generator().manual_seed(seed)
    table = torch.tensor(((0, 0, 1, 1), (0, 1, 0, 1), (1, 1, 0, 0)), dtype=torch.long)
    bt = torch.randint(0, 3, (n,), generator=g)
    pm = table[bt]
    y = torch.randn(n, 4, generator=g)
    return y * (1.0 - pm.float()), y, bt, pm, g


def _special(kind) -> GraphBatch:
    if kind == "already_undirected":
        # both directions present in the input: is_directed() is False, edge list passes through
        n = 7
        half = torch.tensor([[0, 1, 2, 3, 4, 5, 1], [1, 2, 3, 4, 5, 6, 4]])
        ei = torch.cat([half, half.flip(0)], dim=1)
    elif kind == "first_edge_reversed_only":
        # only the FIRST branch has its reverse in the list; the reference looks at nothing else
        # (networks/MPN.py:498-504), so the graph is treated as undirected and NOT doubled
        n = 6
        ei = torch.tensor([[0, 1, 1, 2, 3, 4], [1, 0, 2, 3, 4, 5]])
    elif kind == "no_edges":
        n = 4
        ei = torch.zeros((2, 0), dtype=torch.long)
    elif kind == "isolated_and_parallel":
        # node 8 is isolated (degree 0 => d^-1/2 = inf -> 0), branches (2,3) are parallel, one self loop
        n = 9
        ei = torch.tensor([[0, 1, 2, 2, 2, 3, 4, 5, 6, 6], [1, 2, 3, 3, 3, 4, 5, 6, 7, 6]])
    else:
        raise KeyError(kind)
    x, y, bt, pm, g = _node_payload(n, 77)
    ea = torch.randn(ei.size(1), 2, generator=g)
    return GraphBatch(x=x, y=y, bus_type=bt, pred_mask=pm, edge_index=ei.contiguous(), edge_attr=ea,
                      batch=torch.zeros(n, dtype=torch.long), ptr=torch.tensor([0, n]))


def make_batch(name) -> GraphBatch:
    spec = CASES[name]["batch"]
    if "special" in spec:
        return _special(spec["special"])
    return synthetic_batch(**spec)


def seeded_state_dict(shapes, seed=4321):
    """Deterministic weights for a `{key: shape}` mapping: keys in sorted order, uniform(+-1/sqrt(fan_in))
    for matrices, uniform(+-0.5) for vectors (so TAGConv biases are non-zero and exercised)."""
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for key in sorted(shapes):
        shape = tuple(shapes[key])
        bound = 1.0 / math.sqrt(shape[-1]) if len(shape) == 2 else 0.5
        out[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return out


def load_seeded(model, seed=4321):
    shapes = {k: v.shape for k, v in model.state_dict().items()}
    model.load_state_dict(seeded_state_dict(shapes, seed))
    return model


def dropout_masks(name, n_nodes, seed=99):
    kw = model_kwargs(name)
    n_layers = 2 * kw["n_gnn_layers"] - 1 if kw["n_gnn_layers"] > 1 else 3
    g = torch.Generator().manual_seed(seed)
    keep = 1.0 - kw["dropout_rate"]
    return [(torch.rand(n_nodes, kw["hidden_dim"], generator=g) < keep).float() for _ in range(n_layers - 1)]


def golden_path(name):
    return os.path.join(GOLDEN_DIR, f"{name}.pt")


def rel_err(a, b):
    """(max-abs relative to max-abs, Frobenius relative) -- the two parity figures of SURVEY.md section 8d."""
    a, b = a.double(), b.double()
    den_max = max(float(b.abs().max()), 1e-30) if b.numel() else 1.0
    den_fro = max(float(b.norm()), 1e-30) if b.numel() else 1.0
    if a.numel() == 0:
        return 0.0, 0.0
    return float((a - b).abs().max()) / den_max, float((a - b).norm()) / den_fro
