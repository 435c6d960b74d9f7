"""TEST INFRASTRUCTURE ONLY -- minimal stand-in for the un-vendored `torch_geometric` dependency.

The reference (StavrosOrf/PoweFlowNet) imports PyTorch-Geometric at module top
(`networks/MPN.py:3-4`, `datasets/PowerFlowData.py:10-15`, `utils/training.py:6`), but PyG is not
installed in this image and cannot be (no network).  This package restates, in plain torch, only
the PyG behaviour the hot path relies on (published PyG 2.3-2.5 semantics, summarised in
SURVEY.md section 8c), so that the reference's OWN model code can be imported and executed to
produce the golden vectors under `tests/golden/`.

Nothing in the product package (`poweflownet_b200/`) imports this.  Parity status: the
reference's `networks/MPN.py` runs unmodified on top of it, but the PyG internals themselves
(`MessagePassing.propagate`, `TAGConv`, `gcn_norm`, `degree`, `Batch` collation) are a
restatement => "parity unpinned" for those pieces (no PyG install, no reference fixtures).
"""
__version__ = "0.0.shim"

from . import utils, nn, data, loader, datasets  # noqa: F401,E402
