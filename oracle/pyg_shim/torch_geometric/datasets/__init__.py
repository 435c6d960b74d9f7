"""Import-only stub: `Planetoid` is imported (never used) at datasets/PowerFlowData.py:14."""


class Planetoid:
    def __init__(self, *a, **k):
        raise NotImplementedError("torch_geometric.datasets.Planetoid is outside the hot path; shim stub")
