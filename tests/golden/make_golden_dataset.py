"""Generate the dataset / batch-assembly fixtures by running the REFERENCE's own `datasets/PowerFlowData.py`.

    python tests/golden/make_golden_dataset.py     # needs /root/reference (this container only)

Seeded synthetic raw files in the reference's on-disk format (`raw/case<case>_{edge,node}_features.npy`,
datasets/PowerFlowData.py:58-61,174-204) are written to a scratch directory; the reference `PowerFlowData` -- imported
UNMODIFIED over `oracle/pyg_shim` (its `InMemoryDataset` / `DataLoader` restate PyG's) -- processes, splits and
normalises them, and `DataLoader(batch_size, shuffle=False)` collates mini-batches.  Stored per case: the raw arrays,
the normalisation statistics of every split and selected mini-batches (sample ids + every tensor of the `Batch`).
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import common  # noqa: E402

REFERENCE = os.environ.get("PFN_REFERENCE", "/root/reference")
SHIM = os.path.join(common.ROOT, "oracle", "pyg_shim")
SPLIT = [.5, .2, .3]  # train.py:76-79
FIELDS = ("x", "y", "bus_type", "pred_mask", "edge_index", "edge_attr", "batch", "ptr")

# name -> (reference `case` argument, [(case tag, buses, branches, samples)], batch size)
DATASET_CASES = {
    "ds_case14": ("14", [("14", 14, 20, 40)], 16),
    "ds_mixed": ("mixed", [("118v2", 118, 186, 10), ("14v2", 14, 20, 20)], 6),  # PowerFlowData.mixed_cases order (:67-70)
}


def dataset_golden_path(name):
    return os.path.join(common.GOLDEN_DIR, f"{name}.pt")


def synthetic_raw(n, e_raw, samples, seed):
    """(edge_features [S, E, 4] = (from, to, r, x), node_features [S, n, 6] = (index, type, Vm, Va, P, Q)) as float32
    values (the reference loads float64 .npy files and casts with `.float()`, :178-179)."""
    from poweflownet_b200.data import synthetic_topology
    rng = np.random.default_rng(seed)
    topo = synthetic_topology(n, e_raw, seed=1000 + n).numpy().T
    edges = np.zeros((samples, e_raw, 4), dtype=np.float32)
    edges[:, :, :2] = topo[None]
    edges[:, :, 2:] = (np.abs(rng.normal(0.05, 0.02, size=(samples, e_raw, 2))) + 1e-3).astype(np.float32)
    nodes = np.zeros((samples, n, 6), dtype=np.float32)
    nodes[:, :, 0] = np.arange(n)[None]
    types_ = np.where(rng.random(n) < 0.45, 1, 2)
    types_[0] = 0
    nodes[:, :, 1] = types_[None]
    nodes[:, :, 2] = 1.0 + rng.normal(0, 0.02, size=(samples, n))
    nodes[:, :, 3] = rng.normal(0, 5.0, size=(samples, n))
    nodes[:, :, 4:] = rng.normal(0, 30.0, size=(samples, n, 2))
    return edges, nodes


def raw_arrays(name):
    _, parts, _ = DATASET_CASES[name]
    return [(tag,) + synthetic_raw(n, e, s, seed=17 + k) for k, (tag, n, e, s) in enumerate(parts)]


def main():
    if not os.path.isdir(REFERENCE):
        raise SystemExit(f"{REFERENCE} not found: fixtures can only be regenerated where the reference is mounted")
    for p in (SHIM, REFERENCE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.setdefault("TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD", "1")  # PowerFlowData torch.load()s pickled Data objects
    from datasets.PowerFlowData import PowerFlowData  # the reference's file, byte for byte
    from torch_geometric.loader import DataLoader
    torch.set_num_threads(1)
    for name, (case, _, batch_size) in DATASET_CASES.items():
        raws = raw_arrays(name)
        out = {"raw": {tag: {"edge_features": torch.from_numpy(e), "node_features": torch.from_numpy(nf)} for tag, e, nf in raws},
               "splits": {}, "meta": {"case": case, "split": SPLIT, "batch_size": batch_size,
                                      "source": "reference datasets/PowerFlowData.py over oracle/pyg_shim"}}
        with tempfile.TemporaryDirectory() as root:
            os.makedirs(os.path.join(root, "raw"))
            for tag, e, nf in raws:
                np.save(os.path.join(root, "raw", f"case{tag}_edge_features.npy"), e.astype(np.float64))
                np.save(os.path.join(root, "raw", f"case{tag}_node_features.npy"), nf.astype(np.float64))
            for task in ("train", "val", "test"):
                ds = PowerFlowData(root=root, case=case, split=SPLIT, task=task, normalize=True)
                stats = [t.clone() for t in ds.get_data_means_stds()]
                batches = []
                for k, b in enumerate(DataLoader(ds, batch_size=batch_size, shuffle=False)):
                    ids = torch.arange(k * batch_size, min(len(ds), (k + 1) * batch_size))
                    batches.append({"ids": ids, **{f: getattr(b, f).clone() for f in FIELDS}})
                out["splits"][task] = {"len": len(ds), "stats": stats, "dims": ds.get_data_dimensions(), "batches": batches}
            # a validation set that is handed the training statistics (the xymean=... constructor arguments, :85-108)
            tr = out["splits"]["train"]["stats"]
            ds = PowerFlowData(root=root, case=case, split=SPLIT, task="val", normalize=True, xymean=tr[0], xystd=tr[1],
                               edgemean=tr[2], edgestd=tr[3])
            b = next(iter(DataLoader(ds, batch_size=batch_size, shuffle=False)))
            out["val_with_train_stats"] = {"ids": torch.arange(min(batch_size, len(ds))), **{f: getattr(b, f).clone() for f in FIELDS}}
        torch.save(out, dataset_golden_path(name))
        print(f"{name}: " + ", ".join(f"{t}={v['len']} samples / {len(v['batches'])} batches" for t, v in out["splits"].items()) +
              f" -> {os.path.getsize(dataset_golden_path(name))/1024:.1f} KiB")


if __name__ == "__main__":
    main()
