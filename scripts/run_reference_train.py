#!/usr/bin/env python
"""Run the reference's OWN `train.py` byte for byte, with its missing dependency (torch_geometric, matplotlib)
satisfied by the stand-in under oracle/pyg_shim, optionally with `networks.MPN.MaskEmbdMultiMPN` swapped for the
sm_100a implementation (the drop-in demonstration of INTEGRATION.md).

    python scripts/run_reference_train.py [--reference /root/reference] [--impl reference|b200] [--workdir DIR]
           [--make-synthetic-case 14 --samples 64] -- <train.py arguments>

`train.py` is executed with runpy from `--workdir` (it writes logs/, models/ and <data-dir>/params relative to
the current directory / data dir).  Test infrastructure: needs the reference checkout, which exists only in the
build container.
"""
from __future__ import annotations

import argparse
import os
import runpy
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_synthetic_raw(data_dir: str, case: str, samples: int, seed: int = 0) -> None:
    """`<data_dir>/raw/case<case>_{edge,node}_features.npy` in the reference's format
    (datasets/PowerFlowData.py:58-61,178-204): edges [S, E, 4] = (from, to, r, x); nodes [S, n, 6] = (index, type,
    Vm, Va, P, Q)."""
    import numpy as np
    import torch

    sys.path.insert(0, ROOT)
    from poweflownet_b200.data import CASES, synthetic_topology

    n, e_raw = CASES[case]
    rng = np.random.default_rng(seed)
    topo = synthetic_topology(n, e_raw).numpy().T.astype(np.float64)  # [E, 2]
    edges = np.zeros((samples, e_raw, 4))
    edges[:, :, :2] = topo[None]
    edges[:, :, 2:] = np.abs(rng.normal(0.05, 0.02, size=(samples, e_raw, 2))) + 1e-3
    nodes = np.zeros((samples, n, 6))
    nodes[:, :, 0] = np.arange(n)[None]
    types_ = np.where(rng.random(n) < 0.45, 1, 2)
    types_[0] = 0
    nodes[:, :, 1] = types_[None]
    nodes[:, :, 2] = 1.0 + rng.normal(0, 0.02, size=(samples, n))
    nodes[:, :, 3] = rng.normal(0, 5.0, size=(samples, n))
    nodes[:, :, 4:] = rng.normal(0, 30.0, size=(samples, n, 2))
    os.makedirs(os.path.join(data_dir, "raw"), exist_ok=True)
    np.save(os.path.join(data_dir, "raw", f"case{case}_edge_features.npy"), edges)
    np.save(os.path.join(data_dir, "raw", f"case{case}_node_features.npy"), nodes)
    del torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("PFN_REFERENCE", "/root/reference"))
    ap.add_argument("--impl", choices=["reference", "b200"], default="reference")
    ap.add_argument("--workdir", default=".")
    ap.add_argument("--make-synthetic-case", default=None)
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("rest", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    train_py = os.path.join(a.reference, "train.py")
    if not os.path.exists(train_py):
        raise SystemExit(f"{train_py} not found")
    rest = [x for x in a.rest if x != "--"]
    os.makedirs(a.workdir, exist_ok=True)
    os.chdir(a.workdir)
    os.makedirs(os.path.join("logs", "train_log"), exist_ok=True)  # train.py:182 saves there before creating it (:201)
    if a.make_synthetic_case:
        data_dir = "data"
        for i, tok in enumerate(rest):
            if tok == "--data-dir":
                data_dir = rest[i + 1]
        make_synthetic_raw(data_dir, a.make_synthetic_case, a.samples)
    os.environ.setdefault("TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD", "1")  # train.py torch.load()s pickled Data / Namespace
    os.environ.setdefault("WANDB_MODE", "disabled")
    for p in (os.path.join(ROOT, "oracle", "pyg_shim"), a.reference):
        if p not in sys.path:
            sys.path.insert(0, p)
    if a.impl == "b200":
        sys.path.insert(0, ROOT)
        import networks.MPN as ref_mpn  # the reference module: keeps the six other model classes train.py imports
        from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN

        swapped = types.ModuleType("networks.MPN")
        swapped.__dict__.update({k: v for k, v in ref_mpn.__dict__.items() if not k.startswith("__")})
        swapped.MaskEmbdMultiMPN = MaskEmbdMultiMPN
        sys.modules["networks.MPN"] = swapped
        import networks
        networks.MPN = swapped
    sys.argv = [train_py] + rest
    runpy.run_path(train_py, run_name="__main__")


if __name__ == "__main__":
    main()
