#!/usr/bin/env python
"""Workload for the ncu launch list of the whole training loop (SURVEY.md section 8 f): two optimisation steps of
`training.train_epoch` with MSE and two with PowerImbalance on a device-resident case118v2 dataset, FusedAdamW.

    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_epoch.csv \
        python scripts/profile_epoch.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def main():
    import common
    from poweflownet_b200.data import synthetic_raw_case
    from poweflownet_b200.datasets import PowerFlowData
    from poweflownet_b200.losses import PowerImbalance
    from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
    from poweflownet_b200.optim import FusedAdamW
    from poweflownet_b200.training import train_epoch
    dev = torch.device("cuda", 0)
    ds = PowerFlowData(case="118v2", split=[.5, .2, .3], task="train", device=dev, raw=[synthetic_raw_case("118v2", 520, seed=7)])
    model = common.load_seeded(MaskEmbdMultiMPN(nfeature_dim=4, efeature_dim=2, output_dim=4, hidden_dim=129, n_gnn_layers=4,
                                                K=3, dropout_rate=0.2)).to(dev)
    opt = FusedAdamW(model.parameters(), lr=1e-3)
    g = torch.Generator().manual_seed(0)
    a = train_epoch(model, ds.loader(128, shuffle=True, generator=g, drop_last=True), torch.nn.MSELoss(), opt, dev)
    b = train_epoch(model, ds.loader(128, shuffle=True, generator=g, drop_last=True), PowerImbalance(*ds.get_data_means_stds()), opt, dev)
    torch.cuda.synchronize()
    print(f"mse epoch loss {a:.5f}; power-imbalance epoch loss {b:.5f}; {len(ds)} samples")


if __name__ == "__main__":
    main()
