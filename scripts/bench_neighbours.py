#!/usr/bin/env python
"""Time the kernels either side of the hot path (SURVEY.md section 8 f) alone on one B200: batch assembly from the
device-resident dataset, PowerImbalance loss + gradient, one-launch AdamW.  CUDA events around back-to-back calls,
after warm-up; algorithmic bytes as stated in DESIGN.md section 5.4.  Prints one JSON object.

    python scripts/bench_neighbours.py [--iters 200] > gpurun_out/neighbours.json
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def timed(fn, iters):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / iters  # us per call


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    a = ap.parse_args()
    import common
    from poweflownet_b200 import ops
    from poweflownet_b200._lib import lib
    from poweflownet_b200.data import synthetic_raw_case
    from poweflownet_b200.datasets import PowerFlowData
    from poweflownet_b200.losses import PowerImbalance
    from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
    from poweflownet_b200.optim import FusedAdamW
    dev = torch.device("cuda", 0)
    peak = 6548.8
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    out = {"hbm_peak_gbs": peak, "iters": a.iters, "workload": "case118v2 x 128 (N=15104, E_raw=23808), standard.json"}
    B = 128
    ds = PowerFlowData(case="118v2", split=[.5, .2, .3], task="train", device=dev, raw=[synthetic_raw_case("118v2", 4000, seed=7)])
    order = torch.randperm(len(ds), generator=torch.Generator().manual_seed(0))
    order_dev, order_host = order.to(dev), order.numpy()
    static = ds.batch(order_host[:B])
    k = [0]

    def assemble():
        i = k[0] % (len(ds) // B)
        k[0] += 1
        ds.batch(order_host[i * B:(i + 1) * B], ids_device=order_dev[i * B:(i + 1) * B], out=static)
    n, e_raw = static.num_nodes, int(static.edge_index.size(1))
    us = timed(assemble, a.iters)
    nbytes = n * (24 + 104) + e_raw * (16 + 24)
    out["batch_assemble"] = {"us_per_call": us, "launches_per_call": 2, "algorithmic_bytes": nbytes,
                             "gbs": nbytes / us / 1e3, "frac_of_hbm_peak": nbytes / us / 1e3 / peak,
                             "note": "includes the host side of datasets.PowerFlowData.batch (id bookkeeping, scratch allocation)"}
    stats = ds.get_data_means_stds()
    fn = PowerImbalance(*stats)
    graph = ops.PreparedGraph(static.edge_index, static.edge_attr, n, mode=1)
    x = static.y.clone().requires_grad_(True)
    us = timed(lambda: fn(x, static.edge_index, static.edge_attr, graph), a.iters)
    e = 2 * e_raw
    # k_pi_node: own row 16N + gathered rows 16E + CSR 12E + 4N + (dP,dQ) 8N; k_pi_grad: own 16N + dpq 8N + two CSR walks
    # 2 x (16E + 12E) + 8N row pointers + gathered dpq 8E + gradient 16N
    nbytes = 76 * n + 92 * e
    out["power_imbalance_fwd_bwd"] = {"us_per_call": us, "launches_per_call": 3, "algorithmic_bytes": nbytes,
                                      "gbs": nbytes / us / 1e3, "frac_of_hbm_peak": nbytes / us / 1e3 / peak}
    model = common.load_seeded(MaskEmbdMultiMPN(nfeature_dim=4, efeature_dim=2, output_dim=4, hidden_dim=129, n_gnn_layers=4, K=3,
                                                dropout_rate=0.2)).to(dev)
    for p in model.parameters():
        p.grad = torch.randn_like(p) * 0.01
    opt = FusedAdamW(model.parameters(), lr=1e-3)
    l0 = lib().pfn_launch_count()
    us = timed(opt.step, a.iters)
    n_par = sum(p.numel() for p in model.parameters())
    nbytes = n_par * 4 * 7
    out["adamw_step"] = {"us_per_call": us, "launches_per_call": (lib().pfn_launch_count() - l0) / (a.iters + 10),
                         "parameters": n_par, "tensors": len(list(model.parameters())), "algorithmic_bytes": nbytes,
                         "gbs": nbytes / us / 1e3, "frac_of_hbm_peak": nbytes / us / 1e3 / peak}
    ref = [torch.nn.Parameter(p.detach().clone()) for p in model.parameters()]
    for p in ref:
        p.grad = torch.randn_like(p) * 0.01
    topt = torch.optim.AdamW(ref, lr=1e-3)
    out["torch_adamw_step_us"] = timed(topt.step, a.iters)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
