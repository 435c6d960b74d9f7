"""Tiny forward + backward through the graph-resident kernels, the layer-wise kernels (bulk-copy EdgeAggregation forward,
register-flushed tensor-core GEMMs, grouped weight gradient) for compute-sanitizer memcheck / racecheck / initcheck."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import common
from poweflownet_b200.data import synthetic_batch
from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
from poweflownet_b200.training import fused_mse_step

dev = torch.device("cuda", 0)
for case, b, hid, layers, fused in (("118v2", 2, 129, 3, True), ("14", 11, 64, 3, True), ("118v2", 2, 129, 3, False), ("14", 6, 512, 2, False)):
    kw = dict(common.MODEL_DIMS, hidden_dim=hid, n_gnn_layers=layers, K=3, dropout_rate=0.2)
    m = common.load_seeded(MaskEmbdMultiMPN(**kw)).to(dev).train()
    m.fused = fused
    batch = synthetic_batch(case, b).to(dev)
    loss = fused_mse_step(m, batch)
    torch.cuda.synchronize()
    print(case, b, hid, "fused" if fused else "layer-wise", float(loss), m._tiling_checked)
