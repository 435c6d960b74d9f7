"""Tiny forward + backward through the graph-resident kernels (for compute-sanitizer memcheck / racecheck / initcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import common
from poweflownet_b200.data import synthetic_batch
from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
from poweflownet_b200.training import fused_mse_step

dev = torch.device("cuda", 0)
for case, b, hid in (("118v2", 2, 129), ("14", 11, 64)):
    kw = dict(common.MODEL_DIMS, hidden_dim=hid, n_gnn_layers=3, K=3, dropout_rate=0.2)
    m = common.load_seeded(MaskEmbdMultiMPN(**kw)).to(dev).train()
    batch = synthetic_batch(case, b).to(dev)
    loss = fused_mse_step(m, batch)
    torch.cuda.synchronize()
    print(case, b, hid, float(loss), m._tiling_checked)
