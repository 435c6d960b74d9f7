"""Phase stamps of k_wgrad_group's CTA 0 (PFN_WG_TIMING) for one training step at the bench size (case118v2 x 128, standard.json)."""
import os, sys
os.environ["PFN_WG_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import common
from poweflownet_b200.data import synthetic_batch
from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
from poweflownet_b200.training import fused_mse_step

dev = torch.device("cuda", 0)
kw = dict(common.MODEL_DIMS, hidden_dim=129, n_gnn_layers=4, K=3, dropout_rate=0.2)
model = common.load_seeded(MaskEmbdMultiMPN(**kw)).to(dev).train()
batch = synthetic_batch("118v2", 128).to(dev)
for _ in range(3):
    fused_mse_step(model, batch)
torch.cuda.synchronize()
print("done")
