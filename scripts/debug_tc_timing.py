"""Phase stamps of k_gemm_tc's CTA 0 (PFN_TC_TIMING) for every tensor-core GEMM of one forward + backward at
case6470rte x B, hidden 512: python scripts/debug_tc_timing.py [B] 2> stamps.log"""
import os, sys
if os.environ.get("TC_TIMING", "1") == "1":
    os.environ["PFN_TC_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import common
from poweflownet_b200.data import synthetic_batch
from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
from poweflownet_b200.training import fused_mse_step

b = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda", 0)
kw = dict(common.MODEL_DIMS, hidden_dim=512, n_gnn_layers=2, K=3, dropout_rate=0.2)
model = common.load_seeded(MaskEmbdMultiMPN(**kw)).to(dev).train()
batch = synthetic_batch("6470rte", b).to(dev)
fused_mse_step(model, batch)
torch.cuda.synchronize()
print("done")
