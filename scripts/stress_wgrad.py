"""200-repeat bitwise-determinism stress of the whole backward (grouped weight gradient k_wgrad_group included) at the
bench size (case118v2 x 128, configs/standard.json) and of the layer-wise route (hidden 512): every repeat must
reproduce the first gradients bit for bit.  Companion to the racecheck logs in profiles/."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import common
from poweflownet_b200.data import synthetic_batch
from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
from poweflownet_b200.training import fused_mse_step

dev = torch.device("cuda", 0)
for case, b, hid, layers, reps in (("118v2", 128, 129, 4, 200), ("6470rte", 2, 512, 3, 40)):
    kw = dict(common.MODEL_DIMS, hidden_dim=hid, n_gnn_layers=layers, K=3, dropout_rate=0.0)
    m = common.load_seeded(MaskEmbdMultiMPN(**kw)).to(dev).train()
    batch = synthetic_batch(case, b).to(dev)
    fused_mse_step(m, batch)
    ref = [p.grad.clone() for p in m.parameters()]
    bad = 0
    for i in range(reps):
        fused_mse_step(m, batch)
        bad += sum(0 if torch.equal(a, p.grad) else 1 for a, p in zip(ref, m.parameters()))
    torch.cuda.synchronize()
    print(f"{case} x {b}, hidden {hid}, {layers} GNN layers: {reps} repeats, {bad} gradient tensors differed from the first run")
    assert bad == 0
print("bitwise deterministic")
