"""Compare the graph-resident forward (pfn_mpn_forward_tiled) with the layer-wise forward on the GPU."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import common
from poweflownet_b200.data import synthetic_batch
from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
from poweflownet_b200.training import fused_mse_step

dev = torch.device("cuda", 0)
def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

for case, bs, kw in [("118v2", 4, dict(hidden_dim=129, n_gnn_layers=4, K=3)), ("14", 16, dict(hidden_dim=64, n_gnn_layers=2, K=3)),
                     ("118v2", 128, dict(hidden_dim=129, n_gnn_layers=4, K=3))]:
    full = dict(nfeature_dim=4, efeature_dim=2, output_dim=4, dropout_rate=0.2, **kw)
    torch.manual_seed(0)
    model = common.load_seeded(MaskEmbdMultiMPN(**full)).to(dev)
    batch = synthetic_batch(case, bs).to(dev)
    for mode in ("eval", "train"):
        model.train(mode == "train")
        masks = None
        if mode == "train":
            n_act = 2 * kw["n_gnn_layers"] - 2
            masks = [(torch.rand(batch.num_nodes, kw["hidden_dim"], generator=torch.Generator().manual_seed(i)) < 0.8).float() for i in range(n_act)]
        model._inject_dropout_masks = masks
        outs, acts, grads = {}, {}, {}
        for fused in (False, True):
            model.fused = fused
            model._pool.clear()
            with torch.enable_grad():
                out = model(batch)
            torch.cuda.synchronize()
            outs[fused] = out.detach().clone()
            loss = fused_mse_step(model, batch)
            torch.cuda.synchronize()
            grads[fused] = [p.grad.clone() for p in model._engine_params()]
        print(f"{case} x{bs} {kw} {mode}: out rel err fused vs layerwise = {rel(outs[True], outs[False]):.3e}  "
              f"nan={bool(torch.isnan(outs[True]).any())}  tiling={model._tiling_checked}")
        print("   worst grad rel err:", max(rel(a, b) for a, b in zip(grads[True], grads[False])))
# timing
model.train(); model._inject_dropout_masks = None
for fused in (False, True):
    model.fused = fused
    for _ in range(5): fused_mse_step(model, batch)
    torch.cuda.synchronize(); t = time.time()
    for _ in range(50): fused_mse_step(model, batch)
    torch.cuda.synchronize(); print("fused" if fused else "layerwise", "ms/step", (time.time() - t) / 50 * 1e3)
# kernel time of the fused forward via the library's event hooks
import ctypes as C
from poweflownet_b200 import _lib
lib = _lib.lib()
model.fused = True
lib.pfn_profile_enable(1)
for _ in range(20): fused_mse_step(model, batch)
torch.cuda.synchronize()
tot, cnt = C.c_double(), C.c_int64()
lib.pfn_profile_read(7, C.byref(tot), C.byref(cnt))
print("fused fwd kernel: %.1f us avg over %d launches" % (1e3 * tot.value / max(cnt.value, 1), cnt.value))
lib.pfn_profile_enable(0)
