"""Per-tensor parity report of the CUDA model against the CPU oracle at the bench configuration."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import common
from oracle import pfn_oracle as O
from poweflownet_b200.data import synthetic_batch
from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
b = int(sys.argv[1]) if len(sys.argv) > 1 else 128
case = sys.argv[2] if len(sys.argv) > 2 else "118v2"
hid = int(sys.argv[3]) if len(sys.argv) > 3 else 129
nl = int(sys.argv[4]) if len(sys.argv) > 4 else 4
K = int(sys.argv[5]) if len(sys.argv) > 5 else 3
fused = (sys.argv[6] != "0") if len(sys.argv) > 6 else True
kw = dict(common.MODEL_DIMS, hidden_dim=hid, n_gnn_layers=nl, K=K, dropout_rate=0.0)
batch = synthetic_batch(case, b)
oracle = common.load_seeded(O.MaskEmbdMultiMPN(**kw)).train()
loss_ref, out_ref = O.forward_loss_backward(oracle, batch, "mse")
o64 = common.load_seeded(O.MaskEmbdMultiMPN(**kw)).double().train()
b64 = common.GraphBatch(**{k: (getattr(batch, k).double() if getattr(batch, k).is_floating_point() else getattr(batch, k)) for k in ("x","y","bus_type","pred_mask","edge_index","edge_attr","batch","ptr")})
O.forward_loss_backward(o64, b64, "mse")
m = MaskEmbdMultiMPN(**kw); m.load_state_dict(oracle.state_dict()); m = m.cuda().train()
m.fused = fused
print("env", {k: v for k, v in os.environ.items() if k.startswith("PFN_")}, "case", case, "b", b, "hidden", hid, "L", nl, "K", K, "fused", fused)
db = batch.to("cuda")
out = m(db); loss = torch.nn.functional.mse_loss(out, db.y); loss.backward()
print("out", common.rel_err(out.detach().cpu(), out_ref))
for (k, p), (_, q), (_, r) in zip(m.named_parameters(), oracle.named_parameters(), o64.named_parameters()):
    e = common.rel_err(p.grad.cpu(), q.grad); e64 = common.rel_err(p.grad.cpu().double(), r.grad); eo = common.rel_err(q.grad.double(), r.grad)
    flag = " <<<" if max(e) > 1e-5 else ""
    print(f"{k:36s} vs fp32 oracle {e[0]:.2e} {e[1]:.2e} | vs fp64 {e64[0]:.2e} {e64[1]:.2e} | oracle32 vs fp64 {eo[0]:.2e} {eo[1]:.2e}{flag}")
