"""Step time of BASELINE configs[3]-like workloads on the layer-wise route (case6470rte, configs/large.json)."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import common
from poweflownet_b200 import _lib
from poweflownet_b200.data import synthetic_batch
from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
from poweflownet_b200.training import fused_mse_step

b = int(sys.argv[1]) if len(sys.argv) > 1 else 32
hid = int(sys.argv[2]) if len(sys.argv) > 2 else 512
nl = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda", 0)
kw = dict(common.MODEL_DIMS, hidden_dim=hid, n_gnn_layers=nl, K=3, dropout_rate=0.2)
model = common.load_seeded(MaskEmbdMultiMPN(**kw)).to(dev).train()
batch = synthetic_batch("6470rte", b).to(dev)
lib = _lib.lib()
for _ in range(2):
    fused_mse_step(model, batch)
torch.cuda.synchronize()
lib.pfn_profile_enable(1)
n = 3
t = time.time()
for _ in range(n):
    fused_mse_step(model, batch)
torch.cuda.synchronize()
dt = (time.time() - t) / n
names = ["ea_fwd", "ea_bwd", "hop", "gemm_fwd", "gemm_dgrad", "gemm_wgrad", "prep", "fused_fwd"]
out = {}
for cat, name in enumerate(names):
    tot, cnt = C.c_double(), C.c_int64()
    lib.pfn_profile_read(cat, C.byref(tot), C.byref(cnt))
    out[name] = (round(tot.value / n, 2), cnt.value // n)
print(f"case6470rte x{b} hidden {hid} L={nl}: {dt * 1e3:.1f} ms/step ({b / dt:.1f} graphs/s); per category (ms, launches): {out}")
