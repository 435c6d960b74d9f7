"""Profiling harness for ncu: W warm-up steps, then K steps of the bench workload between
cudaProfilerStart/Stop (run ncu with --profile-from-start off).  Not a benchmark: never quote its timing."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import torch  # noqa: E402

import bench  # noqa: E402
import common  # noqa: E402
from poweflownet_b200.data import synthetic_batch  # noqa: E402
from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN  # noqa: E402
from poweflownet_b200.training import fused_mse_step  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--case", default=bench.CASE)
ap.add_argument("--batch", type=int, default=bench.BATCH)
ap.add_argument("--hidden", type=int, default=bench.MODEL_KW["hidden_dim"])
ap.add_argument("--layers", type=int, default=bench.MODEL_KW["n_gnn_layers"])
args = ap.parse_args()
dev = torch.device("cuda", 0)
kw = dict(bench.MODEL_KW, hidden_dim=args.hidden, n_gnn_layers=args.layers)
model = common.load_seeded(MaskEmbdMultiMPN(**kw)).to(dev).train()
batches = [synthetic_batch(args.case, args.batch, seed=1234 + i).to(dev) for i in range(2)]
for i in range(args.warmup):
    fused_mse_step(model, batches[i % 2])
torch.cuda.synchronize()
torch.cuda.profiler.start()
for i in range(args.steps):
    fused_mse_step(model, batches[i % 2])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", args.steps, "steps")
