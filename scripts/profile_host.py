"""cProfile of the eager fused_mse_step loop (host side): where the Python/ctypes time of a step goes."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import common
import bench
from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
from poweflownet_b200.training import fused_mse_step
dev = torch.device("cuda", 0)
cfg = bench.CONFIGS["standard"]
model = common.load_seeded(MaskEmbdMultiMPN(**cfg["model"])).to(dev).train()
batches = [bench.make_batch(cfg, 1234 + i).to(dev) for i in range(8)]
for i in range(20):
    fused_mse_step(model, batches[i % 8])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(200):
    fused_mse_step(model, batches[i % 8])
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"200 steps: host issue {1e3 * t_issue / 200:.3f} ms/step, wall {1e3 * t_all / 200:.3f} ms/step")
pr = cProfile.Profile()
pr.enable()
for i in range(200):
    fused_mse_step(model, batches[i % 8])
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
