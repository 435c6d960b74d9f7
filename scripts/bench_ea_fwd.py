"""Stand-alone timing of pfn_ea_fwd on the bench workload (same routine bench.py uses for its roofline figure)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from poweflownet_b200 import _lib
from poweflownet_b200.data import synthetic_batch

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
lib = _lib.lib()
case, b, h = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else ("118v2", 128, 129)
batch = synthetic_batch(case, b).to(dev)
us, n = bench.time_ea_fwd_alone(lib, dev, batch, h, n_sets=12 if case == "118v2" else 3)
by = bench.ea_algorithmic_bytes(batch.num_nodes, 2 * int(batch.edge_index.size(1)), h)
print(f"{case} x{b} h={h} BPS={os.environ.get('PFN_EDGE_BPS')} THREADS={os.environ.get('PFN_EDGE_THREADS')}: {us:.2f} us/launch, {by / us / 1e3:.0f} GB/s")
