"""A/B timing of the two EdgeAggregation forward kernels (PFN_EA_FWD=cta|warp) on the bench workload, one process,
same routine as bench.py's roofline figure.  Prints one JSON object."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from poweflownet_b200 import _lib  # noqa: E402
from poweflownet_b200.data import synthetic_batch  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
lib = _lib.lib()
batch = synthetic_batch("118v2", 128).to(dev)
by = bench.ea_algorithmic_bytes(batch.num_nodes, 2 * int(batch.edge_index.size(1)), 129)
peak, _ = bench.load_peaks()
out = {"algorithmic_bytes": by, "hbm_peak_gbs": peak}
for rep in range(2):
    for which in ("cta", "warp"):
        os.environ["PFN_EA_FWD"] = which
        us, n = bench.time_ea_fwd_alone(lib, dev, batch, 129)
        out[f"{which}_{rep}"] = {"us_per_launch": us, "gbs": by / us / 1e3, "frac": by / us / 1e3 / peak}
print(json.dumps(out))
