"""Stand-alone timing of the memory-bound edge kernels through the C ABI (CUDA events on the launching stream):
warm = back-to-back launches on the same buffers (L2-resident at case118 sizes), cold = L2 flushed (write of a
256 MB buffer) before every launch.  Prints one JSON line per kernel/config; used for profiles/*.md."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from poweflownet_b200 import ops  # noqa: E402
from poweflownet_b200.data import synthetic_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--case", default="118v2")
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--hidden", type=int, default=129)
ap.add_argument("--iters", type=int, default=50)
args = ap.parse_args()
dev = torch.device("cuda", 0)
peak = bench.load_peaks()["hbm_gbs"]
b = synthetic_batch(args.case, args.batch).to(dev)
n, h = b.num_nodes, args.hidden
g = ops.PreparedGraph(b.edge_index, b.edge_attr, n, mode=1)
e = g.meta()[1]
hi, hj, s, ds, dhi, dhj = (ops.new_rows(n, h, dev).normal_() for _ in range(6))
w1 = torch.randn(h, 2 * h + 2, device=dev)
dw1 = torch.zeros_like(w1)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def run_graph(name, fns, nbytes, reps=40):
    """`reps` back-to-back launches inside one CUDA graph (no per-launch host cost); `fns` rotates over buffer sets whose
    total footprint exceeds L2 when len(fns) > 1 ("cold"), or reuses one set ("warm")."""
    for label, use in (("warm", fns[:1]), ("rotating", fns)):
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            for f in use:
                f()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(reps):
                use[i % len(use)]()
        times = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) * 1e3 / reps)
        times.sort()
        med = times[len(times) // 2]
        print(json.dumps({"kernel": name, "mode": "graph x%d" % reps, "l2": label, "sets": len(use), "us_per_launch": round(med, 2),
                          "GBps": round(nbytes / med / 1e3, 1), "frac_of_measured_hbm": round(nbytes / med / 1e3 / peak, 3),
                          "threads": os.environ.get("PFN_EDGE_THREADS"), "bps": os.environ.get("PFN_EDGE_BPS")}))


def run(name, fn, nbytes):
    for cold in (False, True):
        for _ in range(5):
            fn()
        times = []
        for _ in range(args.iters):
            if cold:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) * 1e3)
        times.sort()
        med = times[len(times) // 2]
        print(json.dumps({"kernel": name, "case": args.case, "batch": args.batch, "hidden": h, "N": n, "E": e,
                          "l2": "cold" if cold else "warm", "us_median": round(med, 2), "us_min": round(times[0], 2),
                          "algorithmic_bytes": nbytes, "GBps": round(nbytes / med / 1e3, 1), "frac_of_measured_hbm": round(nbytes / med / 1e3 / peak, 3)}))


ea_bytes = bench.ea_algorithmic_bytes(n, e, h)
hop_bytes = 2 * 4 * n * h + 4 * (n + 1) + 4 * e + 4 * n
if os.environ.get("PFN_BENCH_GRAPH"):
    nsets = max(2, int(200e6 // (3 * 4 * n * ops.round_up4(h))))
    sets = [tuple(ops.new_rows(n, h, dev).normal_() for _ in range(6)) for _ in range(min(nsets, 10))]
    run_graph("ea_fwd", [(lambda t=t: ops.ea_fwd(t[0], t[1], g, w1, h, h, t[2])) for t in sets], ea_bytes)
    run_graph("hop", [(lambda t=t: ops.spmm_hop(t[0], g, t[2], h)) for t in sets], hop_bytes)
    sys.exit(0)
run("ea_fwd", lambda: ops.ea_fwd(hi, hj, g, w1, h, h, s), ea_bytes)
run("hop", lambda: ops.spmm_hop(hi, g, s, h), hop_bytes)
scratch_holder = []
run("ea_bwd", lambda: ops.ea_bwd(ds, hi, hj, g, w1, h, h, dhi, dhj, dw1), 5 * 4 * n * h + 2 * (4 * (n + 1) + 12 * e))
