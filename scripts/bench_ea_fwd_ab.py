"""A/B timing of the EdgeAggregation forward kernels on the two roofline workloads:
case118v2 x 128 at hidden 129 (24.0 MB algorithmic) and case6470rte x 32 at hidden 512 (1.28 GB).  Variants = the CTA-slab
kernel (PFN_EA_FWD=cta) and the bulk-copy kernel with its tuning knobs.  Two timing modes per variant: `iters` eager
back-to-back launches between one pair of CUDA events, and the same launches captured once into a CUDA graph and replayed
(no host launch cost).  Operand sets rotate so that no launch finds its operands in L2.  Prints one JSON object per line.

    python scripts/bench_ea_fwd_ab.py [small|large|both]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from poweflownet_b200 import _lib, ops  # noqa: E402
from poweflownet_b200.data import synthetic_batch  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
lib = _lib.lib()
peak = bench.load_peaks()["hbm_gbs"]
KNOBS = ("PFN_EA_FWD", "PFN_EA_STAGES", "PFN_EA_THREADS", "PFN_EA_PREFETCH", "PFN_EA_PRODUCERS", "PFN_EA_BULK", "PFN_EA_CHUNK", "PFN_EDGE_CYCLIC", "PFN_EA_ROUND", "PFN_EA_BULK8", "PFN_EA_CTAS_PER_SM")


def run(case, b, h, variants, iters, n_sets):
    batch = synthetic_batch(case, b).to(dev)
    n, ld = batch.num_nodes, (h + 3) // 4 * 4
    g = ops.PreparedGraph(batch.edge_index, batch.edge_attr, n, mode=1)
    gen = torch.Generator(device=dev).manual_seed(7)
    sets = [(torch.randn(n, ld, device=dev, generator=gen), torch.randn(n, ld, device=dev, generator=gen),
             torch.empty(n, ld, device=dev)) for _ in range(n_sets)]
    we = torch.randn(h, 2, device=dev, generator=gen)
    by = bench.ea_algorithmic_bytes(n, 2 * int(batch.edge_index.size(1)), h)

    def launch(i, stream):
        hi, hj, s = sets[i % n_sets]
        _lib.check(lib.pfn_ea_fwd(hi.data_ptr(), hj.data_ptr(), ld, g.ws.data_ptr(), n, g.e_raw, we.data_ptr(), 2,
                                  s.data_ptr(), ld, h, stream), "pfn_ea_fwd")

    # calibration: an elementwise kernel that moves the same node-matrix bytes (read Hi, read Hj, write S) with perfectly
    # regular accesses -- torch.add(Hi, Hj, out=S) -- timed the same way: what "24 MB per launch, back to back" can reach at all
    if only is None or "torch_add" in only:
        for i in range(n_sets):
            torch.add(sets[i][0], sets[i][1], out=sets[i][2])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            hi, hj, s_ = sets[i % n_sets]
            torch.add(hi, hj, out=s_)
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / iters
        gcal = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            with torch.cuda.graph(gcal, stream=side):
                for i in range(iters):
                    hi, hj, s_ = sets[i % n_sets]
                    torch.add(hi, hj, out=s_)
        gcal.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            gcal.replay()
        e1.record()
        torch.cuda.synchronize()
        us_g = 1e3 * e0.elapsed_time(e1) / (3 * iters)
        nb = 3 * n * ld * 4
        print(json.dumps({"workload": f"{case} x {b}, hidden {h}", "variant": "torch_add (calibration: S = Hi + Hj, same node-matrix bytes)",
                          "us_eager": round(us, 3), "us_graph": round(us_g, 3), "bytes": nb, "frac_eager": round(nb / us / 1e3 / peak, 4),
                          "frac_graph": round(nb / us_g / 1e3 / peak, 4)}), flush=True)
    ref = None
    for name, env in variants:
        for k in KNOBS:
            os.environ.pop(k, None)
        os.environ.update(env)
        cur = torch.cuda.current_stream().cuda_stream
        for i in range(n_sets):
            launch(i, cur)
        torch.cuda.synchronize()
        out = sets[0][2].clone()
        if ref is None:
            ref = out
        same = bool(torch.equal(ref, out))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            launch(i, cur)
        e1.record()
        torch.cuda.synchronize()
        us_eager = 1e3 * e0.elapsed_time(e1) / iters
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                for i in range(iters):
                    launch(i, side.cuda_stream)
        graph.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        us_graph = 1e3 * e0.elapsed_time(e1) / (3 * iters)
        print(json.dumps({"workload": f"{case} x {b}, hidden {h}", "variant": name, "env": env, "bit_identical_to_first": same,
                          "us_eager": round(us_eager, 3), "us_graph": round(us_graph, 3), "algorithmic_bytes": by,
                          "frac_eager": round(by / us_eager / 1e3 / peak, 4), "frac_graph": round(by / us_graph / 1e3 / peak, 4)}), flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else "both"
only = set(sys.argv[2].split(",")) if len(sys.argv) > 2 else None  # optional: comma-separated variant names
P = {"PFN_EA_FWD": "tma"}
small = [("cta", {"PFN_EA_FWD": "cta"}), ("tma", P), ("tma_p10", {**P, "PFN_EA_PRODUCERS": "10"}), ("tma_p12", {**P, "PFN_EA_PRODUCERS": "12"}), ("tma_p16", {**P, "PFN_EA_PRODUCERS": "16"}), ("tma_1sm", {**P, "PFN_EA_CTAS_PER_SM": "1"}), ("tma_2sm_t128", {**P, "PFN_EA_THREADS": "128"}),
         ("tma_2sm_t384_p4", {**P, "PFN_EA_THREADS": "384"}), ("tma_2sm_p2", {**P, "PFN_EA_PRODUCERS": "2"}), ("tma_2sm_p6", {**P, "PFN_EA_PRODUCERS": "6"}),
         ("tma_2sm_s1", {**P, "PFN_EA_STAGES": "1"}), ("tma_2sm_s3", {**P, "PFN_EA_STAGES": "3"}), ("tma_2sm_prefetch", {**P, "PFN_EA_PREFETCH": "1"})]
large = [("cta", {"PFN_EA_FWD": "cta"}), ("tma", P), ("tma_c32", {**P, "PFN_EA_CHUNK": "32"}), ("tma_c64", {**P, "PFN_EA_CHUNK": "64"}), ("tma_c96", {**P, "PFN_EA_CHUNK": "96"}),
         ("tma_c192", {**P, "PFN_EA_CHUNK": "192"}), ("tma_c256", {**P, "PFN_EA_CHUNK": "256"}), ("tma_c64_s3", {**P, "PFN_EA_CHUNK": "64", "PFN_EA_STAGES": "3"}),
         ("tma_p6", {**P, "PFN_EA_PRODUCERS": "6"}), ("tma_t640", {**P, "PFN_EA_THREADS": "640"}),
         ("tma_p10", {**P, "PFN_EA_PRODUCERS": "10"}), ("tma_p12", {**P, "PFN_EA_PRODUCERS": "12"}), ("tma_p14", {**P, "PFN_EA_PRODUCERS": "14"}), ("tma_p16", {**P, "PFN_EA_PRODUCERS": "16"}),
         ("tma_p12_t384", {**P, "PFN_EA_PRODUCERS": "12", "PFN_EA_THREADS": "384"}), ("tma_p16_t256", {**P, "PFN_EA_PRODUCERS": "16", "PFN_EA_THREADS": "256"}), ("tma_2sm", {**P, "PFN_EA_CTAS_PER_SM": "2"})]
if only is not None:
    small = [v for v in small if v[0] in only]
    large = [v for v in large if v[0] in only]
if which in ("small", "both") and small:
    run("118v2", 128, 129, small, iters=int(os.environ.get("AB_ITERS", "240")), n_sets=12)
if which in ("large", "both") and large:
    run("6470rte", 32, 512, large, iters=12, n_sets=2)
