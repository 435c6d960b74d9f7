#!/usr/bin/env python
"""bench.py -- graphs/sec of the PowerFlowNet hot path (MaskEmbdMultiMPN fwd + MSE + bwd) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config standard|large|mixed]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workloads (BASELINE.json `configs`):
  standard (default, configs[1]/[2])  case118v2-shaped synthetic graphs (118 buses, 186 branches), batch 128 PER GPU (weak
            scaling), configs/standard.json (hidden 129, 4 GNN layers, K=3, dropout 0.2), train mode.
  large    (configs[3])  case6470rte x 32 per GPU, configs/large.json (hidden 512, 5 GNN layers, K=3).
  mixed    (configs[4])  ONE global batch of 64 graphs drawn 40:20:4 from case14 / case118v2 / case6470rte,
            configs/extra_large.json (hidden 512, 10 GNN layers), sharded over the ranks by branch count (strong scaling).
One "step" = graph prep + forward + MSE loss + backward over one resident mini-batch (+ the single gradient all-reduce
when N > 1); `value` = graphs of all ranks / max-over-ranks device time.  `e2e` is the same step through the public
API from pinned HOST memory (H2D of the batch and the D2H read of the loss inside the timed region).  `roofline` is the
fused EdgeAggregation message+aggregate forward kernel: algorithmic bytes / its mean device time
over back-to-back launches measured with CUDA events in this run; `roofline_step` lists the kernels the timed step
actually consists of.  `--impl reference` times the reference's CPU path (the oracle restatement -- the reference itself
needs torch_geometric, absent here) on the box's host cores.  The default N=1 run also carries, as extra keys, the
configs[3] workload (`configs3`), the same oracle in eager torch CUDA on this GPU (`reference_gpu`) and whole training
epochs (`train_epoch`); N>1 runs carry `dp_parity` and a strong-scaling leg.  Prints exactly one JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# stdout must carry exactly one JSON line: NCCL prints its "NCCL version ..." banner to stdout at every debug level from
# VERSION up (WARN included), and this image runs with the VERSION level; silence it unless the caller asked for INFO/TRACE
if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):
    os.environ["NCCL_DEBUG"] = "NONE"

_REAL_STDOUT = None


def _claim_stdout():
    """Route everything any library writes to file descriptor 1 (NCCL banners, stray prints) to stderr and keep the real
    stdout for the single JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line: dict):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


UNIT = "graphs/s"
DIMS = dict(nfeature_dim=4, efeature_dim=2, output_dim=4)
CONFIGS = {
    "standard": dict(
        metric="graphs/sec (case118v2, batch 128) fwd+bwd", case="118v2", batch=128, scaling="weak",
        model=dict(DIMS, hidden_dim=129, n_gnn_layers=4, K=3, dropout_rate=0.2),
        workload="case118v2 MaskEmbdMultiMPN configs/standard.json batch=128 per GPU (BASELINE configs[1]); train mode, MSE loss"),
    "large": dict(
        metric="graphs/sec (case6470rte, batch 32) fwd+bwd", case="6470rte", batch=32, scaling="weak",
        model=dict(DIMS, hidden_dim=512, n_gnn_layers=5, K=3, dropout_rate=0.2),
        workload="case6470rte MaskEmbdMultiMPN configs/large.json batch=32 per GPU (BASELINE configs[3]); train mode, MSE loss"),
    "mixed": dict(
        metric="graphs/sec (mixed 14/118/6470, global batch 64) fwd+bwd", case=None, batch=64, scaling="strong",
        mix=(("14", 40), ("118v2", 20), ("6470rte", 4)),
        model=dict(DIMS, hidden_dim=512, n_gnn_layers=10, K=3, dropout_rate=0.2),
        workload="mixed-case generalizer (40 x case14, 20 x case118v2, 4 x case6470rte) MaskEmbdMultiMPN configs/extra_large.json "
                 "global batch=64 sharded over the GPUs by branch count (BASELINE configs[4]); train mode, MSE loss"),
}
N_ROTATE = 8  # distinct resident batches cycled through the timed steps
# names the helper scripts use
CASE, BATCH, MODEL_KW = CONFIGS["standard"]["case"], CONFIGS["standard"]["batch"], CONFIGS["standard"]["model"]


def config_dict(cfg):
    """The `config` object of the JSON line -- identical for both arms (`--impl ours` / `--impl reference`)."""
    return {"workload": cfg["workload"], **{k: cfg["model"][k] for k in ("hidden_dim", "n_gnn_layers", "K", "dropout_rate")}}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", 1400.0)),
                "source": "measured (MEASURED_PEAKS.json: hbm_gbs burst copy; bf16_tflops_sustained)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md: 6.65 TB/s, ~1.4 PFLOP/s sustained)"}


def mixed_cases(cfg, seed=0):
    """The 64 graph names of the configs[4] batch in a fixed seeded order (variable-N batching)."""
    import torch
    names = [n for n, c in cfg["mix"] for _ in range(c)]
    perm = torch.randperm(len(names), generator=torch.Generator().manual_seed(seed)).tolist()
    return [names[i] for i in perm]


def make_batch(cfg, seed, rank=0, world=1, strong=False):
    """This rank's mini-batch (host).  weak: `batch` graphs per rank; strong / mixed: one global batch sharded over ranks."""
    from poweflownet_b200.data import shard_batch, synthetic_batch
    if cfg["case"] is None:
        return shard_batch(synthetic_batch(cases=mixed_cases(cfg), seed=seed), rank, world)
    if strong:
        return shard_batch(synthetic_batch(cfg["case"], cfg["batch"], seed=seed), rank, world)
    return synthetic_batch(cfg["case"], cfg["batch"], seed=seed + 1000 * rank)


def fused_saved_bytes(n_nodes: int, kw) -> int:
    """Activations the graph-resident forward writes for the backward pass: Hi, Hj, S per EdgeAggregation,
    [x_0..x_K] and Y per TAGConv, t1/x0/maskf of mask_embd, the output."""
    h, L, K = kw["hidden_dim"], kw["n_gnn_layers"], kw["K"]
    ld = (h + 3) // 4 * 4
    n_ea, n_tag = L, L - 1
    return 4 * n_nodes * (n_ea * 3 * ld + n_tag * (K + 2) * ld + ld + 4 + 4 + kw["output_dim"])


def fused_tensor_flops(tiles: int, kw) -> float:
    """tcgen05 work of one launch of the tile kernel: 128x128x128 TF32 GEMMs x 3 (split precision) per tile.  The backward
    program multiplies the same number of tiles (dS = G W2, d cur = dHi Wi + dHj Wj, (K+1) segments per TAGConv)."""
    L, K = kw["n_gnn_layers"], kw["K"]
    gemms = 1 + (L - 2) * 3 + 2 + (L - 1) * (K + 1)  # first EA: W2 ; middle EAs: Wi, Wj, W2 ; last EA: Wi, Wj ; TAGs
    return tiles * gemms * 3 * 2.0 * 128 ** 3


def wgrad_tensor_flops(n_nodes: int, kw) -> float:
    """Executed TF32 flops (x3 passes) of the grouped weight gradient: dW2, dWi, dWj per EdgeAggregation, dW_0..dW_K per
    TAGConv, mask_embd -- each 2 * nodes * out * in."""
    h, L, K, nf, out = kw["hidden_dim"], kw["n_gnn_layers"], kw["K"], kw["nfeature_dim"], kw["output_dim"]
    mn = 0
    for li in range(L):
        fin = nf if li == 0 else h
        fout = out if li == L - 1 else h
        mn += fout * h + 2 * h * fin
    mn += (L - 1) * (K + 1) * h * h + 2 * nf * h
    return 3 * 2.0 * n_nodes * mn


def ea_algorithmic_bytes(n_nodes: int, n_edges: int, h: int) -> int:
    """read Hi, read Hj, write S (4*N*h each) + CSR rowptr + src idx + edge_attr + We."""
    return 3 * 4 * n_nodes * h + 4 * (n_nodes + 1) + 4 * n_edges + 8 * n_edges + 4 * 3 * h


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index, self.rows, self.proc, self.thread = gpu_index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu_index), "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.rows.append((time.time(), [c.strip() for c in line.split(",")]))
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 - 0.2 <= t <= t1 + 0.2 and len(r) >= 9] or [r for (_, r) in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}

        def num(v):
            try:
                return float(v)
            except ValueError:
                return None
        sm = [num(r[1]) for r in rows if num(r[1]) is not None]
        reasons = set()
        for r in rows:
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": num(rows[0][2]),
                "power_w_max": max((num(r[3]) or 0.0) for r in rows), "samples": len(rows), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle restatement on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_sample_batch(cfg):
    """(batch, graphs, description) -- the bounded CPU sample of the workload: the whole batch for configs/standard.json;
    for the hidden-512 configurations (tens of TFLOP per step on the host) a slice of it, said in `sample`."""
    from poweflownet_b200.data import synthetic_batch
    if cfg["case"] == "118v2":
        return synthetic_batch("118v2", cfg["batch"], seed=1234), cfg["batch"], f"one case118v2 batch of {cfg['batch']} graphs"
    if cfg["case"] == "6470rte":
        return synthetic_batch("6470rte", 1, seed=1234), 1, "ONE case6470rte graph of the 32-graph batch (the full batch is ~16 TFLOP per step on the host)"
    return (synthetic_batch(cases=["14"] * 10 + ["118v2"] * 5 + ["6470rte"], seed=1234), 16,
            "a 16-graph slice (10 x case14, 5 x case118v2, 1 x case6470rte) of the 64-graph mixed batch")


def time_oracle_cpu(cfg, steps: int, warmup: int, budget_s: float):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import common
    from oracle import pfn_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(1234)
    batch, graphs, what = cpu_sample_batch(cfg)
    model = common.load_seeded(O.MaskEmbdMultiMPN(**cfg["model"])).train()
    times = []
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        model.zero_grad(set_to_none=True)
        t0 = time.perf_counter()
        O.forward_loss_backward(model, batch, "mse")
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_start > budget_s and len(times) >= 3:
            break
    total = sum(times)
    return {"graphs_per_s": graphs * len(times) / total, "ms_per_step": 1e3 * total / len(times), "steps": len(times),
            "cores": cores, "threads": torch.get_num_threads(), "what": what}


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = time_oracle_cpu(cfg, args.steps, args.warmup, budget_s=150.0)
    sample = (f"{r['steps']} steps of fwd+MSE+bwd (oracle restatement of networks/MPN.py + PyG semantics; fp32, torch CPU, "
              f"train mode, dropout 0.2) on {r['what']}")
    line = {"impl": "reference", "metric": cfg["metric"], "value": r["graphs_per_s"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(cfg),
            "cpu_baseline": {"value": r["graphs_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "port", "sample": sample},
            "e2e": {"value": r["graphs_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


def time_oracle_gpu(cfg, dev, batch, steps=10, warmup=3):
    """The "reference GPU path" bar of SURVEY.md section 2.1 / BASELINE.md: the SAME oracle (the reference's op-for-op
    arithmetic: index_select, cat, Linear, scatter_add, autograd) in eager torch CUDA on this B200, same batch, train
    mode, forward + MSE + backward, CUDA events."""
    import torch
    import common
    from oracle import pfn_oracle as O
    torch.manual_seed(1234)
    model = common.load_seeded(O.MaskEmbdMultiMPN(**cfg["model"])).to(dev).train()
    for _ in range(warmup):
        model.zero_grad(set_to_none=True)
        O.forward_loss_backward(model, batch, "mse")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        model.zero_grad(set_to_none=True)
        O.forward_loss_backward(model, batch, "mse")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del model
    return ms


def time_ea_fwd_alone(lib, dev, batch, h, iters=240, n_sets=12):
    """The fused EdgeAggregation message+aggregate forward kernel (pfn_ea_fwd) timed ALONE on `batch`: `iters`
    back-to-back launches bracketed by one pair of CUDA events on the launching stream, rotating over `n_sets` distinct
    (Hi, Hj, S) buffer sets so that no launch finds its operands in L2 (n_sets x operand bytes > 126 MB).  Measured twice:
    launched eagerly from Python and replayed from a CUDA graph holding the same `iters` launches (no host launch cost in
    the timed region); returns (us per launch eager, us per launch graph)."""
    import torch
    from poweflownet_b200 import _lib, ops
    n, ld = batch.num_nodes, (h + 3) // 4 * 4
    g = ops.PreparedGraph(batch.edge_index, batch.edge_attr, n, mode=1)
    gen = torch.Generator(device=dev).manual_seed(7)
    sets = [(torch.randn(n, ld, device=dev, generator=gen), torch.randn(n, ld, device=dev, generator=gen),
             torch.empty(n, ld, device=dev)) for _ in range(n_sets)]
    we = torch.randn(h, 2, device=dev, generator=gen)

    def launch(i, stream):
        hi, hj, s = sets[i % n_sets]
        _lib.check(lib.pfn_ea_fwd(hi.data_ptr(), hj.data_ptr(), ld, g.ws.data_ptr(), n, g.e_raw, we.data_ptr(), 2,
                                  s.data_ptr(), ld, h, stream), "pfn_ea_fwd")
    cur = torch.cuda.current_stream().cuda_stream
    for i in range(n_sets):
        launch(i, cur)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        launch(i, cur)
    e1.record()
    torch.cuda.synchronize()
    us_eager = 1e3 * e0.elapsed_time(e1) / iters
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for i in range(iters):
                launch(i, side.cuda_stream)
    graph.replay()
    torch.cuda.synchronize()
    reps = 3
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    us_graph = 1e3 * e0.elapsed_time(e1) / (reps * iters)
    del sets, graph
    return us_eager, us_graph


def ea_roofline(lib, dev, batch, h, peaks, iters, n_sets, profile_key):
    n_edges = 2 * int(batch.edge_index.size(1))
    ea_bytes = ea_algorithmic_bytes(batch.num_nodes, n_edges, h)
    us_eager, us_graph = time_ea_fwd_alone(lib, dev, batch, h, iters, n_sets)
    us = min(us_eager, us_graph)
    achieved = ea_bytes / (us * 1e-6) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r2_ea_fwd_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            rec = json.load(fh).get(profile_key)
        if rec:
            traffic, traffic_src = rec.get("dram_bytes_per_launch"), rec.get("source")
    return {"kernel": "pfn_ea_fwd (fused EdgeAggregation message+aggregate, forward)", "bound": "hbm", "achieved": achieved,
            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": ea_bytes, "us_per_launch": us, "us_per_launch_eager": us_eager,
            "us_per_launch_cuda_graph": us_graph, "launches_timed": iters, "peak_source": peaks["source"],
            "nodes": batch.num_nodes, "directed_edges": n_edges, "hidden_dim": h}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def time_train_epoch(cfg, dev, epochs=4, samples=4000):
    """Whole optimisation epochs (the reference's utils/training.py:30-80 loop, optimizer included) on a device-resident
    dataset: `training.GraphedEpochs` (batches assembled on the GPU into the static buffers of a captured step, graph
    replay, one-launch AdamW, ONE loss read-back per epoch) for MSE and for the parser-default Masked_L2_loss, and, beside
    it, the eager `training.train_epoch`.  Informational: not the headline metric (which excludes the optimizer)."""
    import torch
    from poweflownet_b200.data import synthetic_raw_case
    from poweflownet_b200.datasets import PowerFlowData
    from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
    from poweflownet_b200.optim import FusedAdamW
    from poweflownet_b200.training import GraphedEpochs, train_epoch
    import common
    case, batch = cfg["case"], cfg["batch"]
    ds = PowerFlowData(case=case, split=[.5, .2, .3], task="train", device=dev, raw=[synthetic_raw_case(case, samples, seed=7)])
    steps = ds.num_batches(batch, drop_last=True)

    def timed_epochs(run):
        first = run()  # warm-up epoch
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = first
        for _ in range(epochs):
            last = run()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), first, last

    out = {"unit": UNIT, "steps_per_epoch": steps, "epochs": epochs, "dataset_samples": len(ds),
           "h2d_bytes_per_epoch": 8 * len(ds), "d2h_bytes_per_epoch": 4,
           "includes": "pfn_batch_assemble from the device-resident dataset (sample ids shuffled on the host), graph prep, "
                       "forward, fused loss, backward, pfn_adamw_step; one loss read-back per epoch"}
    for key, loss in (("mse", "mse"), ("masked_l2", "masked_l2")):
        gen = torch.Generator().manual_seed(0)
        model = common.load_seeded(MaskEmbdMultiMPN(**cfg["model"])).to(dev)
        runner = GraphedEpochs(model, ds, batch, FusedAdamW(model.parameters(), lr=1e-3), loss=loss)
        ms, first, last = timed_epochs(lambda: runner.run_epoch(shuffle=True, generator=gen))
        rec = {"value": epochs * steps * batch / (ms / 1e3), "ms_per_step": ms / (epochs * steps), "loss_first_epoch": first,
               "loss_last_epoch": last, "api": f"poweflownet_b200.training.GraphedEpochs(model, dataset, {batch}, FusedAdamW, loss='{loss}').run_epoch()"}
        if key == "mse":
            out.update(rec)
        else:
            out["masked_l2"] = rec
    gen = torch.Generator().manual_seed(0)
    model = common.load_seeded(MaskEmbdMultiMPN(**cfg["model"])).to(dev)
    opt = FusedAdamW(model.parameters(), lr=1e-3)
    loss_fn = torch.nn.MSELoss()
    ms, first, last = timed_epochs(lambda: train_epoch(model, ds.loader(batch, shuffle=True, generator=gen, drop_last=True), loss_fn, opt, dev))
    out["eager_train_epoch"] = {"value": epochs * steps * batch / (ms / 1e3), "ms_per_step": ms / (epochs * steps),
                                "loss_first_epoch": first, "loss_last_epoch": last,
                                "api": "poweflownet_b200.training.train_epoch(model, dataset.loader(128, shuffle=True), MSELoss(), FusedAdamW, device)"}
    return out


def read_profile(lib):
    from poweflownet_b200 import _lib
    names = ["ea_fwd", "ea_bwd", "hop", "gemm_fwd", "gemm_dgrad", "gemm_wgrad", "prep", "fused_fwd", "fused_bwd"]
    prof = {}
    for cat, name in enumerate(names):
        tot, cnt = C.c_double(), C.c_int64()
        _lib.check(lib.pfn_profile_read(cat, C.byref(tot), C.byref(cnt)), "pfn_profile_read")
        prof[name] = (tot.value, cnt.value)
    return prof


def dp_parity(model, dev, cfg, world, rank, strong):
    """max relative error (max-norm, Frobenius over the whole flat vector) of the N-rank all-reduced gradient against rank 0
    recomputing the GLOBAL batch alone -- dropout off, same weights; SURVEY.md section 4 item 6."""
    import torch
    import torch.distributed as dist
    from poweflownet_b200.training import fused_mse_step
    p_keep = model.dropout.p
    model.dropout.p = 0.0
    try:
        shards = [make_batch(cfg, 77, r, world, strong) for r in range(world)]
        total = sum(s.num_nodes for s in shards) * cfg["model"]["output_dim"]
        fused_mse_step(model, shards[rank].to(dev), total)  # the attached reducer all-reduces the flat gradient
        flat = torch.cat([p.grad.reshape(-1) for p in model._engine_params()]).clone()
        if rank != 0:
            return None
        from poweflownet_b200.data import GraphBatch
        reducer, model._grad_reducer = model._grad_reducer, None
        try:
            off_n, parts = 0, {f: [] for f in ("x", "y", "bus_type", "pred_mask", "edge_index", "edge_attr", "batch", "ptr")}
            for gi, s in enumerate(shards):
                for f in ("x", "y", "bus_type", "pred_mask", "edge_attr"):
                    parts[f].append(getattr(s, f))
                parts["edge_index"].append(s.edge_index + off_n)
                parts["batch"].append(s.batch + sum(x.num_graphs for x in shards[:gi]))
                parts["ptr"].append((s.ptr if gi == 0 else s.ptr[1:]) + off_n)
                off_n += s.num_nodes
            whole = GraphBatch(**{f: torch.cat(v, dim=1 if f == "edge_index" else 0) for f, v in parts.items()})
            fused_mse_step(model, whole.to(dev), total)
            ref = torch.cat([p.grad.reshape(-1) for p in model._engine_params()]).clone()
            # control: rank 0's own shard WITHOUT the reduction -- the reduced gradient must differ from it by O(1),
            # otherwise the comparison above would say nothing about the exchange
            fused_mse_step(model, shards[rank].to(dev), total)
            local = torch.cat([p.grad.reshape(-1) for p in model._engine_params()])
        finally:
            model._grad_reducer = reducer
        d = (flat.double() - ref.double())
        return {"max_rel": float(d.abs().max() / ref.double().abs().max()), "fro_rel": float(d.norm() / ref.double().norm()),
                "global_graphs": sum(s.num_graphs for s in shards), "global_nodes": off_n,
                "unreduced_rank0_vs_global_fro_rel": float((local.double() - ref.double()).norm() / ref.double().norm())}
    finally:
        model.dropout.p = p_keep
        dist.barrier()


def run_ours(args, cfg):
    import torch
    import torch.distributed as dist
    from poweflownet_b200 import _lib, parallel
    from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
    from poweflownet_b200.training import GraphedMSEStep, PipelinedMSESteps, fused_mse_step

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    rank, world, local_rank = parallel.init_from_env("nccl")
    if world != args.gpus:
        raise RuntimeError(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = _lib.lib()
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import common
    kw = cfg["model"]
    strong = cfg["scaling"] == "strong" or args.scaling == "strong"
    peaks = load_peaks()

    torch.manual_seed(1234)
    model = common.load_seeded(MaskEmbdMultiMPN(**kw)).to(dev).train()  # identical weights on every rank
    if world > 1:
        parallel.attach_gradient_allreduce(model)
    n_rot = N_ROTATE if cfg["case"] == "118v2" else 2
    host_batches = [make_batch(cfg, 1234 + i, rank, world, strong).pin_memory() for i in range(n_rot)]
    dev_batches = [b.to(dev) for b in host_batches]
    n_nodes, e_raw = dev_batches[0].num_nodes, int(dev_batches[0].edge_index.size(1))
    graphs_local = dev_batches[0].num_graphs
    counts = torch.tensor([n_nodes * kw["output_dim"], graphs_local], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(counts)
    total_count, graphs_global = int(counts[0]), int(counts[1])  # global element count: the MSE mean is over ALL ranks' nodes
    same_shape = all(b.num_nodes == n_nodes and int(b.edge_index.size(1)) == e_raw for b in dev_batches)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize on both sides; CUDA events; MAX over ranks (ms)."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        t1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        barrier()
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), t0, t1

    step_eager = lambda i: fused_mse_step(model, dev_batches[i % n_rot], total_count)  # noqa: E731
    e2e_eager = lambda i: float(fused_mse_step(model, host_batches[i % n_rot].to(dev, non_blocking=True), total_count).item())  # noqa: E731
    use_graphs = not args.no_graph and same_shape
    graphed = GraphedMSEStep(model, dev_batches[0], total_count) if use_graphs else None
    step_graph = (lambda i: graphed(dev_batches[i % n_rot])) if graphed is not None else None  # noqa: E731
    e2e_graph = (lambda i: float(graphed(host_batches[i % n_rot]).item())) if graphed is not None else None  # noqa: E731
    pipe = PipelinedMSESteps(model, dev_batches[0], total_count) if use_graphs else None

    def e2e_pipe(i):
        # batch i was copied while step i-1 ran; this iteration queues the step of batch i FIRST (so that the GPU does not
        # wait for the host to issue the next copies), then the copy of batch i+1, then reads the loss of batch i
        loss = pipe.step()
        pipe.prefetch(host_batches[(i + 1) % n_rot])
        return float(loss.item())

    def step_pipe(i):
        # resident batches through the same two-buffer pipeline: the device-to-device copy of batch i+1 into the captured
        # step's input buffers runs on the copy stream while batch i computes (GraphedMSEStep alone queues its eight
        # small copies in front of every replay)
        pipe.step()
        pipe.prefetch(dev_batches[(i + 1) % n_rot])
    pending_loss = []

    def e2e_pipe_async(i):
        # as e2e_pipe, but the loss of step i is read after step i + 1 has been queued (4-byte pinned D2H + event per step);
        # informational: shows how much of e2e_pipe is the host waiting on `.item()` rather than GPU work
        pending_loss.append(pipe.step_async())
        pipe.prefetch(host_batches[(i + 1) % n_rot])
        if len(pending_loss) > 1:
            pending_loss.pop(0)()
    for i in range(max(args.warmup, 3)):
        step_eager(i)
        e2e_eager(i)
        if graphed is not None:
            step_graph(i)
            e2e_graph(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)

    # ---- eager steps with the kernel timing hooks on (they bracket launches with events; not capturable) ----
    lib.pfn_profile_enable(1)
    launches0 = lib.pfn_launch_count()
    ms_total, t0, t1 = timed(step_eager, args.steps)
    launches = int(lib.pfn_launch_count() - launches0)
    lib.pfn_profile_enable(0)
    ms_eager, _, _ = timed(step_eager, args.steps)
    prof = read_profile(lib)
    # the headline number: K steps on device-resident batches, launched eagerly (programmatic dependent launches) and,
    # unless --no-graph, also replayed as one captured CUDA graph per step; the faster of the two launch modes is reported
    ms_graph = None
    if graphed is not None:
        ms_graph, _, _ = timed(step_graph, args.steps)
    ms_pipe = None
    if pipe is not None:
        pipe.prefetch(dev_batches[0])
        for i in range(3):
            step_pipe(i)
        ms_pipe = timed(lambda i: step_pipe(i + 3), args.steps)[0]
        pipe.step()  # drain the batch prefetched by the last timed step
    modes = {"eager": ms_eager, "graph": ms_graph, "pipelined": ms_pipe}
    launch_mode = min((k for k, v in modes.items() if v is not None), key=lambda k: modes[k])
    use_graph = launch_mode != "eager"
    ms_total_clean = modes[launch_mode]
    t1b = time.time()
    clocks = sampler.stop(t0, t1b) if rank == 0 else None
    # ---- end to end from pinned host memory (same launch modes) ----
    ms_e2e_eager, _, _ = timed(e2e_eager, args.steps)
    ms_e2e_graph = timed(e2e_graph, args.steps)[0] if graphed is not None else None
    ms_e2e_pipe, ms_e2e_pipe_async = None, None
    if pipe is not None:
        pipe.prefetch(host_batches[0])
        for i in range(3):
            e2e_pipe(i)
        ms_e2e_pipe = timed(lambda i: e2e_pipe(i + 3), args.steps)[0]
        pipe.step()  # drain the batch prefetched by the last timed step
        # informational: the same loop with the loss read one step late (every loss still reaches the host inside the timed
        # region: the last one is read by the final call below, before the closing event)
        pipe.prefetch(host_batches[0])
        for i in range(3):
            e2e_pipe_async(i)

        def async_then_flush(i):
            e2e_pipe_async(i + 3)
            if i == args.steps - 1:
                while pending_loss:
                    pending_loss.pop(0)()
        ms_e2e_pipe_async = timed(async_then_flush, args.steps)[0]
        pipe.step()
    e2e_modes = {"eager": ms_e2e_eager, "graph": ms_e2e_graph, "pipelined": ms_e2e_pipe}
    e2e_mode = min((k for k, v in e2e_modes.items() if v is not None), key=lambda k: e2e_modes[k])
    ms_e2e = e2e_modes[e2e_mode]

    # ---- multi-GPU extras: gradient parity of the data-parallel step, strong-scaling leg ----
    parity, strong_leg = None, None
    if world > 1:
        parity = dp_parity(model, dev, cfg, world, rank, strong)
        if not strong and cfg["case"] is not None:
            sb = [make_batch(cfg, 4321 + i, rank, world, True).to(dev) for i in range(n_rot)]
            cnt = torch.tensor([sb[0].num_nodes * kw["output_dim"]], dtype=torch.int64, device=dev)
            dist.all_reduce(cnt)
            s_total = int(cnt[0])
            s_step = GraphedMSEStep(model, sb[0], s_total) if not args.no_graph else None
            run = (lambda i: s_step(sb[i % n_rot])) if s_step is not None else (lambda i: fused_mse_step(model, sb[i % n_rot], s_total))
            for i in range(5):
                run(i)
            ms_s, _, _ = timed(run, args.steps)
            strong_leg = {"scaling": "strong", "global_batch": cfg["batch"], "graphs_per_rank": sb[0].num_graphs,
                          "ms_per_step": ms_s / args.steps, "value": cfg["batch"] * args.steps / (ms_s / 1e3), "unit": UNIT}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    graphs = graphs_global * args.steps
    value = graphs / (ms_total_clean / 1e3)
    e2e_value = graphs / (ms_e2e / 1e3)
    h = kw["hidden_dim"]
    step_ms_hooks = ms_total / args.steps
    kernel_share = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps,
                        "share_of_step": (v[0] / args.steps) / step_ms_hooks if step_ms_hooks > 0 else None}
                    for k, v in prof.items()}
    fused_route = prof["fused_fwd"][1] > 0
    # ---- roofline of the section-8d kernel on THIS workload's shapes, timed alone in this run ----
    big = n_nodes * ((h + 3) // 4 * 4) * 4 > (32 << 20)
    roof = ea_roofline(lib, dev, dev_batches[0], h, peaks, iters=12 if big else 240, n_sets=2 if big else 12,
                       profile_key="large" if big else "standard")
    ea_ms, ea_cnt = prof["ea_fwd"]
    roof["us_per_launch_inside_step"] = 1e3 * ea_ms / ea_cnt if ea_cnt > 0 else None
    roof["in_timed_step"] = ea_cnt > 0
    roof["how"] = ("pfn_ea_fwd on this workload's shapes timed alone: back-to-back launches between one pair of CUDA events, rotating "
                   "over distinct operand sets larger than L2; the smaller of eager and CUDA-graph-replayed launches; burst HBM peak as "
                   "denominator." + (" At this batch the step itself runs the graph-resident kernels (see roofline_step), where "
                                     "message+aggregate reads Hi/Hj from shared memory and moves no HBM bytes at all." if fused_route else
                                     " The timed step launches this kernel once per EdgeAggregation layer (us_per_launch_inside_step)."))
    # ---- the kernels the timed step consists of ----
    tf32_peak = peaks["bf16_tflops_sustained"] / 2.0
    roofline_step = []
    tiles = (n_nodes + 117) // 118 if cfg["case"] == "118v2" else None

    def step_entry(name, key, flops, hbm_bytes, note):
        ms_k, cnt_k = prof[key]
        if cnt_k <= 0:
            return
        us = 1e3 * ms_k / cnt_k
        per_step = cnt_k / args.steps
        e = {"kernel": name, "us_per_launch": us, "launches_per_step": per_step, "share_of_step": ms_k / args.steps / step_ms_hooks}
        if flops is not None:
            tf = flops / per_step / (us * 1e-6) / 1e12
            e.update({"bound": "tensor", "achieved": tf, "peak": tf32_peak, "unit": "TFLOP/s", "frac": tf / tf32_peak,
                      "flops_note": "executed TF32 flops (3 passes of split precision); peak = bf16_tflops_sustained / 2"})
        if hbm_bytes is not None:
            gb = hbm_bytes / per_step / (us * 1e-6) / 1e9
            e.update({"hbm_algorithmic_gbs": gb, "hbm_frac": gb / peaks["hbm_gbs"]})
        e["note"] = note
        roofline_step.append(e)
    if fused_route and tiles:
        saved = fused_saved_bytes(n_nodes, kw)
        step_entry("k_mpn_fused_fwd<1,0> (whole forward, one tile of whole graphs per CTA)", "fused_fwd", fused_tensor_flops(tiles, kw), saved,
                   "HBM bytes = activations saved for the backward pass")
        step_entry("k_mpn_fused_fwd<1,3> (whole backward data path)", "fused_bwd", fused_tensor_flops(tiles, kw), 2 * saved,
                   "HBM bytes = saved activations re-read + per-layer gradient buffers written (approx. 2x the forward's)")
    else:
        for key, nm in (("gemm_fwd", "k_gemm_tc (forward Linears)"), ("gemm_dgrad", "k_gemm_tc (data gradients)"), ("hop", "k_hop"),
                        ("ea_fwd", "k_ea_fwd_tma"), ("ea_bwd", "k_ea_bwd")):
            step_entry(nm, key, None, None, "layer-wise route")
    step_entry("k_wgrad_group + k_wgrad_group_reduce (all weight gradients)", "gemm_wgrad", wgrad_tensor_flops(n_nodes, kw), None,
               "one grouped launch per step")
    line = {
        "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total_clean / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_dict(cfg),
        "details": {"global_batch": graphs_global, "graphs_rank0": graphs_local, "nodes_rank0": n_nodes, "directed_edges_rank0": 2 * e_raw,
                    "parallelism": (f"dp{world}: graphs sharded per rank" + (" by branch count" if strong else "") +
                                    ", one NCCL all-reduce of the flat fp32 gradient buffer per step") if world > 1 else "single GPU",
                    "l2": f"steps rotate over {n_rot} resident batches; per-step activation+scratch working set > 126 MB L2 (no explicit flush)",
                    "timed_region": "graph prep + forward + fused MSE + backward (+ all-reduce); optimizer.step excluded (SURVEY 8 f4; see train_epoch)",
                    "launch": {"graph": "CUDA graph replay of the captured step (training.GraphedMSEStep)",
                               "pipelined": "CUDA graph replays of the captured step over two input-buffer sets, the device-to-device copy of "
                                            "the next resident batch into the idle set on a side stream (training.PipelinedMSESteps)",
                               "eager": "eager: stream-ordered launches with programmatic dependent launch (training.fused_mse_step)"}[launch_mode],
                    "route": "graph-resident kernels (pfn_mpn_forward_tiled / pfn_mpn_backward_tiled)" if fused_route else "layer-wise kernels"},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": host_batches[0].nbytes(), "d2h_bytes_per_step": 4,
                "api": {"pipelined": "poweflownet_b200.training.PipelinedMSESteps: step() of batch i (CUDA-graph replay), prefetch(pinned "
                                     "batch i+1) on a copy stream, loss.item() of batch i -- one H2D, one step, one D2H per iteration",
                        "graph": "poweflownet_b200.training.GraphedMSEStep(model, batch)(pinned_host_batch) + loss.item()",
                        "eager": "poweflownet_b200.training.fused_mse_step(model, pinned_host_batch.to(device, non_blocking=True)) + loss.item()"}[e2e_mode]},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roof,
        "roofline_step": roofline_step,
        "kernel_time": kernel_share,
        "ms_per_step_with_timing_hooks": step_ms_hooks,
        "ms_per_step_eager": ms_eager / args.steps,
        "ms_per_step_cuda_graph": None if ms_graph is None else ms_graph / args.steps,
        "ms_per_step_pipelined": None if ms_pipe is None else ms_pipe / args.steps,
        "e2e_ms_per_step_eager": ms_e2e_eager / args.steps,
        "e2e_ms_per_step_cuda_graph": None if ms_e2e_graph is None else ms_e2e_graph / args.steps,
        "e2e_ms_per_step_pipelined": None if ms_e2e_pipe is None else ms_e2e_pipe / args.steps,
        # NOT the e2e figure: the pipelined loop with each loss read one step late (PipelinedMSESteps.step_async)
        "e2e_ms_per_step_pipelined_loss_read_one_step_late": None if ms_e2e_pipe_async is None else ms_e2e_pipe_async / args.steps,
    }
    if parity is not None:
        line["dp_parity"] = parity
    if strong_leg is not None:
        line["strong_scaling"] = strong_leg
    extras = world == 1 and not args.no_extras
    if extras:
        r = time_oracle_cpu(cfg, steps=10, warmup=2, budget_s=25.0)
        line["cpu_baseline"] = {"value": r["graphs_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "port",
                                "ms_per_step": r["ms_per_step"],
                                "sample": f"{r['steps']} steps of fwd+MSE+bwd of the oracle (torch CPU fp32, train mode) on {r['what']}, "
                                          f"{r['cores']} host cores"}
        try:
            ms_ref = time_oracle_gpu(cfg, dev, dev_batches[0])
            line["reference_gpu"] = {"value": graphs_local / (ms_ref / 1e3), "unit": UNIT, "ms_per_step": ms_ref,
                                     "what": "the oracle (reference arithmetic, op for op) in eager torch CUDA fp32 on this GPU: same batch, "
                                             "train mode, forward + MSE + backward, 10 steps between CUDA events",
                                     "speedup_of_value": value / (graphs_local / (ms_ref / 1e3))}
        except Exception as exc:  # an auxiliary leg must never cost the headline line
            line["reference_gpu"] = {"error": f"{type(exc).__name__}: {exc}"}
        if args.config == "standard":
            try:
                line["train_epoch"] = time_train_epoch(cfg, dev)
            except Exception as exc:
                line["train_epoch"] = {"error": f"{type(exc).__name__}: {exc}"}
            try:
                line["configs3"] = side_config_leg(CONFIGS["large"], dev, lib, peaks)
            except Exception as exc:
                line["configs3"] = {"error": f"{type(exc).__name__}: {exc}"}
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


def side_config_leg(cfg, dev, lib, peaks, steps=4):
    """BASELINE configs[3] (case6470rte x 32, configs/large.json) inside the default run, so that the driver's record
    carries it: step time on one GPU (layer-wise route, eager launches, CUDA events) and the section-8d kernel's roofline
    fraction at this size (1.28 GB algorithmic per launch)."""
    import torch
    import common
    from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
    from poweflownet_b200.training import fused_mse_step
    torch.manual_seed(1234)
    model = common.load_seeded(MaskEmbdMultiMPN(**cfg["model"])).to(dev).train()
    batches = [make_batch(cfg, 1234 + i).to(dev) for i in range(2)]
    for i in range(2):
        fused_mse_step(model, batches[i])
    torch.cuda.synchronize()
    lib.pfn_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fused_mse_step(model, batches[i % 2])
    e1.record()
    torch.cuda.synchronize()
    lib.pfn_profile_enable(0)
    ms = e0.elapsed_time(e1) / steps
    prof = read_profile(lib)
    out = {"config": config_dict(cfg), "metric": cfg["metric"], "value": cfg["batch"] / (ms / 1e3), "unit": UNIT, "ms_per_step": ms,
           "steps": steps, "nodes": batches[0].num_nodes,
           "kernel_time_ms_per_step": {k: v[0] / steps for k, v in prof.items() if v[1] > 0}}
    del model
    torch.cuda.empty_cache()
    out["roofline"] = ea_roofline(lib, dev, batches[0], cfg["model"]["hidden_dim"], peaks, iters=12, n_sets=2, profile_key="large")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--config", choices=list(CONFIGS), default="standard", help="BASELINE.json workload (default configs[1])")
    ap.add_argument("--scaling", choices=["weak", "strong"], default=None, help="standard/large: shard ONE batch over the ranks instead of one batch per rank")
    ap.add_argument("--no-extras", action="store_true", help="skip the auxiliary legs (cpu_baseline, reference_gpu, train_epoch, configs3)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.steps is None:
        args.steps = 200 if args.config == "standard" else 10
    _claim_stdout()
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
