"""Multi-GPU check + timing of the one-shot NVLink all-reduce (pfn_allreduce_peer) against NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/test_oneshot_allreduce.py

Checks: exact equality with a rank-ordered reference sum (all_gather + sequential fp32 adds), equality across ranks,
300 back-to-back calls with changing data, replay inside a CUDA graph.  Timing: CUDA events around 200 calls, max over ranks."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from poweflownet_b200 import parallel  # noqa: E402

rank, world, local = parallel.init_from_env("nccl")
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
n = int(os.environ.get("AR_N", "354500"))
mode = os.environ.get("AR_MODE", "auto")  # auto | one | two
red = parallel.OneShotAllReduce(n, dev, two_shot=None if mode == "auto" else mode == "two")
n4 = red.n
gen = torch.Generator(device=dev).manual_seed(100 + rank)
ok = True
for it in range(300):
    x = torch.randn(n4, device=dev, generator=gen)
    gathered = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(gathered, x)
    want = gathered[0].clone()
    for r in range(1, world):
        want += gathered[r]
    got = red(x.clone())
    if not torch.equal(got, want):
        ok = False
        print(f"rank {rank} iteration {it}: mismatch, max abs diff {float((got - want).abs().max())}", flush=True)
        break
# graph replay: static buffer, new contents per replay
static = torch.zeros(n4, device=dev)
red(static)  # warm
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    with torch.cuda.graph(g, stream=side):
        red(static)
for it in range(50):
    x = torch.randn(n4, device=dev, generator=gen)
    gathered = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(gathered, x)
    want = gathered[0].clone()
    for r in range(1, world):
        want += gathered[r]
    static.copy_(x)
    g.replay()
    if not torch.equal(static, want):
        ok = False
        print(f"rank {rank} graph replay {it}: mismatch", flush=True)
        break
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)


def timed(fn, iters=200):
    for _ in range(20):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms) * 1e3


buf = torch.randn(n4, device=dev)
us_one = timed(lambda: red(buf))
us_graph = timed(lambda: g.replay())
us_nccl = timed(lambda: dist.all_reduce(buf))
stamps = None
if os.environ.get("PFN_AR_TIMING") == "1":
    import ctypes as C
    from poweflownet_b200 import _lib
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 8)()
    if _lib.lib().pfn_allreduce_debug_stamps(buf) == 0:
        t = list(buf)
        stamps = {"launch_to_wait_done": t[1] - t[0], "push": t[2] - t[1], "exchange1": t[3] - t[2], "sum_and_push2": t[4] - t[3] if t[4] else None,
                  "exchange2": t[5] - t[4] if t[5] else None, "tail": t[6] - (t[5] if t[5] else t[3]), "total_ns": t[6] - t[0]}
if rank == 0:
    print(json.dumps({"phase_ns_last_call_cta0": stamps, "world": world, "two_shot": red.two_shot, "floats": n4, "bytes": 4 * n4, "exact": bool(flag.item()), "us_one_shot": us_one,
                      "us_one_shot_graph_replay": us_graph, "us_nccl_all_reduce": us_nccl}), flush=True)
dist.barrier()
dist.destroy_process_group()
