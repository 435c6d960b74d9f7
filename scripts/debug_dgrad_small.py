"""Stand-alone Linear forward / data gradient at the small shapes of tests/test_gpu_model.py::test_standalone_layers..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from poweflownet_b200 import ops
dev = "cuda:0"
print("env", {k: v for k, v in os.environ.items() if k.startswith("PFN_")})
g = torch.Generator().manual_seed(0)


def rows(t):
    r = ops.new_rows(t.size(0), t.size(1), dev)
    r.zero_()
    r[:, :t.size(1)].copy_(t)
    return r[:, :t.size(1)] if False else r


for n in (146, 15104):
    for (fin, h) in ((4, 33), (33, 33), (129, 129), (4, 129), (33, 64)):
        ldw = 2 * fin + 2
        dy = torch.randn(n, h, generator=g)
        w1 = torch.randn(h, ldw, generator=g)
        for off in (0, fin):
            ref = dy.double() @ w1[:, off:off + fin].double()
            dx = torch.full((n, fin), float("nan"), device=dev)
            ops.linear_dgrad(rows(dy), w1.to(dev), ldw, fin, h, dx, w_offset=off)
            torch.cuda.synchronize()
            e = float((dx.cpu().double() - ref).abs().max() / ref.abs().max())
            print(f"n={n} dgrad K={h} N={fin} off={off}: {e:.2e}")
        x = torch.randn(n, fin, generator=g)
        out = ops.new_rows(n, h, dev)
        ops.linear_fwd(rows(x), w1.to(dev), ldw, fin, h, None, out)
        ref = x.double() @ w1[:, :fin].double().T
        print(f"n={n} fwd K={fin} N={h}: {float((out[:, :h].cpu().double() - ref).abs().max() / ref.abs().max()):.2e}")
