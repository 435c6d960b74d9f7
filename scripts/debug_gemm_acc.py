import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from poweflownet_b200 import ops
dev = "cuda:0"
def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max()), float((a - b).norm() / b.norm())
for (m, k, n) in [(4096, 128, 128), (4096, 32, 128), (4096, 8, 128), (4096, 512, 128), (15104, 128, 144)]:
    g = torch.Generator().manual_seed(k)
    x, w = torch.randn(m, k, generator=g), torch.randn(n, k, generator=g) / k ** 0.5
    ref = x.double() @ w.double().T
    out = ops.new_rows(m, n, dev)
    ops.linear_fwd(ops.new_rows(m, k, dev).copy_(x) if k % 4 == 0 else None, w.to(dev), k, k, n, None, out)
    torch.cuda.synchronize()
    e_tc = rel(out[:, :n].cpu(), ref)
    e_f32 = rel((x @ w.T), ref)
    # what pure single-pass TF32 would give (round both operands to tf32)
    def tf32(t):
        return (t.view(torch.int32) + 0x1000 & ~0x1FFF).view(torch.float32)
    e_1x = rel(tf32(x).double() @ tf32(w).double().T, ref)
    print(f"M={m} K={k} N={n}: tc {e_tc[0]:.2e} {e_tc[1]:.2e} | torch fp32 cpu {e_f32[0]:.2e} {e_f32[1]:.2e} | 1xTF32 model {e_1x[0]:.2e} {e_1x[1]:.2e}")
