"""Accuracy of the tensor-core Linear (3xTF32 on tcgen05) against fp64, by reduction length: rel. error (max-norm,
Frobenius) of pfn_linear_fwd, of torch's fp32 CPU matmul and of a single-pass TF32 model.  PFN_TC_DRAIN=0/1 selects the
plain / register-flushed accumulation (see k_gemm_tc); default = the library's own choice."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from poweflownet_b200 import ops
dev = "cuda:0"
def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max()), float((a - b).norm() / b.norm())
print("env", {k: v for k, v in os.environ.items() if k.startswith("PFN_")})
for (m, k, n) in [(4096, 128, 128), (4096, 32, 128), (4096, 512, 512), (4096, 2048, 256), (15104, 132, 129), (12940, 512, 512)]:
    g = torch.Generator().manual_seed(k)
    x, w = torch.randn(m, k, generator=g), torch.randn(n, k, generator=g) / k ** 0.5
    # positive-mean operands: every partial sum has the same sign, the worst case for a truncating accumulator
    xp, wp = x.abs(), w.abs()
    for tag, (xx, ww) in (("randn", (x, w)), ("abs", (xp, wp)), ("neg", (-xp, wp)), ("relu", (x.clamp_min(0), w))):
        ref = xx.double() @ ww.double().T
        out = ops.new_rows(m, n, dev)
        ops.linear_fwd(ops.new_rows(m, k, dev).copy_(xx), ww.to(dev), k, k, n, None, out)
        torch.cuda.synchronize()
        e_tc = rel(out[:, :n].cpu(), ref)
        e_f32 = rel((xx @ ww.T), ref)
        o = out[:, :n].cpu().double()
        big = ref.abs() > 0.1 * ref.abs().mean()  # (relative errors of near-cancelled sums would dominate the mean)
        shrink = float((torch.sign(ref) * (o - ref) / ref.abs().clamp_min(1e-30))[big].mean())  # < 0: magnitudes come out too small
        shrink32 = float((torch.sign(ref) * ((xx @ ww.T).double() - ref) / ref.abs().clamp_min(1e-30))[big].mean())
        print(f"M={m} K={k} N={n} {tag:5s}: tc {e_tc[0]:.2e} {e_tc[1]:.2e} | torch fp32 cpu {e_f32[0]:.2e} {e_f32[1]:.2e} | mean magnitude error tc {shrink:+.2e} (fp32 cpu {shrink32:+.2e})")
