import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from poweflownet_b200 import ops
torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)
dev = "cuda:0"
def run(m, n_in, n_out, mode):
    g = torch.Generator().manual_seed(0)
    if mode == "ones":
        dy, x = torch.ones(m, n_out), torch.ones(m, n_in)
    elif mode == "pattern":
        dy = torch.zeros(m, n_out); dy[torch.arange(m), torch.arange(m) % n_out] = 1.0   # row m selects column m % n_out
        x = torch.arange(n_in).float()[None, :].repeat(m, 1) + 100 * (torch.arange(m) % n_out)[:, None].float()
    else:
        dy, x = torch.randn(m, n_out, generator=g), torch.randn(m, n_in, generator=g)
    dyd, xd = ops.new_rows(m, n_out, dev), ops.new_rows(m, n_in, dev)
    dyd[:, :n_out] = dy.to(dev); xd[:, :n_in] = x.to(dev)
    dw = torch.full((n_out, n_in), -7.0, device=dev); db = torch.full((n_out,), -7.0, device=dev)
    ops.linear_wgrad(dyd, xd, n_in, n_out, dw, n_in, dbias=db)
    torch.cuda.synchronize()
    ref = dy.double().T @ x.double()
    err = (dw.cpu().double() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-9)
    print(f"m={m} n_in={n_in} n_out={n_out} mode={mode}: rel err {err:.3e}; db err {(db.cpu().double()-dy.double().sum(0)).abs().max().item():.3e}")
    if err > 1e-4 and n_in <= 16:
        print("got\n", dw.cpu()[:8, :16]); print("want\n", ref.float()[:8, :16])
    return err
for m in (32, 64, 100):
    for mode in ("ones", "pattern", "rand"):
        run(m, 8, 8, mode)
run(1000, 129, 129, "rand")
run(1000, 40, 40, "rand")
run(1000, 4, 129, "rand")
run(1000, 129, 4, "rand")
