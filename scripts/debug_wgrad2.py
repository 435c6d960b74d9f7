import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from poweflownet_b200 import ops
torch.set_printoptions(linewidth=220, precision=1, sci_mode=False)
dev = "cuda:0"
m, n_in, n_out = 32, 40, 40
# dY[m, o] = 1 if o == m (nodes 0..31 select output rows 0..31);  X[m, i] = 1000*m + i  =>  dW[o, i] = 1000*o + i for o < 32
dy = torch.zeros(m, n_out); dy[torch.arange(m), torch.arange(m)] = 1.0
x = (1000 * torch.arange(m)[:, None] + torch.arange(n_in)[None, :]).float()
dyd, xd = ops.new_rows(m, n_out, dev), ops.new_rows(m, n_in, dev)
dyd[:, :n_out] = dy.to(dev); xd[:, :n_in] = x.to(dev)
dw = torch.full((n_out, n_in), -7.0, device=dev)
ops.linear_wgrad(dyd, xd, n_in, n_out, dw, n_in)
torch.cuda.synchronize()
ref = dy.double().T @ x.double()
print("MN", os.environ.get("PFN_WG_MN"), "DEBUG", os.environ.get("PFN_WG_DEBUG"), os.environ.get("PFN_WG_LBO"), os.environ.get("PFN_WG_SBO"), os.environ.get("PFN_WG_KSTEP"), "max err", (dw.cpu().double() - ref).abs().max().item(), "nonzero", int((dw != 0).sum()))
print(dw.cpu()[:12, :12])
