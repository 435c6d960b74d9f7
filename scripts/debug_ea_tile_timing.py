"""Phase stamps (%globaltimer, ns) of k_ea_fwd_tile at case118v2 x 128, hidden 129: PFN_EA_TILE_TIMING=1 python scripts/debug_ea_tile_timing.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PFN_EA_TILE_TIMING"] = "1"
import torch  # noqa: E402

from poweflownet_b200 import _lib, ops  # noqa: E402
from poweflownet_b200.data import synthetic_batch  # noqa: E402

dev = torch.device("cuda", 0)
lib = _lib.lib()
batch = synthetic_batch("118v2", 128).to(dev)
n, h, ld = batch.num_nodes, 129, 132
g = ops.PreparedGraph(batch.edge_index, batch.edge_attr, n, mode=1, tile_rows=118)
sets = [(torch.randn(n, ld, device=dev), torch.randn(n, ld, device=dev), torch.empty(n, ld, device=dev)) for _ in range(12)]
we = torch.randn(h, 2, device=dev)
raw = lib
names = ["start", "after pdl_wait", "copies + loads issued", "row pointers staged", "nbr/ea staged", "Hj tile landed", "first row stored", "end"]
for rep in range(3):
    for i in range(24):
        hi, hj, s = sets[i % 12]
        _lib.check(lib.pfn_ea_fwd_tiled(hi.data_ptr(), hj.data_ptr(), ld, g.ws.data_ptr(), n, g.e_raw, we.data_ptr(), 2, s.data_ptr(), ld, h, 118,
                                        torch.cuda.current_stream().cuda_stream), "pfn_ea_fwd_tiled")
    torch.cuda.synchronize()
    host = (C.c_ulonglong * 32)()
    fn = raw.pfn_debug_ea_tile_stamps
    fn.argtypes = [C.POINTER(C.c_ulonglong)]
    fn(host)
    for cta, base in ((0, 0), (100, 16)):
        t0 = host[base]
        print(f"rep {rep} CTA {cta}: " + "  ".join(f"{names[k]} +{host[base + k] - t0}" for k in range(8)))
    print(f"rep {rep} CTA 100 started {int(host[16]) - int(host[0])} ns after CTA 0")
