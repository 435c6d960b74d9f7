"""How far is the fp32 reference from ITSELF?  The oracle (op-for-op restatement of the reference, torch CPU fp32) is run
twice on the same inputs and weights with different intra-op thread counts (different summation orders inside mm /
scatter_add), and once in fp64.  Printed per parameter: rel. error (max-norm, Frobenius) of run A vs run B and of each
vs the fp64 twin.  This is the floor under any "within 1e-5 of the fp32 reference" statement at that size.

    python scripts/oracle_self_noise.py 6470rte 2 512 5 3 [threads_a threads_b]
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import common  # noqa: E402
from oracle import pfn_oracle as O  # noqa: E402
from poweflownet_b200.data import synthetic_batch  # noqa: E402


def run(kw, batch, threads, double=False):
    torch.set_num_threads(threads)
    model = common.load_seeded(O.MaskEmbdMultiMPN(**kw))
    if double:
        model = model.double()
        batch = common.GraphBatch(**{f: (getattr(batch, f).double() if getattr(batch, f).is_floating_point() else getattr(batch, f))
                                     for f in ("x", "y", "bus_type", "pred_mask", "edge_index", "edge_attr", "batch", "ptr")})
    model.train()
    t0 = time.time()
    loss, out = O.forward_loss_backward(model, batch, "mse")
    return float(loss), out.detach(), {k: p.grad.detach().clone() for k, p in model.named_parameters()}, time.time() - t0


def main():
    case, b, h, L, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    ta, tb = (int(sys.argv[6]), int(sys.argv[7])) if len(sys.argv) > 7 else (1, os.cpu_count() or 8)
    kw = dict(common.MODEL_DIMS, hidden_dim=h, n_gnn_layers=L, K=K, dropout_rate=0.0)
    names = case.split(",")
    batch = synthetic_batch(cases=[n for n in names for _ in range(b)]) if len(names) > 1 else synthetic_batch(case, b)
    la, oa, ga, sa = run(kw, batch, ta)
    lb, ob, gb, sb = run(kw, batch, tb)
    ld, od, gd, sd = run(kw, batch, tb, double=True)
    rows, worst = {}, {"a_vs_b": 0.0, "a_vs_fp64": 0.0, "b_vs_fp64": 0.0}
    for k in ga:
        r = {"a_vs_b": common.rel_err(ga[k], gb[k]), "a_vs_fp64": common.rel_err(ga[k].double(), gd[k]),
             "b_vs_fp64": common.rel_err(gb[k].double(), gd[k])}
        rows[k] = r
        for n in worst:
            worst[n] = max(worst[n], *r[n])
    res = {"case": case, "graphs": b * len(names), "nodes": batch.num_nodes, "hidden_dim": h, "n_gnn_layers": L, "K": K,
           "threads": [ta, tb], "seconds": [sa, sb, sd], "loss": [la, lb, ld],
           "out_a_vs_b": common.rel_err(oa, ob), "out_a_vs_fp64": common.rel_err(oa.double(), od),
           "worst_grad_rel_err": worst,
           "per_tensor_over_1e-5": {k: v for k, v in rows.items() if max(max(x) for x in v.values()) > 1e-5}}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
