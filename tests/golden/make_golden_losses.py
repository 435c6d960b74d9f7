"""Generate the loss fixtures by running the REFERENCE's own loss code.

    python tests/golden/make_golden_losses.py     # needs /root/reference (this container only)

`/root/reference/utils/custom_loss_functions.py` is imported UNMODIFIED over the PyG stand-in in
`oracle/pyg_shim/` (its `MessagePassing` implements `flow='target_to_source'` and `update(aggregated, x)`).
For each case the reference `PowerImbalance` (:99-286) and `MixedMSEPoweImbalance(alpha=0.9)` (:289-306, the
alpha train.py:101 uses) are evaluated in fp32 and fp64 on seeded predictions; losses and the gradients w.r.t.
the predictions go to `tests/golden/loss_<case>.pt`.  The fixtures travel; the GPU box never needs the reference.
"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import common  # noqa: E402

REFERENCE = os.environ.get("PFN_REFERENCE", "/root/reference")
SHIM = os.path.join(common.ROOT, "oracle", "pyg_shim")

LOSS_CASES = ("tiny", "case14_small", "case118_h33", "mixed", "already_undirected", "first_edge_reversed_only",
              "isolated_and_parallel")


def loss_stats(dtype=torch.float32):
    """(xymean, xystd, edgemean, edgestd) in the shapes `PowerFlowData.get_data_means_stds` returns
    (datasets/PowerFlowData.py:115-117): per-unit voltage, angle in degrees, P / Q; branch r / x."""
    return (torch.tensor([[1.02, -7.5, 0.25, 0.08]], dtype=dtype), torch.tensor([[0.03, 6.0, 0.6, 0.3]], dtype=dtype),
            torch.tensor([[0.035, 0.11]], dtype=dtype), torch.tensor([[0.02, 0.05]], dtype=dtype))


def loss_predictions(name):
    """Seeded stand-in for the model output: the z-scored targets plus noise."""
    batch = common.make_batch(name)
    g = torch.Generator().manual_seed(4242)
    return batch, batch.y + 0.1 * torch.randn(batch.y.shape, generator=g)


def loss_golden_path(name):
    return os.path.join(common.GOLDEN_DIR, f"loss_{name}.pt")


def import_reference_losses():
    if not os.path.isdir(REFERENCE):
        raise SystemExit(f"{REFERENCE} not found: fixtures can only be regenerated where the reference is mounted")
    for p in (SHIM, REFERENCE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import utils.custom_loss_functions as ref_losses  # the reference's file, byte for byte
    assert os.path.realpath(ref_losses.__file__).startswith(os.path.realpath(REFERENCE))
    return ref_losses


def main():
    ref = import_reference_losses()
    torch.set_num_threads(1)
    for name in LOSS_CASES:
        batch, pred = loss_predictions(name)
        out = {}
        for tag, dt in (("", torch.float32), ("_fp64", torch.float64)):
            stats = loss_stats(dt)
            x = pred.to(dt).clone().requires_grad_(True)
            ea, y = batch.edge_attr.to(dt), batch.y.to(dt)
            pi = ref.PowerImbalance(*stats)
            loss = pi(x, batch.edge_index, ea)
            loss.backward()
            out["pi_loss" + tag], out["pi_grad" + tag] = loss.detach().clone(), x.grad.clone()
            x2 = pred.to(dt).clone().requires_grad_(True)
            mixed = ref.MixedMSEPoweImbalance(*stats, alpha=0.9)
            loss2 = mixed(x2, batch.edge_index, ea, y)
            loss2.backward()
            out["mixed_loss" + tag], out["mixed_grad" + tag] = loss2.detach().clone(), x2.grad.clone()
        out["meta"] = {"source": "reference utils/custom_loss_functions.py over oracle/pyg_shim", "alpha": 0.9,
                       "torch": str(torch.__version__)}
        torch.save(out, loss_golden_path(name))
        print(f"{name:28s} N={batch.num_nodes:5d} pi={float(out['pi_loss']):.6e} (fp64 {float(out['pi_loss_fp64']):.6e}) "
              f"mixed={float(out['mixed_loss']):.6e} -> {os.path.getsize(loss_golden_path(name))/1024:.1f} KiB")


if __name__ == "__main__":
    main()
