"""Generate the golden fixtures by running the REFERENCE's own model code.

    python tests/golden/make_golden.py            # needs /root/reference (this container only)

`/root/reference/networks/MPN.py` is imported UNMODIFIED; its only missing dependency,
`torch_geometric` (absent from the image), is satisfied by the minimal stand-in under
`oracle/pyg_shim/`.  For each case in `common.CASES` the reference `MaskEmbdMultiMPN` is built,
loaded with seeded weights and run (a) in eval mode and (b) in train mode with MSE loss + backward,
where `model.dropout` -- a plain attribute of the reference module (networks/MPN.py:496) -- is
swapped for a test double that applies pre-drawn keep-masks so other implementations can replay
the identical masks.  Inputs, outputs, loss and parameter gradients go to `tests/golden/<case>.pt`
(`summary_only` cases store gradient norms + strided samples instead of full gradients).
The fixtures travel with the repo; the GPU box never needs /root/reference.
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


def import_reference_mpn():
    if not os.path.isdir(REFERENCE):
        raise SystemExit(f"{REFERENCE} not found: fixtures can only be regenerated where the reference is mounted")
    for p in (SHIM, REFERENCE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import networks.MPN as ref_mpn  # the reference's file, byte for byte
    assert os.path.realpath(ref_mpn.__file__).startswith(os.path.realpath(REFERENCE))
    return ref_mpn


class ReplayDropout(torch.nn.Module):
    """Stands in for `nn.Dropout(p)`: multiplies by the next pre-drawn keep-mask and rescales by 1/(1-p)."""

    def __init__(self, masks, p):
        super().__init__()
        self.masks, self.p, self.calls = masks, p, 0

    def forward(self, x):
        m = self.masks[self.calls]
        self.calls += 1
        return x * m / (1.0 - self.p)


class _MaskFp64:
    """fp64 twin only: the reference casts the mask with `.float()` (networks/MPN.py:533); hand it an
    object whose `.float()` yields fp64 so the double-precision model type-checks."""

    def __init__(self, t):
        self.t = t

    def float(self):
        return self.t.double()


def strided_sample(t, k=257):
    flat = t.reshape(-1)
    step = max(1, flat.numel() // k)
    return flat[::step][:k].clone()


def main():
    ref = import_reference_mpn()
    torch.set_num_threads(1)  # deterministic scatter/reduction order
    for name, spec in common.CASES.items():
        kw = common.model_kwargs(name)
        batch = common.make_batch(name)
        model = common.load_seeded(ref.MaskEmbdMultiMPN(**kw))
        out = {"inputs": {f: getattr(batch, f) for f in ("x", "y", "bus_type", "pred_mask", "edge_index",
                                                          "edge_attr", "batch", "ptr")}}
        # integer work of networks/MPN.py:498-523
        out["is_directed"] = torch.tensor(bool(model.is_directed(batch.edge_index)))
        ei_u, ea_u = model.undirect_graph(batch.edge_index, batch.edge_attr)
        out["undirected_edge_index"], out["undirected_edge_attr"] = ei_u.clone(), ea_u.clone()
        # (a) eval forward, fp32 and the fp64 twin
        model.eval()
        with torch.no_grad():
            out["eval_out"] = model(batch).clone()
        model64 = common.load_seeded(ref.MaskEmbdMultiMPN(**kw)).double().eval()
        b64 = common.GraphBatch(**{k: (v.double() if v.is_floating_point() else v) for k, v in out["inputs"].items()})
        b64.pred_mask = _MaskFp64(batch.pred_mask)
        with torch.no_grad():
            out["eval_out_fp64"] = model64(b64).clone()
        # (b) train forward + MSE + backward with replayed dropout masks
        masks = common.dropout_masks(name, batch.num_nodes)
        model.train()
        model.dropout = ReplayDropout(masks, kw["dropout_rate"])
        model.zero_grad()
        y = model(batch)
        loss = torch.nn.functional.mse_loss(y, batch.y)
        loss.backward()
        assert model.dropout.calls == len(masks), (model.dropout.calls, len(masks))
        out["train_out"], out["train_loss"] = y.detach().clone(), loss.detach().clone()
        grads = {k: p.grad.clone() for k, p in model.named_parameters()}
        if spec.get("summary_only"):
            out["grad_norm"] = {k: g.double().norm().float() for k, g in grads.items()}
            out["grad_absmax"] = {k: g.abs().max() for k, g in grads.items()}
            out["grad_sample"] = {k: strided_sample(g) for k, g in grads.items()}
        else:
            out["grads"] = grads
        out["meta"] = {"model_kwargs": kw, "weights_seed": 4321, "mask_seed": 99,
                       "source": "reference networks/MPN.py over oracle/pyg_shim", "torch": str(torch.__version__)}
        torch.save(out, common.golden_path(name))
        size = os.path.getsize(common.golden_path(name))
        print(f"{name:28s} N={batch.num_nodes:5d} E_raw={batch.edge_index.size(1):5d} directed={bool(out['is_directed'])} "
              f"loss={float(loss):.6f} -> {size/1024:.1f} KiB")


if __name__ == "__main__":
    main()
