"""The algebra the CUDA engine implements (tests/emulation.py) against the reference golden vectors:
restructured forward (per-node GEMMs, hoisted second Linear, K-segmented TAGConv) and the hand-derived
backward.  CPU only; this is the no-GPU guard on csrc/engine.cu's derivations."""
import pytest
import torch

import common
from emulation import EngineEmulation
from oracle import pfn_oracle as O

ALL = list(common.CASES)
TOL = 1e-5  # BASELINE.json north_star: 1e-5 relative fp32


def _setup(name):
    gold = torch.load(common.golden_path(name), weights_only=True)
    kw = gold["meta"]["model_kwargs"]
    model = common.load_seeded(O.MaskEmbdMultiMPN(**kw))
    batch = common.GraphBatch(**gold["inputs"])
    und = O.undirect_graph(batch.edge_index, batch.edge_attr)
    return gold, kw, model, batch, und


@pytest.mark.parametrize("name", ALL)
def test_restructured_forward_eval(name):
    gold, kw, model, batch, und = _setup(name)
    emu = EngineEmulation(model.state_dict(), kw)
    out = emu.forward(batch, und, training=False)
    assert max(common.rel_err(out, gold["eval_out"])) < TOL
    emu64 = EngineEmulation(model.state_dict(), kw, dtype=torch.float64)
    out64 = emu64.forward(batch, und, training=False)
    assert max(common.rel_err(out64, gold["eval_out_fp64"])) < 1e-12


@pytest.mark.parametrize("name", ALL)
def test_hand_derived_backward(name):
    gold, kw, model, batch, und = _setup(name)
    masks = common.dropout_masks(name, batch.num_nodes)
    for dtype, tol in ((torch.float32, TOL), (torch.float64, 1e-11)):
        emu = EngineEmulation(model.state_dict(), kw, dtype=dtype)
        out = emu.forward(batch, und, training=True, masks=masks)
        dout = 2.0 * (out - batch.y.to(dtype)) / out.numel()
        grads = emu.backward(dout)
        if dtype == torch.float32:
            assert max(common.rel_err(out, gold["train_out"])) < tol
            ref = gold.get("grads")
        else:
            m64 = common.load_seeded(O.MaskEmbdMultiMPN(**kw)).double().train()
            b64 = common.GraphBatch(**{k: (v.double() if v.is_floating_point() else v) for k, v in gold["inputs"].items()})
            O.forward_loss_backward(m64, b64, "mse", dropout_masks=masks)
            ref = {k: p.grad for k, p in m64.named_parameters()}
        if ref is None:
            for k, g in grads.items():
                nrm = float(gold["grad_norm"][k])
                assert abs(float(g.double().norm()) - nrm) <= tol * nrm + 1e-12, k
            continue
        assert set(ref) == set(grads)
        for k in ref:
            assert max(common.rel_err(grads[k], ref[k])) < tol, (k, common.rel_err(grads[k], ref[k]))
