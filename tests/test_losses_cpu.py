"""CPU checks for the loss / optimizer neighbours of the hot path:
the oracle restatements against fixtures produced by the reference's own `utils/custom_loss_functions.py`
(tests/golden/make_golden_losses.py), the hand-derived gradient the CUDA kernel implements against autograd,
and the AdamW restatement against torch's optimizer."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import make_golden_losses as mgl  # noqa: E402
from emulation import power_imbalance_emulation  # noqa: E402
from oracle import pfn_oracle as oracle  # noqa: E402


def _gold(name):
    return torch.load(mgl.loss_golden_path(name), weights_only=False)


@pytest.mark.parametrize("name", mgl.LOSS_CASES)
def test_oracle_power_imbalance_matches_reference(name):
    torch.set_num_threads(1)
    gold = _gold(name)
    batch, pred = mgl.loss_predictions(name)
    for tag, dt, tol in (("", torch.float32, 0.0), ("_fp64", torch.float64, 0.0)):
        stats = mgl.loss_stats(dt)
        x = pred.to(dt).clone().requires_grad_(True)
        loss = oracle.power_imbalance(x, batch.edge_index, batch.edge_attr.to(dt), *stats)
        loss.backward()
        assert torch.equal(loss.detach(), gold["pi_loss" + tag]), (float(loss), float(gold["pi_loss" + tag]))
        assert torch.equal(x.grad, gold["pi_grad" + tag])
        x2 = pred.to(dt).clone().requires_grad_(True)
        mixed = oracle.mixed_mse_power_imbalance(x2, batch.edge_index, batch.edge_attr.to(dt), batch.y.to(dt), stats, alpha=0.9)
        mixed.backward()
        assert torch.equal(mixed.detach(), gold["mixed_loss" + tag])
        assert torch.equal(x2.grad, gold["mixed_grad" + tag])


@pytest.mark.parametrize("name", mgl.LOSS_CASES)
def test_hand_derived_gradient_matches_autograd(name):
    """The formulas of k_pi_grad (tests/emulation.py restates them) against autograd of the oracle, in fp64."""
    batch, pred = mgl.loss_predictions(name)
    stats = mgl.loss_stats(torch.float64)
    x = pred.double().clone().requires_grad_(True)
    loss = oracle.power_imbalance(x, batch.edge_index, batch.edge_attr.double(), *stats)
    loss.backward()
    und = oracle.undirect_graph(batch.edge_index, batch.edge_attr.double())
    loss_e, dx_e = power_imbalance_emulation(pred.double(), und, stats)
    assert abs(float(loss_e) - float(loss.detach())) <= 1e-12 * abs(float(loss.detach()))
    assert float((dx_e - x.grad).abs().max()) <= 1e-11 * float(x.grad.abs().max())
    gold = _gold(name)
    assert float((dx_e - gold["pi_grad_fp64"]).abs().max()) <= 1e-11 * float(gold["pi_grad_fp64"].abs().max())


def test_fp32_emulation_within_contract():
    """Error budget: the fp32 evaluation of the same formulas against the fp64 reference values."""
    for name in mgl.LOSS_CASES:
        batch, pred = mgl.loss_predictions(name)
        und = oracle.undirect_graph(batch.edge_index, batch.edge_attr)
        loss_e, dx_e = power_imbalance_emulation(pred, und, mgl.loss_stats())
        gold = _gold(name)
        ref = gold["pi_grad_fp64"]
        err = float((dx_e.double() - ref).abs().max() / ref.abs().max())
        err_ref = float((gold["pi_grad"].double() - ref).abs().max() / ref.abs().max())
        assert err <= max(1e-5, 4 * err_ref), (name, err, err_ref)
        assert abs(float(loss_e) - float(gold["pi_loss_fp64"])) <= 1e-5 * abs(float(gold["pi_loss_fp64"]))


def test_oracle_adamw_matches_torch():
    torch.manual_seed(0)
    shapes = [(129, 10), (129,), (4, 129), (1,), (129, 129)]
    params = [torch.randn(s) for s in shapes]
    ref_params = [torch.nn.Parameter(p.clone()) for p in params]
    opt = torch.optim.AdamW(ref_params, lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, foreach=False)
    ms = [torch.zeros_like(p) for p in params]
    vs = [torch.zeros_like(p) for p in params]
    for step in range(1, 6):
        grads = [torch.randn(s) * 0.1 for s in shapes]
        for p, g in zip(ref_params, grads):
            p.grad = g.clone()
        opt.step()
        oracle.adamw_step(params, grads, ms, vs, step, lr=3e-3)
        for p, q in zip(params, ref_params):
            assert torch.equal(p, q.detach())
