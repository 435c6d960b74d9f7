"""Losses and optimizer next to the hot path on the B200:
`poweflownet_b200.losses.{Masked_L2_loss, PowerImbalance, MixedMSEPoweImbalance}` against the fixtures produced by the
reference's own utils/custom_loss_functions.py (tests/golden/loss_*.pt) and against the CPU oracle at full size;
`poweflownet_b200.optim.FusedAdamW` against `torch.optim.AdamW`.  Tolerance 1e-5 relative fp32; where the fp32
reference itself is further than that from its fp64 twin (the branch injections cancel), the kernel has to stay within
4x the reference's own rounding error."""
import copy

import pytest
import torch

import common
import make_golden_losses as mgl
from oracle import pfn_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5
DEV = "cuda:0"


def _losses():
    from poweflownet_b200 import losses
    return losses


def _stats_dev():
    return tuple(s.to(DEV) for s in mgl.loss_stats())


def _check(got, ref32, ref64, what):
    got, ref64 = got.detach().cpu().double(), ref64.double()
    scale = float(ref64.abs().max())
    err = float((got - ref64).abs().max()) / scale
    err_ref = float((ref32.double() - ref64).abs().max()) / scale
    assert err <= max(TOL, 4 * err_ref), (what, err, err_ref)
    e32 = common.rel_err(got.float(), ref32)
    assert max(e32) <= max(TOL, 8 * err_ref), (what, e32, err_ref)


@pytest.mark.parametrize("name", mgl.LOSS_CASES)
def test_power_imbalance_matches_reference(name):
    gold = torch.load(mgl.loss_golden_path(name), weights_only=False)
    batch, pred = mgl.loss_predictions(name)
    x = pred.to(DEV).requires_grad_(True)
    fn = _losses().PowerImbalance(*mgl.loss_stats())  # statistics arrive on the host, as train.py:97 passes them
    loss = fn(x, batch.edge_index.to(DEV), batch.edge_attr.to(DEV))
    assert loss.shape == () and loss.dtype == torch.float32
    loss.backward()
    _check(loss, gold["pi_loss"], gold["pi_loss_fp64"], "loss")
    _check(x.grad, gold["pi_grad"], gold["pi_grad_fp64"], "d loss / d x")


@pytest.mark.parametrize("name", ["case14_small", "mixed", "isolated_and_parallel"])
def test_mixed_loss_matches_reference(name):
    gold = torch.load(mgl.loss_golden_path(name), weights_only=False)
    batch, pred = mgl.loss_predictions(name)
    x = pred.to(DEV).requires_grad_(True)
    fn = _losses().MixedMSEPoweImbalance(*_stats_dev(), alpha=0.9)
    loss = fn(x, batch.edge_index.to(DEV), batch.edge_attr.to(DEV), batch.y.to(DEV))
    (3.0 * loss).backward()  # a non-unit incoming gradient must scale the stored one
    _check(loss, gold["mixed_loss"], gold["mixed_loss_fp64"], "mixed loss")
    _check(x.grad / 3.0, gold["mixed_grad"], gold["mixed_grad_fp64"], "d mixed / d x")


def test_power_imbalance_full_size_against_the_oracle():
    """BASELINE config 2 size (case118v2 x 128): CPU oracle in fp32 and fp64 as the checker."""
    batch = common.synthetic_batch(case="118v2", batch_size=128, seed=1234)
    g = torch.Generator().manual_seed(7)
    pred = batch.y + 0.1 * torch.randn(batch.y.shape, generator=g)
    ref = {}
    for tag, dt in (("32", torch.float32), ("64", torch.float64)):
        x = pred.to(dt).clone().requires_grad_(True)
        loss = O.power_imbalance(x, batch.edge_index, batch.edge_attr.to(dt), *mgl.loss_stats(dt))
        loss.backward()
        ref[tag] = (loss.detach(), x.grad)
    x = pred.to(DEV).requires_grad_(True)
    fn = _losses().PowerImbalance(*mgl.loss_stats())
    loss = fn(x, batch.edge_index.to(DEV), batch.edge_attr.to(DEV))
    loss.backward()
    _check(loss, ref["32"][0], ref["64"][0], "loss")
    _check(x.grad, ref["32"][1], ref["64"][1], "d loss / d x")
    # deterministic: no atomics anywhere on the path
    x2 = pred.to(DEV).requires_grad_(True)
    loss2 = fn(x2, batch.edge_index.to(DEV), batch.edge_attr.to(DEV))
    loss2.backward()
    assert torch.equal(loss2, loss) and torch.equal(x2.grad, x.grad)


def test_power_imbalance_reuses_a_prepared_graph_and_skips_the_gradient_without_grad():
    from poweflownet_b200 import ops
    from poweflownet_b200._lib import lib
    batch, pred = mgl.loss_predictions("case14_small")
    ei, ea, x = batch.edge_index.to(DEV), batch.edge_attr.to(DEV), pred.to(DEV)
    fn = _losses().PowerImbalance(*mgl.loss_stats())
    graph = ops.PreparedGraph(ei, ea, x.size(0), mode=1)
    n0 = lib().pfn_launch_count()
    with torch.no_grad():
        val = fn(x, ei, ea, graph)
    assert lib().pfn_launch_count() - n0 == 2  # k_pi_node + k_pi_final: no graph prep, no gradient kernel
    gold = torch.load(mgl.loss_golden_path("case14_small"), weights_only=False)
    _check(val, gold["pi_loss"], gold["pi_loss_fp64"], "loss (eval)")
    assert not val.requires_grad


def test_power_imbalance_errors_like_the_reference():
    fn = _losses().PowerImbalance(*mgl.loss_stats())
    x = torch.zeros((4, 4), device=DEV, requires_grad=True)
    with pytest.raises(IndexError):  # custom_loss_functions.py:133 indexes edge_index[0, 0] unguarded
        fn(x, torch.zeros((2, 0), dtype=torch.long, device=DEV), torch.zeros((0, 2), device=DEV))
    with pytest.raises(RuntimeError):  # no CPU path
        fn(torch.zeros((4, 4)), torch.zeros((2, 1), dtype=torch.long), torch.zeros((1, 2)))


def test_masked_l2_module_matches_oracle():
    batch = common.make_batch("case118_h33")
    g = torch.Generator().manual_seed(3)
    pred = batch.y + 0.3 * torch.randn(batch.y.shape, generator=g)
    for regularize, coeff in ((True, 1), (True, 0.25), (False, 1)):
        xr = pred.clone().requires_grad_(True)
        want = O.masked_l2_loss(xr, batch.y, batch.pred_mask, regularize, coeff)
        want.backward()
        x = pred.to(DEV).requires_grad_(True)
        got = _losses().Masked_L2_loss(regularize, coeff)(x, batch.y.to(DEV), batch.pred_mask.to(DEV))
        got.backward()
        assert max(common.rel_err(got.detach().cpu(), want.detach())) < TOL
        assert max(common.rel_err(x.grad.cpu(), xr.grad)) < TOL


def test_training_dispatch_with_power_imbalance_reaches_the_parameters():
    """utils/training.py:63-68: `masked_out = out*pred_mask + x*(1-pred_mask)`, `loss_fn(masked_out, edge_index,
    edge_attr)`, `loss.backward()` -- parameter gradients against the oracle model + oracle loss on the CPU."""
    from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
    name = "case14_small"
    kw = common.model_kwargs(name)
    batch = common.make_batch(name)
    ref_model = common.load_seeded(O.MaskEmbdMultiMPN(**kw)).eval()
    out = ref_model(batch)
    masked = out * batch.pred_mask + batch.x * (1 - batch.pred_mask)
    want = O.power_imbalance(masked, batch.edge_index, batch.edge_attr, *mgl.loss_stats())
    want.backward()
    model = common.load_seeded(MaskEmbdMultiMPN(**kw)).to(DEV).eval()
    data = batch.to(DEV)
    out = model(data)
    masked = out * data.pred_mask + data.x * (1 - data.pred_mask)
    got = _losses().PowerImbalance(*mgl.loss_stats())(masked, data.edge_index, data.edge_attr)
    got.backward()
    assert max(common.rel_err(got.detach().cpu(), want.detach())) < 2e-5
    for (k, p), (_, q) in zip(model.named_parameters(), ref_model.named_parameters()):
        assert max(common.rel_err(p.grad.cpu(), q.grad)) < 5e-5, k


# ---- AdamW --------------------------------------------------------------------------------------------
def _param_set(seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = [(129, 10), (129,), (129, 129), (4, 129), (4,), (1,), (129, 260), (300, 257)]
    return [torch.randn(s, generator=g) for s in shapes]


def test_fused_adamw_matches_torch_adamw():
    from poweflownet_b200.optim import FusedAdamW
    init = _param_set()
    ref = [torch.nn.Parameter(p.clone()) for p in init]
    mine = [torch.nn.Parameter(p.clone().to(DEV)) for p in init]
    kw = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    opt_ref = torch.optim.AdamW(ref, foreach=False, **kw)
    opt = FusedAdamW(mine, **kw)
    sched_ref = torch.optim.lr_scheduler.OneCycleLR(opt_ref, max_lr=1e-2, total_steps=12)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-2, total_steps=12)
    g = torch.Generator().manual_seed(1)
    from poweflownet_b200._lib import lib
    for step in range(8):
        grads = [0.1 * torch.randn(p.shape, generator=g) for p in init]
        for p, q, gr in zip(ref, mine, grads):
            p.grad, q.grad = gr.clone(), gr.to(DEV)
        if step == 3:  # a parameter without a gradient is skipped and keeps its step count
            ref[1].grad = mine[1].grad = None
        n0 = lib().pfn_launch_count()
        opt_ref.step()
        opt.step()
        assert lib().pfn_launch_count() - n0 == (1 if step < 4 else 2)  # the straggler gets its own bias corrections
        sched_ref.step()
        sched.step()
        for k, (p, q) in enumerate(zip(ref, mine)):
            assert max(common.rel_err(q.detach().cpu(), p.detach())) < 1e-6, (step, k)
    for p, q in zip(ref, mine):
        for key in ("exp_avg", "exp_avg_sq"):
            assert max(common.rel_err(opt.state[q][key].cpu(), opt_ref.state[p][key])) < 1e-6
        assert float(opt.state[q]["step"]) == float(opt_ref.state[p]["step"])


def test_fused_adamw_state_dict_interchanges_with_torch():
    from poweflownet_b200.optim import FusedAdamW
    init = _param_set(5)[:4]
    a = [torch.nn.Parameter(p.clone().to(DEV)) for p in init]
    b = [torch.nn.Parameter(p.clone().to(DEV)) for p in init]
    opt_a, opt_b = FusedAdamW(a, lr=2e-3), torch.optim.AdamW(b, lr=2e-3, foreach=False)
    g = torch.Generator().manual_seed(2)
    grads = [[0.1 * torch.randn(p.shape, generator=g).to(DEV) for p in init] for _ in range(4)]
    for gs in grads[:2]:
        for p, q, gr in zip(a, b, gs):
            p.grad, q.grad = gr.clone(), gr.clone()
        opt_a.step()
        opt_b.step()
    # swap the optimizer states and continue: both pairs must stay together
    sd_a, sd_b = copy.deepcopy(opt_a.state_dict()), copy.deepcopy(opt_b.state_dict())
    opt_a.load_state_dict(sd_b)
    opt_b.load_state_dict(sd_a)
    for gs in grads[2:]:
        for p, q, gr in zip(a, b, gs):
            p.grad, q.grad = gr.clone(), gr.clone()
        opt_a.step()
        opt_b.step()
    for p, q in zip(a, b):
        assert max(common.rel_err(p.detach().cpu(), q.detach().cpu())) < 1e-6


def test_fused_adamw_rejects_cpu_parameters():
    from poweflownet_b200.optim import FusedAdamW
    p = torch.nn.Parameter(torch.zeros(3))
    p.grad = torch.ones(3)
    with pytest.raises(RuntimeError):
        FusedAdamW([p]).step()


def test_fused_adamw_trains_the_model_like_torch_adamw():
    """Three optimisation steps of the standard model on one batch: same losses with either optimizer."""
    from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
    from poweflownet_b200.optim import FusedAdamW
    from poweflownet_b200.training import fused_mse_step
    kw = common.model_kwargs("case118_h33")
    kw["dropout_rate"] = 0.0
    data = common.make_batch("case118_h33").to(DEV)
    losses = []
    for cls in (FusedAdamW, torch.optim.AdamW):
        model = common.load_seeded(MaskEmbdMultiMPN(**kw)).to(DEV).train()
        opt = cls(model.parameters(), lr=1e-3)
        run = []
        for _ in range(3):
            opt.zero_grad()
            run.append(float(fused_mse_step(model, data).item()))
            opt.step()
        losses.append(run)
    assert losses[0][0] == losses[1][0]
    for a, b in zip(*losses):
        assert abs(a - b) <= 1e-5 * abs(b), losses
    assert losses[0][2] < losses[0][0]
