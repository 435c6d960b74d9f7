"""Whole-model parity on the B200: `poweflownet_b200.networks.MPN.MaskEmbdMultiMPN` (C ABI ->
sm_100a kernels) against (a) the golden vectors produced by the reference's own networks/MPN.py and
(b) the CPU oracle at BASELINE.json's full sizes.  Tolerance 1e-5 relative fp32 (north_star)."""
import pytest
import torch

import common
from oracle import pfn_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5
DEV = "cuda:0"
ALL = list(common.CASES)


def _model(kw, state=None):
    from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
    m = MaskEmbdMultiMPN(**kw)
    if state is None:
        common.load_seeded(m)
    else:
        m.load_state_dict(state)
    return m.to(DEV)


def _close(got, want, what, tol=TOL):
    e = common.rel_err(got.detach().cpu(), want)
    assert max(e) < tol, (what, e)


@pytest.mark.parametrize("name", ALL)
def test_state_dict_keys_and_shapes_match_reference(name):
    gold = torch.load(common.golden_path(name), weights_only=True)
    m = _model(gold["meta"]["model_kwargs"])
    ref = gold["grads"] if "grads" in gold else gold["grad_norm"]
    assert list(m.state_dict().keys()) == list(ref.keys())
    if "grads" in gold:
        for k, v in m.state_dict().items():
            assert tuple(v.shape) == tuple(gold["grads"][k].shape), k


@pytest.mark.parametrize("name", ALL)
def test_eval_forward_matches_reference(name):
    gold = torch.load(common.golden_path(name), weights_only=True)
    m = _model(gold["meta"]["model_kwargs"]).eval()
    batch = common.GraphBatch(**gold["inputs"]).to(DEV)
    with torch.no_grad():
        out = m(batch)
    assert out.shape == gold["eval_out"].shape and out.dtype == torch.float32
    _close(out, gold["eval_out"], "eval_out vs fp32 reference")
    _close(out, gold["eval_out_fp64"].float(), "eval_out vs fp64 twin")


@pytest.mark.parametrize("name", ALL)
def test_train_step_matches_reference(name):
    gold = torch.load(common.golden_path(name), weights_only=True)
    m = _model(gold["meta"]["model_kwargs"]).train()
    batch = common.GraphBatch(**gold["inputs"]).to(DEV)
    m._inject_dropout_masks = common.dropout_masks(name, batch.num_nodes)
    out = m(batch)
    loss = torch.nn.functional.mse_loss(out, batch.y)
    loss.backward()
    _close(out, gold["train_out"], "train_out")
    assert abs(float(loss) - float(gold["train_loss"])) < TOL * abs(float(gold["train_loss"]))
    for k, p in m.named_parameters():
        assert p.grad is not None, k
        if "grads" in gold:
            _close(p.grad, gold["grads"][k], k)
        else:
            nrm = float(gold["grad_norm"][k])
            assert abs(float(p.grad.double().norm()) - nrm) < TOL * nrm + 1e-12, k
            flat = p.grad.reshape(-1).cpu()
            step = max(1, flat.numel() // 257)
            s = flat[::step][:257]
            assert float((s - gold["grad_sample"][k]).abs().max()) < TOL * float(gold["grad_absmax"][k]) + 1e-12, k


def test_fused_mse_step_equals_autograd_step():
    from poweflownet_b200.training import fused_mse_step
    name = "case118_h33"
    gold = torch.load(common.golden_path(name), weights_only=True)
    m = _model(gold["meta"]["model_kwargs"]).train()
    batch = common.GraphBatch(**gold["inputs"]).to(DEV)
    m._inject_dropout_masks = common.dropout_masks(name, batch.num_nodes)
    loss = fused_mse_step(m, batch)
    assert abs(float(loss) - float(gold["train_loss"])) < TOL * abs(float(gold["train_loss"]))
    for k, p in m.named_parameters():
        _close(p.grad, gold["grads"][k], k)


# Gradient parity at BASELINE.json's real shapes.  Three effects put a flat "1e-5 against the fp32 reference" out of reach
# there; the criterion below names each (measured values: profiles/r2_parity.md, scripts/debug_parity.py):
#  (1) The fp32 reference is itself further than 1e-5 from the exact (fp64) gradient once sums run over 10^4..10^5 nodes:
#      1.1e-5..1.8e-5 (Frobenius), up to 7.9e-5 (max-norm) on configs/large.json x case6470rte.  Allowance: SLACK x the
#      reference's own distance to its fp64 twin (measured worst ratio 1.3 on layer gradients, 2.15 on mask_embd).
#  (2) d ReLU is a step function.  A pre-activation p = W1 [x_i | x_j | e] + b1 whose magnitude is at rounding level
#      (|p| < ~1e-6 of its scale) gets its mask decided by the order of the fp32 additions, which differs between ANY two
#      implementations (the reference sums one 2f+2-long dot product per edge, this library adds Hi[i] + Hj[j] + We e).
#      One flipped unit moves every gradient upstream of it by a discrete amount -- measured 2.2e-5 (Frobenius) / 6.2e-5
#      (max-norm) on configs/wide.json x 16 case118v2 graphs, as one jump at layers.4.edge_aggr while everything downstream
#      agrees to 8e-7; the same jumps are what (1) consists of.  With ~1e-7 of all units at risk, batches of >= 10^6
#      edge-channel units nearly always contain one.  Allowance: FLIP_TOL, granted only if the fp64 twin really holds units
#      with |p| <= FLIP_MARGIN x rms(p) (counted with forward hooks on the twin); otherwise the strict bound applies.
#  (3) Depth: rounding errors of ~1e-7 per GEMM stage add up over the 2 x (2 n_gnn_layers - 1) stages of a forward + backward
#      (measured 8e-7 at 11 layer entries once the tensor core's truncating accumulate is flushed and de-biased, see
#      k_gemm_tc<true>; 1-2e-5 before).  Allowance: DEPTH_TOL per layer entry beyond the first five.
SLACK = 2.5
FLIP_TOL = 1e-4
FLIP_MARGIN = 2e-6
DEPTH_TOL = 1e-6


def _units_at_risk(twin, batch64):
    """Number of EdgeAggregation pre-activations of the fp64 twin with |p| <= FLIP_MARGIN x rms(p) (per layer): the oracle's
    forward is re-run with a counting stand-in for `torch.relu` inside its `edge_aggregation`."""
    counts = []
    real = O.torch.relu

    class _Counting:
        def __getattr__(self, name):
            return getattr(torch, name)

        @staticmethod
        def relu(t):
            if t.dim() == 2 and t.size(0) == batch64.edge_index.size(1) * (2 if O.is_directed(batch64.edge_index) else 1):
                counts.append(int((t.abs() <= FLIP_MARGIN * t.pow(2).mean().sqrt()).sum()))
            return real(t)
    saved = O.torch
    O.torch = _Counting()
    try:
        with torch.no_grad():
            twin(batch64)
    finally:
        O.torch = saved
    assert len(counts) == sum(1 for layer in twin.layers if hasattr(layer, "edge_aggr")), counts
    return sum(counts)


def _assert_grad_parity(model, oracle, twin, n_gnn_layers, at_risk):
    tol = max(TOL, TOL + DEPTH_TOL * (2 * n_gnn_layers - 1 - 5))
    flip = FLIP_TOL if at_risk > 0 else 0.0
    worst = {}
    for (k, p), (_, q), (_, r) in zip(model.named_parameters(), oracle.named_parameters(), twin.named_parameters()):
        exact = r.grad
        ours_vs_exact = max(common.rel_err(p.grad.cpu().double(), exact))
        ref_vs_exact = max(common.rel_err(q.grad.double(), exact))
        ours_vs_ref = max(common.rel_err(p.grad.cpu(), q.grad))
        assert ours_vs_exact <= max(tol, SLACK * ref_vs_exact, flip), (k, "vs fp64 twin", ours_vs_exact, ref_vs_exact, at_risk)
        assert ours_vs_ref <= max(tol, (SLACK + 1) * ref_vs_exact, flip), (k, "vs fp32 oracle", ours_vs_ref, ref_vs_exact, at_risk)
        worst[k] = (ours_vs_exact, ref_vs_exact, ours_vs_ref)
    return worst


FULL_SIZE_CASES = {
    # BASELINE configs[1]: case118v2 x 128, configs/standard.json (graph-resident kernels)
    "standard_118x128": (dict(case="118v2", batch_size=128), dict(hidden_dim=129, n_gnn_layers=4, K=3)),
    "6470x2_h64": (dict(case="6470rte", batch_size=2), dict(hidden_dim=64, n_gnn_layers=3, K=3)),
    # BASELINE configs[3]: configs/large.json at its real width AND depth on 6470-bus graphs (layer-wise kernels)
    "large_6470x2": (dict(case="6470rte", batch_size=2), dict(hidden_dim=512, n_gnn_layers=5, K=3)),
    # BASELINE configs[4]: configs/extra_large.json (hidden 512, TEN GNN layers = 19 layer entries) on a small batch that
    # mixes the three grid sizes (variable-N batching, datasets/PowerFlowData.py:67-70,151-155)
    "extra_large_mixed": (dict(cases=["14"] * 4 + ["118v2"] * 2 + ["6470rte"]), dict(hidden_dim=512, n_gnn_layers=10, K=3)),
    # the only configuration the reference's own runs.sh uses (runs.sh:4-12): configs/wide.json, K = 6, L = 6, case6470rte
    "wide_6470x2": (dict(case="6470rte", batch_size=2), dict(hidden_dim=129, n_gnn_layers=6, K=6)),
    # wide.json on small graphs: K = 6 exceeds the graph-resident kernel's four TAGConv segments -> layer-wise route
    "wide_118x16": (dict(case="118v2", batch_size=16), dict(hidden_dim=129, n_gnn_layers=6, K=6)),
}


def test_deep_wide_config_without_units_at_risk_meets_the_strict_bound():
    """configs/wide.json (K = 6, L = 6: 11 layer entries, the deepest K-segmented GEMMs) on the first one-graph batch whose
    fp64 pre-activations all stay clear of zero (no ReLU mask can flip): every gradient within the strict depth-scaled
    bound of the EXACT gradient -- the arithmetic itself, separated from the step function's discontinuity."""
    from poweflownet_b200.data import synthetic_batch
    kw = dict(common.MODEL_DIMS, dropout_rate=0.0, hidden_dim=129, n_gnn_layers=6, K=6)
    for seed in range(40, 60):
        batch = synthetic_batch("118v2", 1, seed=seed)  # ~3e5 edge-channel units: about every second seed has none at risk
        b64 = common.GraphBatch(**{f: (getattr(batch, f).double() if getattr(batch, f).is_floating_point() else getattr(batch, f))
                                   for f in ("x", "y", "bus_type", "pred_mask", "edge_index", "edge_attr", "batch", "ptr")})
        twin = common.load_seeded(O.MaskEmbdMultiMPN(**kw)).double().train()
        if _units_at_risk(twin, b64) == 0:
            break
    else:
        pytest.fail("no batch without units at risk among 20 seeds")
    oracle = common.load_seeded(O.MaskEmbdMultiMPN(**kw)).train()
    O.forward_loss_backward(oracle, batch, "mse")
    O.forward_loss_backward(twin, b64, "mse")
    m = _model(kw, oracle.state_dict()).train()
    dbatch = batch.to(DEV)
    torch.nn.functional.mse_loss(m(dbatch), dbatch.y).backward()
    _assert_grad_parity(m, oracle, twin, kw["n_gnn_layers"], 0)


@pytest.mark.parametrize("name", list(FULL_SIZE_CASES))
def test_full_size_against_oracle(name):
    """Forward + MSE + backward at the real shapes of BASELINE.json's configurations vs the CPU oracle and its fp64 twin,
    dropout off (p=0 in train mode exercises the train path deterministically)."""
    from poweflownet_b200.data import synthetic_batch
    spec, cfg = FULL_SIZE_CASES[name]
    kw = dict(common.MODEL_DIMS, dropout_rate=0.0, **cfg)
    batch = synthetic_batch(**spec)
    oracle = common.load_seeded(O.MaskEmbdMultiMPN(**kw)).train()
    loss_ref, out_ref = O.forward_loss_backward(oracle, batch, "mse")
    twin = common.load_seeded(O.MaskEmbdMultiMPN(**kw)).double().train()
    b64 = common.GraphBatch(**{f: (getattr(batch, f).double() if getattr(batch, f).is_floating_point() else getattr(batch, f))
                               for f in ("x", "y", "bus_type", "pred_mask", "edge_index", "edge_attr", "batch", "ptr")})
    _, out64 = O.forward_loss_backward(twin, b64, "mse")
    m = _model(kw, oracle.state_dict()).train()
    dbatch = batch.to(DEV)
    out = m(dbatch)
    loss = torch.nn.functional.mse_loss(out, dbatch.y)
    loss.backward()
    ref_out_err = max(common.rel_err(out_ref.double(), out64))
    assert max(common.rel_err(out.detach().cpu().double(), out64)) <= max(TOL, SLACK * ref_out_err)
    assert max(common.rel_err(out.detach().cpu(), out_ref)) <= max(TOL, (SLACK + 1) * ref_out_err)
    assert abs(float(loss) - float(loss_ref)) < TOL * float(loss_ref)
    _assert_grad_parity(m, oracle, twin, kw["n_gnn_layers"], _units_at_risk(twin, b64))


def test_masked_l2_loss_through_autograd():
    """Parser-default loss (utils/custom_loss_functions.py:10-46) composed in torch on top of the CUDA model."""
    name = "mixed"
    gold = torch.load(common.golden_path(name), weights_only=True)
    kw = gold["meta"]["model_kwargs"]
    batch = common.GraphBatch(**gold["inputs"])
    masks = common.dropout_masks(name, batch.num_nodes)
    oracle = common.load_seeded(O.MaskEmbdMultiMPN(**kw)).train()
    O.forward_loss_backward(oracle, batch, "masked_l2", dropout_masks=masks)
    m = _model(kw).train()
    m._inject_dropout_masks = masks
    db = batch.to(DEV)
    O.masked_l2_loss(m(db), db.y, db.pred_mask).backward()
    for (k, p), (_, q) in zip(m.named_parameters(), oracle.named_parameters()):
        _close(p.grad, q.grad, k)


def test_standalone_layers_match_oracle_layers():
    from poweflownet_b200.networks.MPN import EdgeAggregation, TAGConv
    # the layers below draw their initial weights from torch's global generator: pin it, so that the case does not depend
    # on how many random numbers earlier tests consumed (an unlucky draw puts a ReLU pre-activation at rounding level,
    # where the mask -- hence dx -- legitimately differs between two summation orders; see FULL_SIZE_CASES)
    torch.manual_seed(20240229)
    batch = common.make_batch("mixed")
    ei, ea = O.undirect_graph(batch.edge_index, batch.edge_attr)
    n = batch.num_nodes
    g = torch.Generator().manual_seed(0)
    for fin, h, fout in ((4, 33, 33), (33, 33, 4), (129, 129, 129)):
        x = torch.randn(n, fin, generator=g)
        ref = O.EdgeAggregation(fin, 2, h, fout)
        mine = EdgeAggregation(fin, 2, h, fout)
        mine.load_state_dict(ref.state_dict())
        mine = mine.to(DEV)
        xr, xm = x.clone().requires_grad_(True), x.clone().to(DEV).requires_grad_(True)
        yr, ym = ref(xr, ei, ea), mine(xm, ei.to(DEV), ea.to(DEV))
        _close(ym, yr.detach(), f"EA {fin}->{fout}")
        w = torch.randn(yr.shape, generator=g)
        (yr * w).sum().backward()
        (ym * w.to(DEV)).sum().backward()
        _close(xm.grad, xr.grad, "EA dx")
        for (k, p), (_, q) in zip(mine.named_parameters(), ref.named_parameters()):
            _close(p.grad, q.grad, f"EA {k}")
    for fin, fout, K in ((33, 33, 3), (16, 7, 2), (129, 129, 3)):
        x = torch.randn(n, fin, generator=g)
        ref = O.TAGConv(fin, fout, K)
        with torch.no_grad():
            ref.bias.uniform_(-0.5, 0.5)
        mine = TAGConv(fin, fout, K)
        mine.load_state_dict(ref.state_dict())
        mine = mine.to(DEV)
        xr, xm = x.clone().requires_grad_(True), x.clone().to(DEV).requires_grad_(True)
        yr, ym = ref(xr, ei), mine(xm, ei.to(DEV))
        _close(ym, yr.detach(), f"TAG {fin}->{fout}")
        w = torch.randn(yr.shape, generator=g)
        (yr * w).sum().backward()
        (ym * w.to(DEV)).sum().backward()
        _close(xm.grad, xr.grad, "TAG dx")
        for (k, p), (_, q) in zip(mine.named_parameters(), ref.named_parameters()):
            _close(p.grad, q.grad, f"TAG {k}")


def test_train_mode_dropout_is_seeded_and_eval_is_deterministic():
    from poweflownet_b200.data import synthetic_batch
    kw = dict(common.MODEL_DIMS, hidden_dim=129, n_gnn_layers=4, K=3, dropout_rate=0.2)
    m = _model(kw)
    batch = synthetic_batch("118v2", 8).to(DEV)
    m.train()
    torch.manual_seed(5)
    a = m(batch).detach().clone()
    torch.manual_seed(5)
    b = m(batch).detach().clone()
    c = m(batch).detach().clone()
    assert torch.equal(a, b) and not torch.equal(a, c)
    m.eval()
    with torch.no_grad():
        e1, e2 = m(batch).clone(), m(batch).clone()
    assert torch.equal(e1, e2) and not torch.equal(e1, a)
    # expectation of the dropout output stays near the eval output (inverted dropout): loose statistical check
    m.train()
    with torch.no_grad():
        mean = torch.stack([m(batch) for _ in range(64)]).mean(0)
    assert float((mean - e1).norm() / e1.norm()) < 0.5


def test_backward_twice_is_rejected_and_input_grad_flows():
    batch = common.make_batch("tiny").to(DEV)
    m = _model(common.model_kwargs("tiny")).eval()
    x = batch.x.clone().requires_grad_(True)
    batch.x = x
    out = m(batch)
    out.sum().backward(retain_graph=True)
    oracle = common.load_seeded(O.MaskEmbdMultiMPN(**common.model_kwargs("tiny"))).eval()
    cb = common.make_batch("tiny")
    cb.x = cb.x.clone().requires_grad_(True)
    oracle(cb).sum().backward()
    _close(x.grad, cb.x.grad, "d out / d x")
    with pytest.raises(RuntimeError):
        out.sum().backward()


def test_cpu_tensors_are_refused_loudly():
    from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
    m = MaskEmbdMultiMPN(**common.model_kwargs("tiny"))
    with pytest.raises(RuntimeError, match="CUDA"):
        m(common.make_batch("tiny"))


def test_cuda_graph_step_matches_eager_step():
    """GraphedMSEStep (captured forward + MSE + backward) replays give the same loss/gradients as the eager path,
    follow new batch contents, and draw new dropout masks on every replay."""
    from poweflownet_b200.data import synthetic_batch
    from poweflownet_b200.training import GraphedMSEStep, fused_mse_step
    kw = dict(common.MODEL_DIMS, hidden_dim=33, n_gnn_layers=3, K=2, dropout_rate=0.0)
    m = _model(kw).train()
    b1, b2 = synthetic_batch("14", 8, seed=1).to(DEV), synthetic_batch("14", 8, seed=2).to(DEV)
    step = GraphedMSEStep(m, b1)
    for b in (b1, b2, b1):
        loss_g = float(step(b))
        grads_g = [p.grad.clone() for p in m.parameters()]
        loss_e = float(fused_mse_step(m, b))
        assert abs(loss_g - loss_e) <= 1e-6 * abs(loss_e)
        for g, p in zip(grads_g, m.parameters()):
            assert torch.equal(g, p.grad)  # same kernels, same order: bitwise identical
    kw["dropout_rate"] = 0.3
    m2 = _model(kw).train()
    step2 = GraphedMSEStep(m2, b1)
    l1, l2 = float(step2(b1)), float(step2(b1))
    assert l1 != l2  # the device-resident seed changed between replays


def test_loss_curve_matches_oracle_over_optimizer_steps():
    """`train_epoch`'s loop (utils/training.py:55-77) with AdamW (train.py:123) for 12 steps, dropout off so both sides are
    deterministic: the CUDA model's loss curve tracks the CPU oracle's."""
    from poweflownet_b200.data import synthetic_batch
    kw = dict(common.MODEL_DIMS, hidden_dim=33, n_gnn_layers=3, K=3, dropout_rate=0.0)
    oracle = common.load_seeded(O.MaskEmbdMultiMPN(**kw)).train()
    mine = _model(kw, oracle.state_dict()).train()
    opt_o = torch.optim.AdamW(oracle.parameters(), lr=1e-3)
    opt_m = torch.optim.AdamW(mine.parameters(), lr=1e-3)
    batches = [synthetic_batch("14", 16, seed=100 + i) for i in range(4)]
    curve_o, curve_m = [], []
    for step in range(12):
        b = batches[step % 4]
        opt_o.zero_grad()
        lo = torch.nn.functional.mse_loss(oracle(b), b.y)
        lo.backward()
        opt_o.step()
        db = b.to(DEV)
        opt_m.zero_grad()
        lm = torch.nn.functional.mse_loss(mine(db), db.y)
        lm.backward()
        opt_m.step()
        curve_o.append(float(lo))
        curve_m.append(float(lm))
    assert curve_o[-1] < curve_o[0]  # it trains
    for a, b in zip(curve_m, curve_o):
        assert abs(a - b) <= 1e-4 * abs(b), (curve_m, curve_o)


@pytest.mark.parametrize("case,graphs,hidden,layers,repeats", [("118v2", 128, 129, 4, 30), ("6470rte", 1, 512, 3, 6)])
def test_gradients_are_bitwise_reproducible(case, graphs, hidden, layers, repeats):
    """No atomics anywhere on the path (fixed-order split-K sums in k_wgrad_group_reduce, segment sums in edge order): every
    repeat of the step reproduces the first run's 35+ gradient tensors bit for bit, on the graph-resident route (bench
    size; grouped weight gradients with dY^T in tensor memory) and on the layer-wise route (hidden 512)."""
    from poweflownet_b200.data import synthetic_batch
    from poweflownet_b200.training import fused_mse_step
    kw = dict(common.MODEL_DIMS, hidden_dim=hidden, n_gnn_layers=layers, K=3, dropout_rate=0.0)
    m = _model(kw).train()
    batch = synthetic_batch(case, graphs).to(DEV)
    first_loss = float(fused_mse_step(m, batch))
    first = [p.grad.clone() for p in m.parameters()]
    for _ in range(repeats):
        loss = float(fused_mse_step(m, batch))
        assert loss == first_loss
        differing = [k for (k, p), g in zip(m.named_parameters(), first) if not torch.equal(p.grad, g)]
        assert not differing, differing
