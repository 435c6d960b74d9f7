"""The graph-resident forward (`pfn_mpn_forward_tiled`, csrc/fused_fwd.cu): one kernel for the whole layer stack,
one tile of whole graphs per CTA.  Checked against the reference goldens, against the layer-wise kernels, and for its
behaviour when the caller's closed-tile promise does not hold (the kernel validates it itself)."""
import ctypes as C

import pytest
import torch

import common
from oracle import pfn_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5
DEV = "cuda:0"
FUSED_CASES = ["case14_small", "case118_standard"]  # equal-sized graphs, hidden_dim 64 / 129


def _model(kw, fused=True):
    from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
    m = common.load_seeded(MaskEmbdMultiMPN(**kw)).to(DEV)
    m.fused = fused
    return m


def _close(got, want, what, tol=TOL):
    e = common.rel_err(got.detach().cpu(), want.detach().cpu())
    assert max(e) < tol, (what, e)


def _launches():
    from poweflownet_b200 import _lib
    return int(_lib.lib().pfn_launch_count())


@pytest.mark.parametrize("name", FUSED_CASES)
def test_fused_path_is_taken_and_matches_reference(name):
    gold = torch.load(common.golden_path(name), weights_only=True)
    m = _model(gold["meta"]["model_kwargs"]).eval()
    batch = common.GraphBatch(**gold["inputs"]).to(DEV)
    assert m._tile_rows(batch) > 0, "this batch should be eligible for the graph-resident kernel"
    with torch.no_grad():
        m(batch)  # first call of a shape also reads back the tile validation
        n0 = _launches()
        out = m(batch)
        n_fused = _launches() - n0
    assert all(m._tiling_checked.values()) and len(m._tiling_checked) == 1
    # graph prep (1 launch when the batch is laid out tile by tile, else 5) + weight packing (1) + ONE forward kernel
    assert n_fused <= 7, n_fused
    _close(out, gold["eval_out"], "eval_out vs fp32 reference")
    _close(out, gold["eval_out_fp64"].float(), "eval_out vs fp64 twin")
    m.fused = False
    with torch.no_grad():
        n0 = _launches()
        out_lw = m(batch)
        n_layerwise = _launches() - n0
    assert n_layerwise > 2 * n_fused
    _close(out, out_lw, "graph-resident vs layer-wise forward", tol=3e-6)


@pytest.mark.parametrize("name", FUSED_CASES)
def test_fused_train_step_matches_reference(name):
    gold = torch.load(common.golden_path(name), weights_only=True)
    m = _model(gold["meta"]["model_kwargs"]).train()
    batch = common.GraphBatch(**gold["inputs"]).to(DEV)
    m._inject_dropout_masks = common.dropout_masks(name, batch.num_nodes)
    out = m(batch)
    assert m._tiling_checked and all(m._tiling_checked.values())
    loss = torch.nn.functional.mse_loss(out, batch.y)
    loss.backward()
    _close(out, gold["train_out"], "train_out")
    assert abs(float(loss) - float(gold["train_loss"])) < TOL * abs(float(gold["train_loss"]))
    for k, p in m.named_parameters():
        if "grads" in gold:
            _close(p.grad, gold["grads"][k], k)
        else:
            nrm = float(gold["grad_norm"][k])
            assert abs(float(p.grad.double().norm()) - nrm) < TOL * nrm + 1e-12, k


def test_saved_activations_match_layerwise():
    """The fused forward must leave the activation workspace the backward kernels read (Hi, Hj, S, x_k, Y, t1, x0)
    as the layer-wise forward does: compare the gradients a layer-wise backward computes from either."""
    from poweflownet_b200.data import synthetic_batch
    kw = dict(common.MODEL_DIMS, hidden_dim=129, n_gnn_layers=3, K=2, dropout_rate=0.0)
    batch = synthetic_batch("118v2", 5).to(DEV)
    grads = {}
    for fused in (True, False):
        m = _model(kw, fused).train()
        torch.nn.functional.mse_loss(m(batch), batch.y).backward()
        grads[fused] = {k: p.grad.clone() for k, p in m.named_parameters()}
    for k in grads[True]:
        _close(grads[True][k], grads[False][k], k, tol=3e-6)


def test_dropout_generator_statistics_and_determinism():
    """Counter-based dropout inside the fused epilogues: same seed tensor -> same output, and the same keep masks as
    the layer-wise path draws."""
    from poweflownet_b200.data import synthetic_batch
    kw = dict(common.MODEL_DIMS, hidden_dim=129, n_gnn_layers=2, K=3, dropout_rate=0.2)
    batch = synthetic_batch("118v2", 8).to(DEV)
    m = _model(kw).train()
    m._seed_device = torch.tensor([1234567], dtype=torch.int64, device=DEV)
    with torch.no_grad():
        a, b = m(batch), m(batch)
        m._seed_device = torch.tensor([7654321], dtype=torch.int64, device=DEV)
        c = m(batch)
    assert torch.equal(a, b) and not torch.equal(a, c)
    with torch.no_grad():
        d = m.eval()(batch)
    assert not torch.equal(a, d)  # dropout really was applied in train mode
    # the generator is shared with the layer-wise epilogues (same hash of (row, column, layer, seed)): same masks
    m.train()
    m.fused = False
    with torch.no_grad():
        e = m(batch)
    _close(c, e, "same seed, graph-resident vs layer-wise", tol=3e-6)


def test_broken_tile_promise_is_detected_and_falls_back():
    """Equal N / num_graphs but an edge that crosses the assumed tile boundary: the kernel flags it, the module
    re-runs the layer-wise path for that shape, and the result is still the reference's."""
    from poweflownet_b200.data import synthetic_batch
    kw = dict(common.MODEL_DIMS, hidden_dim=64, n_gnn_layers=2, K=3, dropout_rate=0.0)
    batch = synthetic_batch("14", 18)  # 252 nodes: tiles of 9 graphs = 126 rows
    ei = batch.edge_index.clone()
    ei[:, -1] = torch.tensor([125, 126])  # joins the last node of tile 0 to the first node of tile 1
    bad = common.GraphBatch(**{f: (ei if f == "edge_index" else getattr(batch, f)) for f in
                               ("x", "y", "bus_type", "pred_mask", "edge_index", "edge_attr", "batch", "ptr")})
    oracle = common.load_seeded(O.MaskEmbdMultiMPN(**kw)).eval()
    with torch.no_grad():
        want = oracle(bad)
    m = _model(kw).eval()
    db = bad.to(DEV)
    assert m._tile_rows(db) == 126
    with torch.no_grad():
        out = m(db)
    # every tiling was tried (equal tiles on the one-launch and on the general preparation, then whole graphs packed from
    # `ptr`: the same boundary) and refused
    assert len(m._tiling_checked) >= 2 and not any(m._tiling_checked.values())
    assert m._tile_rows(db) == 0  # the shape is remembered as not tileable
    assert not torch.isnan(out).any()
    _close(out, want, "fallback result")
    with torch.no_grad():
        _close(m(db), want, "second call (layer-wise from the start)")


def test_uniform_tiling_refused_then_variable_tiles_accepted():
    """N divisible by num_graphs although the graphs differ in size (10- and 18-bus grids alternating): the equal-tile
    guess cuts a graph in two, the kernel refuses it, and the module retries with whole graphs packed from `ptr` on the
    graph-resident route instead of dropping to the layer-wise kernels."""
    from poweflownet_b200.data import synthetic_batch
    kw = dict(common.MODEL_DIMS, hidden_dim=64, n_gnn_layers=2, K=3, dropout_rate=0.0)
    batch = synthetic_batch(cases=[(10, 13), (18, 25)] * 6)  # 168 nodes, 12 graphs: N / G = 14 -> tiles of 126 rows guessed
    oracle = common.load_seeded(O.MaskEmbdMultiMPN(**kw)).eval()
    with torch.no_grad():
        want = oracle(batch)
    m = _model(kw).eval()
    db = batch.to(DEV)
    assert m._tiling(db)[0] == 126 and m._tiling(db)[2] is not None
    n0 = _launches()
    with torch.no_grad():
        out = m(db)
    first = _launches() - n0
    e = int(db.edge_index.size(1))
    # (168 rows are not a whole number of 126-row tiles, so the one-launch preparation does not apply to this shape)
    assert m._tiling_checked == {(168, e, 126, 0): False, (168, e, 128, 12): True}
    _close(out, want, "variable tiles after a refused uniform tiling")
    assert m._tiling(db)[:1] == (128,) and m._tiling(db)[1] is not None
    n0 = _launches()
    with torch.no_grad():
        _close(m(db), want, "second call (variable tiles from the start)")
    assert _launches() - n0 < first


def test_tile_validation_is_repeated_periodically():
    """The closed-tile flag is read back for the first batch of a shape and then every `_tiling_recheck` batches: a later
    batch of the same shape that breaks the promise is caught at the next re-check (its own output rows are NaN-poisoned
    by the kernel in the meantime, never silently wrong)."""
    from poweflownet_b200.data import synthetic_batch
    kw = dict(common.MODEL_DIMS, hidden_dim=64, n_gnn_layers=2, K=3, dropout_rate=0.0)
    import types
    good = synthetic_batch("14", 18)
    ei = good.edge_index.clone()
    ei[:, -1] = torch.tensor([125, 126])
    bad = common.GraphBatch(**{f: (ei if f == "edge_index" else getattr(good, f)) for f in
                               ("x", "y", "bus_type", "pred_mask", "edge_index", "edge_attr", "batch", "ptr")})
    oracle = common.load_seeded(O.MaskEmbdMultiMPN(**kw)).eval()
    with torch.no_grad():
        want_bad = oracle(bad)
    m = _model(kw).eval()
    m._tiling_recheck = 3

    def no_ptr(b):  # a batch object without `ptr` (only num_graphs): no variable-tile second attempt
        d = b.to(DEV)
        return types.SimpleNamespace(x=d.x, y=d.y, pred_mask=d.pred_mask, edge_index=d.edge_index, edge_attr=d.edge_attr,
                                     num_graphs=b.num_graphs)
    dg, dbad = no_ptr(good), no_ptr(bad)
    sig = (good.num_nodes, int(good.edge_index.size(1)), 126, 0)
    fast_sig = ("prep_tiled", good.num_nodes, int(good.edge_index.size(1)), 126)  # equal tiles on the one-launch preparation
    with torch.no_grad():
        m(dg)                      # call 0: validated
        assert m._tiling_checked == {fast_sig: True}
        out1 = m(dbad)             # call 1: not re-validated -> the broken tiles come back NaN, never silently wrong
        assert torch.isnan(out1).any()
        m(dg)                      # call 2
        out3 = m(dbad)             # call 3: re-validated -> refused -> layer-wise result
    assert m._tiling_checked[fast_sig] is False and m._tiling_checked[sig] is False
    _close(out3, want_bad, "layer-wise result after the periodic re-check")


def test_tiled_entry_point_rejects_unsupported_configurations():
    from poweflownet_b200 import _lib
    from poweflownet_b200._lib import MpnDesc
    lib = _lib.lib()
    ok = MpnDesc(4, 2, 4, 129, 4, 3, 0.2, 0)
    assert lib.pfn_mpn_fused_supported(C.byref(ok), 118) == 1
    assert lib.pfn_mpn_fused_supported(C.byref(ok), 129) == 0          # more than 128 rows per tile
    assert lib.pfn_mpn_fused_supported(C.byref(MpnDesc(4, 2, 4, 512, 5, 3, 0.2, 0)), 118) == 0   # configs/large.json width
    assert lib.pfn_mpn_fused_supported(C.byref(MpnDesc(4, 2, 4, 33, 4, 3, 0.2, 0)), 118) == 0
    assert lib.pfn_mpn_fused_supported(C.byref(MpnDesc(4, 2, 4, 64, 2, 3, 0.2, 0)), 126) == 1
    rc = lib.pfn_mpn_forward_tiled(C.byref(MpnDesc(4, 2, 4, 512, 5, 3, 0.2, 0)), None, None, None, 0, 0, None, None, None, 0, 0,
                                   None, None, None, 118, None, 0, None)
    assert rc == -2 and b"graph-resident" in lib.pfn_last_error()


def test_pipelined_steps_match_eager_steps():
    """PipelinedMSESteps (double-buffered CUDA-graph replays, H2D of batch i+1 under the step of batch i) returns the
    losses and leaves the gradients of plain eager steps on the same batches."""
    from poweflownet_b200.data import synthetic_batch
    from poweflownet_b200.training import PipelinedMSESteps, fused_mse_step
    kw = dict(common.MODEL_DIMS, hidden_dim=129, n_gnn_layers=3, K=3, dropout_rate=0.0)
    host = [synthetic_batch("118v2", 6, seed=100 + i).pin_memory() for i in range(4)]
    m = _model(kw).train()
    want, want_grads = [], None
    for b in host:
        want.append(float(fused_mse_step(m, b.to(DEV)).item()))
        want_grads = [p.grad.clone() for p in m._engine_params()]
    pipe = PipelinedMSESteps(m, host[0].to(DEV))
    got = []
    pipe.prefetch(host[0])
    for i in range(len(host)):
        if i + 1 < len(host):
            pipe.prefetch(host[i + 1])
        got.append(float(pipe.step().item()))
    for a, b in zip(got, want):
        assert abs(a - b) <= 1e-6 * abs(b), (got, want)
    for p, g in zip(m._engine_params(), want_grads):
        _close(p.grad, g, "gradient after the last pipelined step", tol=1e-6)
    with pytest.raises(RuntimeError):
        pipe.step()


@pytest.mark.parametrize("cfg", [dict(hidden_dim=129, n_gnn_layers=3, K=3), dict(hidden_dim=64, n_gnn_layers=2, K=2)])
def test_chained_backward_equals_per_layer_backward(cfg, monkeypatch):
    """Mode 3 (the whole backward data path in one launch, gradient kept in the planes between layers) against one
    launch per layer (modes 1 / 2, PFN_BWD_CHAIN=0) and against the layer-wise kernels."""
    from poweflownet_b200.data import synthetic_batch
    kw = dict(common.MODEL_DIMS, dropout_rate=0.0, **cfg)
    batch = synthetic_batch("118v2" if cfg["hidden_dim"] == 129 else "14", 9).to(DEV)
    grads = {}
    for name, fused, chain in (("chain", True, "1"), ("per_layer", True, "0"), ("layerwise", False, "1")):
        monkeypatch.setenv("PFN_BWD_CHAIN", chain)
        m = _model(kw, fused).train()
        n0 = _launches()
        torch.nn.functional.mse_loss(m(batch), batch.y).backward()
        grads[name] = ({k: p.grad.clone() for k, p in m.named_parameters()}, _launches() - n0)
    assert grads["chain"][1] < grads["per_layer"][1] < grads["layerwise"][1]
    for k in grads["chain"][0]:
        _close(grads["chain"][0][k], grads["per_layer"][0][k], k + " (chain vs per-layer)", tol=1e-6)
        _close(grads["chain"][0][k], grads["layerwise"][0][k], k + " (chain vs layer-wise)", tol=3e-6)


@pytest.mark.parametrize("name", ["mixed", "case118_standard"])
@pytest.mark.parametrize("regularize", [True, False])
def test_fused_masked_l2_step_matches_reference(name, regularize):
    """`fused_masked_l2_step` (Masked_L2_loss value + gradient from one fused head, element counts on the device)
    against the oracle's restatement of utils/custom_loss_functions.py:10-46 + autograd."""
    from poweflownet_b200.training import fused_masked_l2_step
    gold = torch.load(common.golden_path(name), weights_only=True)
    kw = gold["meta"]["model_kwargs"]
    batch = common.GraphBatch(**gold["inputs"])
    masks = common.dropout_masks(name, batch.num_nodes)
    oracle = common.load_seeded(O.MaskEmbdMultiMPN(**kw)).train()
    out = oracle(batch, dropout_masks=masks)
    want = O.masked_l2_loss(out, batch.y, batch.pred_mask, regularize=regularize)
    want.backward()
    m = _model(kw).train()
    m._inject_dropout_masks = masks
    loss = fused_masked_l2_step(m, batch.to(DEV), regularize=regularize)
    assert abs(float(loss) - float(want)) < TOL * abs(float(want))
    for (k, p), (_, q) in zip(m.named_parameters(), oracle.named_parameters()):
        _close(p.grad, q.grad, k)


def test_mixed_size_batches_take_variable_tiles():
    """Batches that mix graph sizes (the reference's --case mixed: 118-bus and 14-bus grids) run on the graph-resident
    kernels with tiles packed from `ptr` on the device; results equal the layer-wise kernels' and the oracle's."""
    from poweflownet_b200.data import synthetic_batch
    kw = dict(common.MODEL_DIMS, hidden_dim=129, n_gnn_layers=3, K=3, dropout_rate=0.0)
    names = ["14", "118v2", "14", "14", "14", "118v2", "118v2", "14", "14", "14", "14", "14", "14", "14", "14", "14", "14", "118v2"]
    batch = synthetic_batch(cases=names)
    oracle = common.load_seeded(O.MaskEmbdMultiMPN(**kw)).train()
    loss_ref, out_ref = O.forward_loss_backward(oracle, batch, "mse")
    db = batch.to(DEV)
    grads, outs = {}, {}
    for fused in (True, False):
        m = _model(kw, fused).train()
        tile, ptr, _ = m._tiling(db)
        assert (tile, ptr is not None) == ((128, True) if fused else (0, False))
        n0 = _launches()
        out = m(db)
        torch.nn.functional.mse_loss(out, db.y).backward()
        outs[fused] = (out.detach().clone(), _launches() - n0)
        grads[fused] = {k: p.grad.clone() for k, p in m.named_parameters()}
        if fused:
            assert list(m._tiling_checked.values()) == [True]
    assert outs[True][1] < outs[False][1] / 3
    _close(outs[True][0], out_ref, "out vs oracle")
    _close(outs[True][0], outs[False][0], "out vs layer-wise", tol=3e-6)
    for (k, q) in oracle.named_parameters():
        _close(grads[True][k], q.grad, k + " vs oracle")
        _close(grads[True][k], grads[False][k], k + " vs layer-wise", tol=3e-6)


def test_variable_tiles_reject_large_graphs():
    """A batch that mixes a 6470-bus graph with small ones cannot be tiled: the device-side packing flags it and the
    module falls back to the layer-wise kernels for that shape."""
    from poweflownet_b200.data import synthetic_batch
    kw = dict(common.MODEL_DIMS, hidden_dim=64, n_gnn_layers=2, K=3, dropout_rate=0.0)
    batch = synthetic_batch(cases=["14", (300, 420), "14"])
    oracle = common.load_seeded(O.MaskEmbdMultiMPN(**kw)).eval()
    with torch.no_grad():
        want = oracle(batch)
    m = _model(kw).eval()
    db = batch.to(DEV)
    assert m._tiling(db)[0] == 128
    with torch.no_grad():
        out = m(db)
    assert list(m._tiling_checked.values()) == [False] and m._tiling(db)[0] == 0
    _close(out, want, "fallback result")
