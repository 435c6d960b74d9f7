"""Device-resident dataset + on-GPU batch assembly (`poweflownet_b200.datasets.PowerFlowData`, `pfn_batch_assemble`)
against the fixtures produced by the reference's own datasets/PowerFlowData.py + DataLoader, and against the CPU oracle
at BASELINE config 2 size.  Every tensor of the batch must be bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import common
import make_golden_dataset as mgd
from oracle import pfn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _raws(gold):
    return [(gold["raw"][tag]["edge_features"], gold["raw"][tag]["node_features"]) for tag in gold["raw"]]


def _dataset(gold, task, **kw):
    from poweflownet_b200.datasets import PowerFlowData
    return PowerFlowData(case=gold["meta"]["case"], split=mgd.SPLIT, task=task, device=DEV, raw=_raws(gold), **kw)


def _same(batch, want):
    for f in mgd.FIELDS:
        got = getattr(batch, f).cpu()
        assert got.dtype == want[f].dtype and got.shape == want[f].shape, f
        assert torch.equal(got, want[f]), f


@pytest.mark.parametrize("name", list(mgd.DATASET_CASES))
def test_batches_match_the_reference_loader(name):
    torch.set_num_threads(1)
    gold = torch.load(mgd.dataset_golden_path(name), weights_only=False)
    for task, sp in gold["splits"].items():
        ds = _dataset(gold, task)
        assert len(ds) == sp["len"] and ds.get_data_dimensions() == tuple(sp["dims"])
        for a, b in zip(ds.get_data_means_stds(), sp["stats"]):
            assert a.shape == b.shape and torch.equal(a, b)
        for bt in sp["batches"]:
            _same(ds.batch(bt["ids"]), bt)
        bs = gold["meta"]["batch_size"]
        loaded = list(ds.loader(batch_size=bs, shuffle=False))
        assert len(loaded) == len(sp["batches"]) == ds.num_batches(bs)
        for got, bt in zip(loaded, sp["batches"]):
            _same(got, bt)
    tr = gold["splits"]["train"]["stats"]
    ds = _dataset(gold, "val", xymean=tr[0], xystd=tr[1], edgemean=tr[2], edgestd=tr[3])
    _same(ds.batch(gold["val_with_train_stats"]["ids"]), gold["val_with_train_stats"])


def test_arbitrary_sample_orders_and_repeats_mixed_sizes():
    gold = torch.load(mgd.dataset_golden_path("ds_mixed"), weights_only=False)
    ds = _dataset(gold, "train")
    samples = O.process_split(_raws(gold), mgd.SPLIT, "train")
    stats = gold["splits"]["train"]["stats"]
    for ids in ([14, 0, 7, 7, 3, 9, 1], [5], list(range(14, -1, -1))):
        want = O.collate_batch(samples, ids, stats)
        _same(ds.batch(ids), want)
    raw = _dataset(gold, "train", normalize=False)
    _same(raw.batch([2, 11, 4]), O.collate_batch(samples, [2, 11, 4], None))


def test_shuffled_loader_visits_every_sample_once():
    gold = torch.load(mgd.dataset_golden_path("ds_case14"), weights_only=False)
    ds = _dataset(gold, "train")
    g = torch.Generator().manual_seed(0)
    ys = torch.cat([b.y.cpu() for b in ds.loader(batch_size=6, shuffle=True, generator=g)])
    everything = ds.batch(list(range(len(ds)))).y.cpu()
    assert ys.shape == everything.shape and not torch.equal(ys, everything)
    key = lambda t: sorted(map(tuple, t.reshape(-1, 14 * 4).tolist()))  # noqa: E731 -- one row per graph
    assert key(ys) == key(everything)
    assert sum(1 for _ in ds.loader(batch_size=6, drop_last=True)) == len(ds) // 6


def test_random_bus_type_transform_touches_only_bus_type():
    gold = torch.load(mgd.dataset_golden_path("ds_case14"), weights_only=False)
    plain, rnd = _dataset(gold, "train"), _dataset(gold, "train", random_bus_type=True)
    ids = list(range(16))
    a, b, c = plain.batch(ids), rnd.batch(ids, seed=1), rnd.batch(ids, seed=2)
    for f in mgd.FIELDS:
        if f != "bus_type":
            assert torch.equal(getattr(a, f), getattr(b, f)), f
    assert set(b.bus_type.unique().tolist()) == {0, 1} and not torch.equal(b.bus_type, c.bus_type)
    assert abs(float(b.bus_type.float().mean()) - 0.5) < 0.15
    assert torch.equal(b.bus_type, rnd.batch(ids, seed=1).bus_type)


def test_bad_sample_ids_are_refused_on_host_and_device():
    from poweflownet_b200._lib import lib
    gold = torch.load(mgd.dataset_golden_path("ds_case14"), weights_only=False)
    ds = _dataset(gold, "train")
    with pytest.raises(IndexError):
        ds.batch([0, len(ds)])
    with pytest.raises(ValueError):
        ds.batch([])
    # device-side guard: ids handed over on the GPU that disagree with the host's list leave the outputs untouched
    lying = torch.tensor([0, 10 ** 6], dtype=torch.int64, device=DEV)
    out = ds.batch([0, 1], ids_device=lying)
    torch.cuda.synchronize()
    flag = C.c_int32(0)
    scratch = torch.empty(int(lib().pfn_batch_assemble_scratch_bytes(2)), dtype=torch.uint8, device=DEV)
    rc = lib().pfn_batch_assemble(ds._cases, 1, lying.data_ptr(), 2, 28, 40, ds._norm, 0, out.x.data_ptr(), out.y.data_ptr(),
                                  out.bus_type.data_ptr(), out.pred_mask.data_ptr(), out.edge_index.data_ptr(),
                                  out.edge_attr.data_ptr(), out.batch.data_ptr(), out.ptr.data_ptr(), scratch.data_ptr(),
                                  torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    assert lib().pfn_batch_assemble_status(scratch.data_ptr(), 2, C.byref(flag), torch.cuda.current_stream().cuda_stream) == 0
    assert flag.value == 1


def test_reads_the_reference_file_layout(tmp_path):
    from poweflownet_b200.datasets import PowerFlowData
    gold = torch.load(mgd.dataset_golden_path("ds_mixed"), weights_only=False)
    os.makedirs(tmp_path / "raw")
    for tag, g in gold["raw"].items():
        np.save(tmp_path / "raw" / f"case{tag}_edge_features.npy", g["edge_features"].numpy().astype(np.float64))
        np.save(tmp_path / "raw" / f"case{tag}_node_features.npy", g["node_features"].numpy().astype(np.float64))
    ds = PowerFlowData(root=str(tmp_path), case="mixed", split=mgd.SPLIT, task="test", device=DEV)
    sp = gold["splits"]["test"]
    assert len(ds) == sp["len"]
    for bt in sp["batches"]:
        _same(ds.batch(bt["ids"]), bt)
    with pytest.raises(RuntimeError):
        PowerFlowData(case="14", split=mgd.SPLIT, device="cpu", raw=_raws(gold))


def test_full_size_batch_feeds_the_model():
    """BASELINE config 2: 128 graphs of case118v2 drawn from a 250-sample split; bit-exact against the oracle's collation,
    and the tiled forward accepts the assembled batch (its `ptr` marks 118-bus graphs)."""
    from poweflownet_b200.datasets import PowerFlowData
    from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
    e, nf = mgd.synthetic_raw(118, 186, 500, seed=3)
    raws = [(torch.from_numpy(e), torch.from_numpy(nf))]
    ds = PowerFlowData(case="118v2", split=mgd.SPLIT, task="train", device=DEV, raw=raws)
    assert len(ds) == 250
    samples = O.process_split(raws, mgd.SPLIT, "train")
    g = torch.Generator().manual_seed(5)
    ids = torch.randperm(250, generator=g)[:128]
    batch = ds.batch(ids)
    want = O.collate_batch(samples, ids, ds.get_data_means_stds())
    _same(batch, want)
    kw = common.model_kwargs("case118_standard")
    model = common.load_seeded(MaskEmbdMultiMPN(**kw)).to(DEV).eval()
    with torch.no_grad():
        out = model(batch)
    ref = common.load_seeded(O.MaskEmbdMultiMPN(**kw)).eval()
    with torch.no_grad():
        want_out = ref(common.GraphBatch(**want))
    assert max(common.rel_err(out.cpu(), want_out)) < 1e-5


@pytest.mark.parametrize("loss_name", ["mse", "masked_l2", "power_imbalance", "mixed"])
def test_train_epoch_matches_the_reference_loop(loss_name):
    """`training.train_epoch` (utils/training.py:30-80) over the device-resident loader with the one-launch AdamW,
    against the same epoch on the CPU: oracle model, oracle losses, torch.optim.AdamW, the reference's loop."""
    import torch.nn.functional as F
    from poweflownet_b200 import losses
    from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
    from poweflownet_b200.optim import FusedAdamW
    from poweflownet_b200.training import train_epoch
    torch.set_num_threads(1)
    gold = torch.load(mgd.dataset_golden_path("ds_case14"), weights_only=False)
    ds = _dataset(gold, "train")
    stats = ds.get_data_means_stds()
    kw = common.model_kwargs("case14_small")
    kw["dropout_rate"] = 0.0
    # CPU: the reference's loop, restated
    ref = common.load_seeded(O.MaskEmbdMultiMPN(**kw)).train()
    opt = torch.optim.AdamW(ref.parameters(), lr=1e-3, foreach=False)
    samples = O.process_split(_raws(gold), mgd.SPLIT, "train")
    ref_losses, n_attr = [], 8
    for bt in gold["splits"]["train"]["batches"]:
        data = common.GraphBatch(**O.collate_batch(samples, bt["ids"], gold["splits"]["train"]["stats"]))
        opt.zero_grad()
        out = ref(data)
        if loss_name == "mse":
            loss = F.mse_loss(out, data.y)
        elif loss_name == "masked_l2":
            loss = O.masked_l2_loss(out, data.y, data.pred_mask)
        elif loss_name == "power_imbalance":
            loss = O.power_imbalance(out * data.pred_mask + data.x * (1 - data.pred_mask), data.edge_index, data.edge_attr, *stats)
        else:
            loss = O.mixed_mse_power_imbalance(out, data.edge_index, data.edge_attr, data.y, stats, alpha=0.9)
        loss.backward()
        opt.step()
        ref_losses.append(float(loss.detach()))
    want = sum(v * n_attr for v in ref_losses) / (n_attr * len(ref_losses))
    # GPU
    model = common.load_seeded(MaskEmbdMultiMPN(**kw))
    fn = {"mse": torch.nn.MSELoss(), "masked_l2": losses.Masked_L2_loss(),
          "power_imbalance": losses.PowerImbalance(*stats),
          "mixed": losses.MixedMSEPoweImbalance(*stats, alpha=0.9)}[loss_name]
    opt2 = FusedAdamW(model.parameters(), lr=1e-3)
    got = train_epoch(model, ds.loader(batch_size=16, shuffle=False), fn, opt2, DEV)
    assert abs(got - want) <= 2e-4 * abs(want), (got, want, ref_losses)
    assert next(model.parameters()).is_cuda and model.training


def test_graphed_epochs_equal_the_eager_loop():
    """`training.GraphedEpochs` (assembly into the static buffers of a captured step + replay + AdamW) walks the same
    trajectory as `training.train_epoch` on the same sample order."""
    from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
    from poweflownet_b200.optim import FusedAdamW
    from poweflownet_b200.training import GraphedEpochs, train_epoch
    gold = torch.load(mgd.dataset_golden_path("ds_case14"), weights_only=False)
    ds = _dataset(gold, "train")  # 20 samples
    kw = common.model_kwargs("case14_small")
    kw["dropout_rate"] = 0.0
    runs = []
    for graphed in (False, True):
        model = common.load_seeded(MaskEmbdMultiMPN(**kw)).to(DEV)
        opt = FusedAdamW(model.parameters(), lr=1e-3)
        if graphed:
            runner = GraphedEpochs(model, ds, 10, opt)
            losses = [runner.run_epoch(shuffle=False) for _ in range(3)]
        else:
            losses = [train_epoch(model, ds.loader(10, shuffle=False), torch.nn.MSELoss(), opt, DEV) for _ in range(3)]
        runs.append((losses, [p.detach().clone() for p in model.parameters()]))
    for a, b in zip(runs[0][0], runs[1][0]):
        assert abs(a - b) <= 1e-6 * abs(a), runs
    assert runs[0][0][2] < runs[0][0][0]
    for p, q in zip(runs[0][1], runs[1][1]):
        assert torch.allclose(p, q, rtol=0, atol=2e-6)
    # shuffled epochs still see every sample once: 2 full batches of 10
    g = torch.Generator().manual_seed(0)
    assert runner.run_epoch(shuffle=True, generator=g) > 0


@pytest.mark.parametrize("loss_name", ["mse", "masked_l2", "power_imbalance", "mixed"])
def test_evaluate_epoch_matches_the_reference_loop(loss_name):
    """`training.evaluate_epoch` (utils/evaluation.py:53-104) on the device loader against the same loop on the CPU
    oracle, mixed-size dataset (118- and 14-bus graphs in one batch)."""
    import torch.nn.functional as F
    from poweflownet_b200 import losses
    from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
    from poweflownet_b200.training import evaluate_epoch
    torch.set_num_threads(1)
    gold = torch.load(mgd.dataset_golden_path("ds_mixed"), weights_only=False)
    ds = _dataset(gold, "test")  # 9 samples, batches of 6 + 3
    stats = ds.get_data_means_stds()
    kw = common.model_kwargs("mixed")
    ref = common.load_seeded(O.MaskEmbdMultiMPN(**kw)).eval()
    vals = []
    with torch.no_grad():
        for bt in gold["splits"]["test"]["batches"]:
            data = common.GraphBatch(**{f: bt[f] for f in mgd.FIELDS})
            out = ref(data)
            if loss_name == "mse":
                vals.append(float(F.mse_loss(out, data.y)))
            elif loss_name == "masked_l2":
                vals.append(float(O.masked_l2_loss(out, data.y, data.pred_mask)))
            elif loss_name == "power_imbalance":
                masked = out * data.pred_mask + data.pred_mask * (1 - data.pred_mask)
                vals.append(float(O.power_imbalance(masked, data.edge_index, data.edge_attr, *stats)))
            else:
                vals.append(float(O.mixed_mse_power_imbalance(out, data.edge_index, data.edge_attr, data.y, stats, alpha=0.9)))
    want = sum(vals) / len(vals)
    model = common.load_seeded(MaskEmbdMultiMPN(**kw)).to(DEV)
    fn = {"mse": torch.nn.MSELoss(), "masked_l2": losses.Masked_L2_loss(),
          "power_imbalance": losses.PowerImbalance(*stats),
          "mixed": losses.MixedMSEPoweImbalance(*stats, alpha=0.9)}[loss_name]
    got = evaluate_epoch(model, ds.loader(batch_size=6, shuffle=False), fn, DEV)
    assert abs(got - want) <= 2e-5 * abs(want), (got, want)
    assert not model.training


def test_graphed_epochs_with_the_parser_default_loss():
    """`GraphedEpochs(loss="masked_l2")` (Masked_L2_loss, the reference's default --train_loss_fn,
    utils/argument_parser.py:36-37) walks the same trajectory as the eager `train_epoch` with the loss module."""
    from poweflownet_b200.losses import Masked_L2_loss
    from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
    from poweflownet_b200.optim import FusedAdamW
    from poweflownet_b200.training import GraphedEpochs, train_epoch
    gold = torch.load(mgd.dataset_golden_path("ds_case14"), weights_only=False)
    ds = _dataset(gold, "train")  # 20 samples
    kw = common.model_kwargs("case14_small")
    kw["dropout_rate"] = 0.0
    runs = []
    for graphed in (False, True):
        model = common.load_seeded(MaskEmbdMultiMPN(**kw)).to(DEV)
        opt = FusedAdamW(model.parameters(), lr=1e-3)
        if graphed:
            runner = GraphedEpochs(model, ds, 10, opt, loss="masked_l2", regularize=True, regcoeff=0.5)
            losses = [runner.run_epoch(shuffle=False) for _ in range(3)]
        else:
            losses = [train_epoch(model, ds.loader(10, shuffle=False), Masked_L2_loss(regularize=True, regcoeff=0.5), opt, DEV) for _ in range(3)]
        runs.append((losses, [p.detach().clone() for p in model.parameters()]))
    for a, b in zip(runs[0][0], runs[1][0]):
        assert abs(a - b) <= 1e-6 * abs(a), runs
    assert runs[0][0][2] < runs[0][0][0]
    for p, q in zip(runs[0][1], runs[1][1]):
        assert torch.allclose(p, q, rtol=0, atol=2e-6)


def test_graphed_epochs_on_a_mixed_size_dataset_bucket_by_batch_shape():
    """`--case mixed` (118- and 14-bus samples in one dataset, datasets/PowerFlowData.py:67-70): the shape of a mini-batch
    depends on how many samples of each case it drew; `GraphedEpochs` captures one graph per shape on first sight and
    replays it afterwards, and the trajectory equals the eager loop's on the same sample order."""
    from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
    from poweflownet_b200.optim import FusedAdamW
    from poweflownet_b200.training import GraphedEpochs, train_epoch
    gold = torch.load(mgd.dataset_golden_path("ds_mixed"), weights_only=False)
    ds = _dataset(gold, "train")
    assert len(set(int(v) for v in ds._n)) == 2  # two graph sizes
    kw = common.model_kwargs("mixed")
    kw["dropout_rate"] = 0.0
    bs = 4
    runs = []
    for graphed in (False, True):
        model = common.load_seeded(MaskEmbdMultiMPN(**kw)).to(DEV)
        opt = FusedAdamW(model.parameters(), lr=1e-3)
        losses = []
        for ep in range(3):
            g = torch.Generator().manual_seed(100 + ep)
            if graphed:
                if ep == 0:
                    runner = GraphedEpochs(model, ds, bs, opt, loss="mse", max_graphs=2)
                losses.append(runner.run_epoch(shuffle=True, generator=g))
            else:
                losses.append(train_epoch(model, ds.loader(bs, shuffle=True, generator=g, drop_last=True), torch.nn.MSELoss(), opt, DEV))
        runs.append((losses, [p.detach().clone() for p in model.parameters()]))
    for a, b in zip(runs[0][0], runs[1][0]):
        assert abs(a - b) <= 2e-6 * abs(a), runs
    for p, q in zip(runs[0][1], runs[1][1]):
        assert torch.allclose(p, q, rtol=0, atol=5e-6)
    assert 1 <= len(runner.steps) <= 2  # at most `max_graphs` captured shapes; any further shape ran eagerly


def test_evaluation_between_graph_replays_does_not_touch_the_captured_workspace():
    """a captured step has its activation / scratch workspace addresses baked in.  An eval forward of
    the SAME (N, E_raw) shape between replays (train, validate, train: the reference's epoch loop) must neither take nor
    free that workspace: replays before and after the evaluation produce the same gradients for the same batch, also after
    the caching allocator has been emptied and refilled."""
    from poweflownet_b200.data import synthetic_batch
    from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
    from poweflownet_b200.training import GraphedMSEStep
    kw = dict(common.MODEL_DIMS, hidden_dim=64, n_gnn_layers=2, K=3, dropout_rate=0.0)
    model = common.load_seeded(MaskEmbdMultiMPN(**kw)).to(DEV).train()
    batch = synthetic_batch("14", 18, seed=3).to(DEV)
    other = synthetic_batch("14", 18, seed=4).to(DEV)
    step = GraphedMSEStep(model, batch)
    assert not model._pool  # the capture ran on a pool private to the step; nothing of it is reachable from the model
    loss0 = float(step(batch))
    g0 = [p.grad.clone() for p in model.parameters()]
    model.eval()
    with torch.no_grad():
        for _ in range(3):
            model(other)  # same shape: must run on workspaces of its own
    n_ws = sum(len(v) for v in model._pool.values())
    assert n_ws == 1  # ... which it hands back even under no_grad (it used to keep and drop them)
    torch.cuda.empty_cache()
    junk = [torch.full((1 << 22,), float("nan"), device=DEV) for _ in range(8)]  # reuse whatever the allocator freed
    model.train()
    loss1 = float(step(batch))
    assert loss1 == loss0
    for a, b in zip(g0, [p.grad for p in model.parameters()]):
        assert torch.equal(a, b)
    del junk


def test_graph_replays_draw_a_fresh_dropout_seed_each():
    """consecutive replays queued back to back (the host runs ahead of the device) must not share a dropout
    seed; under a fixed torch seed the sequence is reproducible."""
    from poweflownet_b200.data import synthetic_batch
    from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
    from poweflownet_b200.training import GraphedMSEStep
    kw = dict(common.MODEL_DIMS, hidden_dim=64, n_gnn_layers=2, K=3, dropout_rate=0.5)
    batch = synthetic_batch("14", 18, seed=3).to(DEV)
    seqs = []
    for _ in range(2):
        model = common.load_seeded(MaskEmbdMultiMPN(**kw)).to(DEV).train()
        step = GraphedMSEStep(model, batch)
        torch.manual_seed(99)
        total = torch.zeros(200, device=DEV)
        for i in range(200):  # queued without any synchronisation
            total[i:i + 1] = step(batch)
        seqs.append(total.cpu())
    assert torch.equal(seqs[0], seqs[1])  # reproducible
    assert len(set(seqs[0].tolist())) > 190  # and (almost surely) all different: no two replays shared a mask
