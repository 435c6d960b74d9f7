"""Host-side routing of the nn.Module mirror (no GPU needed): which batches are promised to the graph-resident kernels,
and with what tiling.  The decision uses shape information only (N, num_graphs / ptr), never device data."""
import ctypes as C

import pytest
import torch

import common


def _model(**cfg):
    from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
    return MaskEmbdMultiMPN(**dict(common.MODEL_DIMS, dropout_rate=0.2, **cfg))


def test_uniform_batches_get_whole_graph_tiles():
    from poweflownet_b200.data import synthetic_batch
    m = _model(hidden_dim=129, n_gnn_layers=4, K=3)
    assert m._tiling(synthetic_batch("118v2", 5)) == (118, None, None)      # one 118-bus graph per 128-row tile
    assert m._tiling(synthetic_batch("14", 20)) == (126, None, None)        # nine 14-bus graphs per tile
    assert m._tiling(synthetic_batch("6470rte", 1)) == (0, None, None)      # too large: layer-wise kernels
    m.fused = False
    assert m._tiling(synthetic_batch("118v2", 5)) == (0, None, None)


def test_unsupported_widths_and_depths_take_the_layerwise_route():
    from poweflownet_b200.data import synthetic_batch
    b = synthetic_batch("118v2", 3)
    assert _model(hidden_dim=64, n_gnn_layers=2, K=3)._tiling(b)[0] == 118
    assert _model(hidden_dim=33, n_gnn_layers=4, K=3)._tiling(b)[0] == 0     # not 129, not a multiple of 16
    assert _model(hidden_dim=512, n_gnn_layers=5, K=3)._tiling(b)[0] == 0    # configs/large.json width
    assert _model(hidden_dim=129, n_gnn_layers=4, K=5)._tiling(b)[0] == 0    # more than four TAGConv segments
    assert _model(hidden_dim=129, n_gnn_layers=10, K=3)._tiling(b)[0] == 0   # 19 layers > the kernel's layer table


def test_mixed_sizes_need_ptr_on_the_device():
    from poweflownet_b200.data import synthetic_batch
    m = _model(hidden_dim=129, n_gnn_layers=4, K=3)
    mixed = synthetic_batch(cases=["14", "118v2", "14"])
    assert m._tiling(mixed) == (0, None, None)  # ptr is a CPU tensor here: the variable-size tiling is a device-side pass
    m._tiling_checked[(mixed.num_nodes, int(mixed.edge_index.size(1)), 128, 3)] = False
    assert m._tiling(mixed) == (0, None, None)


def test_a_failed_validation_is_remembered_per_shape():
    from poweflownet_b200.data import synthetic_batch
    m = _model(hidden_dim=129, n_gnn_layers=4, K=3)
    b = synthetic_batch("118v2", 4)
    assert m._tiling(b)[0] == 118
    m._tiling_checked[(b.num_nodes, int(b.edge_index.size(1)), 118, 0)] = False
    assert m._tiling(b) == (0, None, None)
    assert m._tiling(synthetic_batch("118v2", 5))[0] == 118  # another shape is unaffected


def test_fused_supported_matrix():
    from poweflownet_b200 import _lib
    from poweflownet_b200._lib import MpnDesc
    lib = _lib.lib()
    ok = lambda h, L, K, rows, nf=4, out=4, ef=2: lib.pfn_mpn_fused_supported(C.byref(MpnDesc(nf, ef, out, h, L, K, 0.2, 0)), rows)  # noqa: E731
    assert ok(129, 4, 3, 118) == 1 and ok(128, 4, 3, 128) == 1 and ok(64, 2, 0, 126) == 1
    assert ok(129, 4, 3, 0) == 0 and ok(129, 4, 3, 129) == 0
    assert ok(130, 4, 3, 118) == 0 and ok(48, 4, 3, 118) == 1 and ok(16, 4, 3, 118) == 0
    assert ok(129, 1, 3, 118) == 0            # the reference's single-layer constructor is shape-inconsistent
    assert ok(129, 4, 3, 118, nf=6) == 0 and ok(129, 4, 3, 118, out=5) == 0 and ok(129, 4, 3, 118, ef=4) == 0


def test_epoch_batches_partition_samples_across_ranks():
    """datasets.epoch_batches: the sample-id plan of one pass (pure host logic of the device-resident loader)."""
    import torch
    from poweflownet_b200.datasets import epoch_batches
    plain = epoch_batches(10, 4)
    assert [b.tolist() for b in plain] == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]]
    assert [b.tolist() for b in epoch_batches(10, 4, drop_last=True)] == [[0, 1, 2, 3], [4, 5, 6, 7]]
    world = 3
    per_rank = [epoch_batches(50, 4, True, torch.Generator().manual_seed(7), True, r, world) for r in range(world)]
    assert len({len(b) for b in per_rank}) == 1 and len(per_rank[0]) == (50 // 4) // world
    seen = torch.cat([torch.cat(b) for b in per_rank])
    assert seen.numel() == seen.unique().numel() == world * len(per_rank[0]) * 4  # disjoint, full batches only
    again = epoch_batches(50, 4, True, torch.Generator().manual_seed(7), True, 1, world)
    assert all(torch.equal(a, b) for a, b in zip(again, per_rank[1]))
    import pytest
    with pytest.raises(ValueError):
        epoch_batches(10, 4, rank=2, world=2)
