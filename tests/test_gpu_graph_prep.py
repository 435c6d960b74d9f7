"""Integer graph work on the B200 through the C ABI, bit-exact against the oracle / golden vectors:
is_directed, undirect_graph (networks/MPN.py:498-523), stable CSR by target and by source, degrees."""
import pytest
import torch

import common
from emulation import stable_csr
from oracle import pfn_oracle as O

pytestmark = pytest.mark.gpu
ALL = list(common.CASES)


def _graph(batch, mode=1):
    from poweflownet_b200 import ops
    dev = torch.device("cuda", 0)
    return ops.PreparedGraph(batch.edge_index.to(dev), batch.edge_attr.to(dev), batch.num_nodes, mode=mode)


def _check_csr(g, ei, ea, n):
    arr = {k: v.cpu() for k, v in g.arrays().items()}
    e = ei.size(1)
    for side, by_target in (("t", True), ("s", False)):
        rowptr, nbr, order = stable_csr(ei, n, by_target)
        assert torch.equal(arr[f"rowptr_{side}"].long(), rowptr), side
        assert torch.equal(arr[f"nbr_{side}"][:e].long(), nbr), side
        assert torch.equal(arr[f"eid_{side}"][:e].long(), order), side
        assert torch.equal(arr[f"ea_{side}"][:e], ea[order]), side
    deg = torch.bincount(ei[1], minlength=n).float()
    assert torch.equal(arr["deg"], deg)
    dis = torch.where(deg > 0, 1.0 / deg.sqrt(), torch.zeros_like(deg))
    assert torch.allclose(arr["dis"], dis, rtol=2e-7, atol=0)


@pytest.mark.parametrize("name", ALL)
def test_prep_matches_golden(name):
    gold = torch.load(common.golden_path(name), weights_only=True)
    batch = common.GraphBatch(**gold["inputs"])
    g = _graph(batch)
    directed, e, err = g.meta()
    assert err == 0
    assert directed == bool(gold["is_directed"])
    assert e == gold["undirected_edge_index"].size(1)
    ei, ea = g.export()
    assert ei.dtype == torch.int64 and torch.equal(ei.cpu(), gold["undirected_edge_index"])
    assert torch.equal(ea.cpu(), gold["undirected_edge_attr"])
    _check_csr(g, gold["undirected_edge_index"], gold["undirected_edge_attr"], batch.num_nodes)


@pytest.mark.parametrize("case,b", [("118v2", 128), ("6470rte", 4)])
def test_prep_full_size(case, b):
    from poweflownet_b200.data import synthetic_batch
    batch = synthetic_batch(case, b)
    g = _graph(batch)
    directed, e, err = g.meta()
    assert (directed, e, err) == (True, 2 * batch.edge_index.size(1), 0)
    ei, ea = O.undirect_graph(batch.edge_index, batch.edge_attr)
    _check_csr(g, ei, ea, batch.num_nodes)


def test_mode0_takes_edges_as_given():
    batch = common.make_batch("tiny")
    g = _graph(batch, mode=0)
    assert g.meta() == (False, batch.edge_index.size(1), 0)
    _check_csr(g, batch.edge_index, batch.edge_attr, batch.num_nodes)


def test_out_of_range_node_id_is_flagged():
    batch = common.make_batch("tiny")
    batch.edge_index[1, 3] = batch.num_nodes + 5
    g = _graph(batch)
    assert g.meta()[2] == 1


def test_module_helpers_match_reference_semantics():
    from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
    m = MaskEmbdMultiMPN(4, 2, 4, 8, 2, 2, 0.0)
    dev = torch.device("cuda", 0)
    for name in ALL:
        gold = torch.load(common.golden_path(name), weights_only=True)
        ei, ea = gold["inputs"]["edge_index"].to(dev), gold["inputs"]["edge_attr"].to(dev)
        assert m.is_directed(ei) == bool(gold["is_directed"]), name
        uei, uea = m.undirect_graph(ei, ea)
        assert torch.equal(uei.cpu(), gold["undirected_edge_index"]) and torch.equal(uea.cpu(), gold["undirected_edge_attr"]), name


# ---- one-launch preparation for batches laid out tile by tile (pfn_graph_prep_tiled) ------------------------------------
def _tile_status(g):
    import ctypes as C
    from poweflownet_b200 import _lib
    v = C.c_int32(0)
    _lib.check(_lib.lib().pfn_graph_tile_status(g.ws.data_ptr(), C.byref(v), torch.cuda.current_stream().cuda_stream), "pfn_graph_tile_status")
    return int(v.value)


def _same_arrays(a, b, e):
    for k, va in a.arrays().items():
        vb = b.arrays()[k]
        n = e if k.startswith(("nbr", "eid", "ea")) else va.size(0)
        assert torch.equal(va[:n], vb[:n]), k


@pytest.mark.parametrize("case,b,tile_rows", [("118v2", 128, 118), ("14", 18, 126), ("14", 27, 126), ("14", 5, 14), ("118v2", 3, 118),
                                               ((9, 12), 42, 126), ((128, 384), 3, 128)])
def test_tiled_prep_writes_what_the_general_prep_writes(case, b, tile_rows):
    from poweflownet_b200 import ops
    from poweflownet_b200.data import synthetic_batch
    batch = synthetic_batch(cases=[case] * b, seed=5)
    dev = torch.device("cuda", 0)
    ei, ea = batch.edge_index.to(dev), batch.edge_attr.to(dev)
    general = ops.PreparedGraph(ei, ea, batch.num_nodes, mode=1)
    tiled = ops.PreparedGraph(ei, ea, batch.num_nodes, mode=1, tile_rows=tile_rows)
    assert tiled.tiled and not general.tiled
    assert tiled.meta() == general.meta() == (True, 2 * ei.size(1), 0)
    assert _tile_status(tiled) == 0
    _same_arrays(tiled, general, 2 * ei.size(1))
    uei, uea = O.undirect_graph(batch.edge_index, batch.edge_attr)
    _check_csr(tiled, uei, uea, batch.num_nodes)


def test_tiled_prep_on_an_already_undirected_batch():
    """Both directions listed graph by graph (first edge's reverse present): is_directed is False, nothing is appended."""
    from poweflownet_b200 import ops
    from poweflownet_b200.data import synthetic_batch
    one = synthetic_batch(cases=["14"] * 18, seed=2)
    per = one.edge_index.size(1) // 18
    cols, attrs = [], []
    for g in range(18):
        e = one.edge_index[:, g * per:(g + 1) * per]
        cols += [e, e.flip(0)]
        attrs += [one.edge_attr[g * per:(g + 1) * per]] * 2
    ei, ea = torch.cat(cols, 1).cuda(), torch.cat(attrs, 0).cuda()
    general = ops.PreparedGraph(ei, ea, one.num_nodes, mode=1)
    tiled = ops.PreparedGraph(ei, ea, one.num_nodes, mode=1, tile_rows=126)
    assert tiled.tiled and tiled.meta() == general.meta() == (False, ei.size(1), 0)
    assert _tile_status(tiled) == 0
    _same_arrays(tiled, general, ei.size(1))
    for mode in (0,):  # edges as given
        _same_arrays(ops.PreparedGraph(ei, ea, one.num_nodes, mode=mode, tile_rows=126), ops.PreparedGraph(ei, ea, one.num_nodes, mode=mode), ei.size(1))


def test_tiled_prep_refuses_layouts_it_cannot_place():
    """Columns that are not grouped tile by tile (shuffled edge order), or an edge that leaves its tile, raise the tile flag
    and make the tile's neighbour ids -1 (so the graph-resident kernels poison it); shapes that do not qualify report
    `tiled == False` and take the general path."""
    from poweflownet_b200 import ops
    from poweflownet_b200.data import synthetic_batch
    batch = synthetic_batch(cases=["14"] * 18, seed=5)
    dev = torch.device("cuda", 0)
    ei, ea = batch.edge_index.to(dev), batch.edge_attr.to(dev)
    perm = torch.randperm(ei.size(1), generator=torch.Generator().manual_seed(0)).to(dev)
    shuffled = ops.PreparedGraph(ei[:, perm].contiguous(), ea[perm].contiguous(), batch.num_nodes, mode=1, tile_rows=126)
    assert shuffled.tiled and _tile_status(shuffled) == 1
    assert int(shuffled.arrays()["nbr_t"].min()) == -1
    crossing = ei.clone()
    crossing[:, -1] = torch.tensor([125, 126], device=dev)
    g = ops.PreparedGraph(crossing, ea, batch.num_nodes, mode=1, tile_rows=126)
    assert g.tiled and _tile_status(g) == 1
    nbr = g.arrays()["nbr_t"]
    e_tile = nbr.numel() // 2
    assert int(nbr[:e_tile].min()) >= 0 and bool((nbr[e_tile:] == -1).all())  # tile 0 is fine, tile 1 holds the stray edge
    # not a multiple of the tile count / too many rows per tile / no edges: the general path
    assert not ops.PreparedGraph(ei[:, :-1].contiguous(), ea[:-1].contiguous(), batch.num_nodes, mode=1, tile_rows=126).tiled
    assert not ops.PreparedGraph(ei, ea, batch.num_nodes, mode=1, tile_rows=252).tiled
    assert not ops.PreparedGraph(ei[:, :0].contiguous(), ea[:0].contiguous(), batch.num_nodes, mode=1, tile_rows=126).tiled


def test_tiled_prep_flags_out_of_range_ids():
    from poweflownet_b200 import ops
    from poweflownet_b200.data import synthetic_batch
    batch = synthetic_batch(cases=["14"] * 18, seed=5)
    ei = batch.edge_index.cuda().clone()
    ei[1, 3] = batch.num_nodes + 5
    g = ops.PreparedGraph(ei, batch.edge_attr.cuda(), batch.num_nodes, mode=1, tile_rows=126)
    assert g.tiled and g.meta()[2] == 1 and _tile_status(g) == 1
