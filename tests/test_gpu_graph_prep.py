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
