"""The oracle restatement (`oracle/pfn_oracle.py`) against the golden vectors produced by the
reference's own `networks/MPN.py` (tests/golden/make_golden.py), plus an independent loop-level
restatement of the graph arithmetic for the tiny cases.  CPU only."""
import numpy as np
import pytest
import torch

import common
from oracle import pfn_oracle as O

ALL = list(common.CASES)


def _load(name):
    return torch.load(common.golden_path(name), weights_only=True)


def _batch(gold):
    return common.GraphBatch(**gold["inputs"])


@pytest.mark.parametrize("name", ALL)
def test_inputs_reproducible(name):
    gold = _load(name)
    b = common.make_batch(name)
    for k, v in gold["inputs"].items():
        assert torch.equal(getattr(b, k), v), k


@pytest.mark.parametrize("name", ALL)
def test_integer_graph_work_bit_exact(name):
    gold = _load(name)
    b = _batch(gold)
    assert O.is_directed(b.edge_index) == bool(gold["is_directed"])
    ei, ea = O.undirect_graph(b.edge_index, b.edge_attr)
    assert ei.dtype == torch.int64 and torch.equal(ei, gold["undirected_edge_index"])
    assert torch.equal(ea, gold["undirected_edge_attr"])


@pytest.mark.parametrize("name", ALL)
def test_state_dict_layout_matches_reference(name):
    gold = _load(name)
    model = O.MaskEmbdMultiMPN(**gold["meta"]["model_kwargs"])
    ref_keys = gold["grads"].keys() if "grads" in gold else gold["grad_norm"].keys()
    assert list(model.state_dict().keys()) == list(ref_keys)
    if "grads" in gold:
        for k, p in model.named_parameters():
            assert tuple(p.shape) == tuple(gold["grads"][k].shape), k


def test_param_count_standard():
    m = O.MaskEmbdMultiMPN(4, 2, 4, 129, 4, 3, 0.2)
    assert sum(p.numel() for p in m.parameters()) == 354500
    assert len(m.state_dict()) == 35


@pytest.mark.parametrize("name", ALL)
def test_eval_forward_matches_reference(name):
    gold = _load(name)
    torch.set_num_threads(1)
    model = common.load_seeded(O.MaskEmbdMultiMPN(**gold["meta"]["model_kwargs"])).eval()
    with torch.no_grad():
        out = model(_batch(gold))
    assert torch.equal(out, gold["eval_out"]) or max(common.rel_err(out, gold["eval_out"])) < 1e-6
    m64 = common.load_seeded(O.MaskEmbdMultiMPN(**gold["meta"]["model_kwargs"])).double().eval()
    b = _batch(gold)
    b64 = common.GraphBatch(**{k: (v.double() if v.is_floating_point() else v) for k, v in gold["inputs"].items()})
    with torch.no_grad():
        out64 = m64(b64)
    assert max(common.rel_err(out64, gold["eval_out_fp64"])) < 1e-12
    # error budget: fp32 reference vs its fp64 twin
    assert max(common.rel_err(gold["eval_out"], gold["eval_out_fp64"])) < 1e-5


@pytest.mark.parametrize("name", ALL)
def test_train_step_matches_reference(name):
    gold = _load(name)
    torch.set_num_threads(1)
    kw = gold["meta"]["model_kwargs"]
    model = common.load_seeded(O.MaskEmbdMultiMPN(**kw)).train()
    b = _batch(gold)
    masks = common.dropout_masks(name, b.num_nodes)
    loss, out = O.forward_loss_backward(model, b, "mse", dropout_masks=masks)
    assert max(common.rel_err(out, gold["train_out"])) < 1e-6
    assert abs(float(loss) - float(gold["train_loss"])) <= 1e-6 * abs(float(gold["train_loss"]))
    for k, p in model.named_parameters():
        if "grads" in gold:
            assert max(common.rel_err(p.grad, gold["grads"][k])) < 2e-6, k
        else:
            assert abs(float(p.grad.double().norm()) - float(gold["grad_norm"][k])) <= 2e-6 * float(gold["grad_norm"][k]) + 1e-12, k
            s = p.grad.reshape(-1)
            step = max(1, s.numel() // 257)
            assert max(common.rel_err(s[::step][:257], gold["grad_sample"][k])) < 2e-6, k


# ------------------------------------------------------------------------------------------
# independent restatement with explicit loops (numpy, fp64) of the two graph operators
# ------------------------------------------------------------------------------------------
def _loops_edge_aggregation(x, ei, ea, w1, b1, w2, b2):
    n = x.shape[0]
    out = np.zeros((n, w2.shape[0]))
    for e in range(ei.shape[1]):
        j, i = ei[0, e], ei[1, e]  # source j -> target i
        z = np.concatenate([x[i], x[j], ea[e]])
        out[i] += w2 @ np.maximum(w1 @ z + b1, 0.0) + b2
    return out


def _loops_tag(x, ei, ws, bias):
    n = x.shape[0]
    deg = np.zeros(n)
    for e in range(ei.shape[1]):
        deg[ei[1, e]] += 1
    dis = np.where(deg > 0, 1.0 / np.sqrt(np.maximum(deg, 1e-300)), 0.0)
    out = x @ ws[0].T
    for wk in ws[1:]:
        nx = np.zeros_like(x)
        for e in range(ei.shape[1]):
            j, i = ei[0, e], ei[1, e]
            nx[i] += dis[j] * dis[i] * x[j]
        x = nx
        out = out + x @ wk.T
    return out + bias


@pytest.mark.parametrize("name", ["tiny", "isolated_and_parallel", "already_undirected", "no_edges"])
def test_operators_against_explicit_loops(name):
    gold = _load(name)
    b = _batch(gold)
    kw = gold["meta"]["model_kwargs"]
    model = common.load_seeded(O.MaskEmbdMultiMPN(**kw)).double().eval()
    ei, ea = O.undirect_graph(b.edge_index, b.edge_attr.double())
    g = torch.Generator().manual_seed(3)
    x = torch.randn(b.num_nodes, kw["hidden_dim"], generator=g, dtype=torch.double)
    ea_layer, tag_layer = model.layers[2], model.layers[1]
    with torch.no_grad():
        got_ea = ea_layer(x, ei, ea).numpy()
        got_tag = tag_layer(x, ei).numpy()
    l0, l2 = ea_layer.edge_aggr[0], ea_layer.edge_aggr[2]
    want_ea = _loops_edge_aggregation(x.numpy(), ei.numpy(), ea.numpy(), l0.weight.detach().numpy(), l0.bias.detach().numpy(),
                                      l2.weight.detach().numpy(), l2.bias.detach().numpy())
    want_tag = _loops_tag(x.numpy(), ei.numpy(), [l.weight.detach().numpy() for l in tag_layer.lins], tag_layer.bias.detach().numpy())
    np.testing.assert_allclose(got_ea, want_ea, rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(got_tag, want_tag, rtol=1e-11, atol=1e-11)
