"""Each CUDA kernel family on the B200 through the C ABI against plain fp32/fp64 torch on the CPU:
dense Linear forward / data-gradient / weight-gradient (all epilogues), the fused EdgeAggregation
message+aggregate kernel and its backward, the TAGConv hop (forward and transposed), fused MSE.
Tolerance: 1e-5 relative (max-abs / max-abs and Frobenius), BASELINE.json north_star."""
import pytest
import torch

import common
from emulation import seg_sum

pytestmark = pytest.mark.gpu
TOL = 1e-5
DEV = "cuda:0"


def _rows(t, ld=None):
    """copy a [n, w] CPU matrix into a padded device matrix (ld = round_up4(w)) and return the [:, :w] view"""
    from poweflownet_b200 import ops
    buf = ops.new_rows(t.size(0), t.size(1), DEV)
    buf[:, :t.size(1)].copy_(t)
    return buf


def _assert_close(got, want, tol=TOL, what=""):
    e = common.rel_err(got.cpu(), want)
    assert max(e) < tol, (what, e)


@pytest.mark.parametrize("m,n_in,n_out", [(1000, 129, 129), (777, 4, 129), (515, 129, 4), (300, 64, 64), (130, 33, 70),
                                          (15104, 129, 129), (64, 512, 512)])
def test_linear_forward_epilogues(m, n_in, n_out):
    from poweflownet_b200 import ops
    g = torch.Generator().manual_seed(m + n_in)
    x, w, b = torch.randn(m, n_in, generator=g), torch.randn(n_out, n_in, generator=g) / n_in ** 0.5, torch.randn(n_out, generator=g)
    rs, add = torch.rand(m, generator=g) * 3, torch.randn(m, n_out, generator=g)
    keep = (torch.rand(m, n_out, generator=g) < 0.8).float()
    xd, wd, bd = _rows(x), w.to(DEV), b.to(DEV)
    ref = x.double() @ w.double().T
    out = ops.new_rows(m, n_out, DEV)
    ops.linear_fwd(xd, wd, n_in, n_in, n_out, None, out)
    _assert_close(out[:, :n_out], ref, what="plain")
    ops.linear_fwd(xd, wd, n_in, n_in, n_out, bd, out, rowscale=rs.to(DEV), addend=_rows(add), act=ops.ACT_RELU)
    _assert_close(out[:, :n_out], torch.relu(ref + rs.double()[:, None] * b.double() + add.double()), what="bias*rowscale+add+relu")
    ops.linear_fwd(xd, wd, n_in, n_in, n_out, bd, out, act=ops.ACT_DROPOUT_RELU, p=0.2, inj_mask=keep.to(DEV))
    _assert_close(out[:, :n_out], torch.relu((ref + b.double()) * keep.double() / 0.8), what="injected dropout+relu")


def test_dropout_generator_statistics_and_determinism():
    from poweflownet_b200 import ops
    m, n = 4096, 129
    x, w = torch.ones(m, 1), torch.ones(n, 1)
    a, b, c = (ops.new_rows(m, n, DEV) for _ in range(3))
    ops.linear_fwd(_rows(x), w.to(DEV), 1, 1, n, None, a, act=ops.ACT_DROPOUT_RELU, p=0.2, seed=1234)
    ops.linear_fwd(_rows(x), w.to(DEV), 1, 1, n, None, b, act=ops.ACT_DROPOUT_RELU, p=0.2, seed=1234)
    ops.linear_fwd(_rows(x), w.to(DEV), 1, 1, n, None, c, act=ops.ACT_DROPOUT_RELU, p=0.2, seed=99)
    a, b, c = a[:, :n], b[:, :n], c[:, :n]
    assert torch.equal(a, b) and not torch.equal(a, c)
    vals = torch.unique(a).cpu()
    assert torch.allclose(vals, torch.tensor([0.0, 1.25]))  # kept entries scaled by 1/(1-p)
    keep = float((a > 0).float().mean())
    assert abs(keep - 0.8) < 0.005
    assert float(((a > 0) & (c > 0)).float().mean()) == pytest.approx(0.64, abs=0.01)  # independent across seeds
    assert float((a > 0).float().mean(0).std()) < 0.02 and float((a > 0).float().mean(1).std()) < 0.06


@pytest.mark.parametrize("m,n_in,n_out", [(1000, 129, 129), (777, 4, 129), (515, 129, 4), (15104, 129, 129), (3000, 512, 64)])
def test_linear_dgrad_wgrad(m, n_in, n_out):
    from poweflownet_b200 import ops
    g = torch.Generator().manual_seed(7 * m + n_out)
    x, w, dy = torch.randn(m, n_in, generator=g), torch.randn(n_out, n_in, generator=g), torch.randn(m, n_out, generator=g)
    ym, rs = torch.randn(m, n_in, generator=g), torch.rand(m, generator=g) * 4
    dx = ops.new_rows(m, n_in, DEV)
    ops.linear_dgrad(_rows(dy), w.to(DEV), n_in, n_in, n_out, dx)
    _assert_close(dx[:, :n_in], dy.double() @ w.double(), what="dgrad")
    ops.linear_dgrad(_rows(dy), w.to(DEV), n_in, n_in, n_out, dx, ymask=_rows(ym), scale=1.25)
    _assert_close(dx[:, :n_in], (dy.double() @ w.double()) * (ym > 0).double() * 1.25, what="dgrad masked")
    dw, db = torch.empty(n_out, n_in, device=DEV), torch.empty(n_out, device=DEV)
    ops.linear_wgrad(_rows(dy), _rows(x), n_in, n_out, dw, n_in, dbias=db)
    _assert_close(dw, dy.double().T @ x.double(), what="wgrad")
    _assert_close(db, dy.double().sum(0), what="bias grad")
    ops.linear_wgrad(_rows(dy), _rows(x), n_in, n_out, dw, n_in, dbias=db, rowscale=rs.to(DEV))
    _assert_close(db, (rs.double()[:, None] * dy.double()).sum(0), what="row-scaled bias grad")
    # twice the same launch => bitwise identical (fixed-order split-K reduction, no atomics)
    dw2 = torch.empty_like(dw)
    ops.linear_wgrad(_rows(dy), _rows(x), n_in, n_out, dw2, n_in)
    ops.linear_wgrad(_rows(dy), _rows(x), n_in, n_out, dw, n_in)
    assert torch.equal(dw, dw2)


def _edge_case(name, h, seed=0):
    from poweflownet_b200 import ops
    from oracle import pfn_oracle as O
    batch = common.make_batch(name) if isinstance(name, str) else name
    n = batch.num_nodes
    ei, ea = O.undirect_graph(batch.edge_index, batch.edge_attr)
    g = torch.Generator().manual_seed(seed)
    hi, hj, ds = (torch.randn(n, h, generator=g) for _ in range(3))
    fin = 3
    w1 = torch.randn(h, 2 * fin + 2, generator=g)
    graph = ops.PreparedGraph(batch.edge_index.to(DEV), batch.edge_attr.to(DEV), n, mode=1)
    return n, ei, ea, hi, hj, ds, w1, fin, graph


@pytest.mark.parametrize("name,h", [("tiny", 8), ("isolated_and_parallel", 20), ("case118_h33", 33), ("mixed", 129),
                                    ("no_edges", 8), ("case14_small", 64), ("mixed", 512), ("tiny", 1100)])
def test_ea_message_aggregate_forward_backward(name, h):
    from poweflownet_b200 import ops
    n, ei, ea, hi, hj, ds, w1, fin, graph = _edge_case(name, h)
    src, tgt = ei[0], ei[1]
    we = w1[:, 2 * fin:].double()
    pre = hi.double()[tgt] + hj.double()[src] + ea.double() @ we.T
    s_ref = seg_sum(torch.relu(pre), tgt, n)
    hid, hjd, s = _rows(hi), _rows(hj), ops.new_rows(n, h, DEV)
    ops.ea_fwd(hid, hjd, graph, w1.to(DEV), fin, h, s)
    _assert_close(s[:, :h], s_ref, what="S")
    ge = ds.double()[tgt] * (pre > 0).double()
    dhi, dhj, dw1 = ops.new_rows(n, h, DEV), ops.new_rows(n, h, DEV), torch.zeros(h, 2 * fin + 2, device=DEV)
    ops.ea_bwd(_rows(ds), hid, hjd, graph, w1.to(DEV), fin, h, dhi, dhj, dw1)
    _assert_close(dhi[:, :h], seg_sum(ge, tgt, n), what="dHi")
    _assert_close(dhj[:, :h], seg_sum(ge, src, n), what="dHj")
    _assert_close(dw1[:, 2 * fin:], ge.T @ ea.double(), what="dWe")
    assert float(dw1[:, :2 * fin].abs().max()) == 0.0  # only the We column block is written


def test_ea_forward_full_size_case118_b128():
    from poweflownet_b200 import ops
    from poweflownet_b200.data import synthetic_batch
    batch = synthetic_batch("118v2", 128)
    n, ei, ea, hi, hj, ds, w1, fin, graph = _edge_case(batch, 129, seed=5)
    src, tgt = ei[0], ei[1]
    pre = hi[tgt] + hj[src] + ea @ w1[:, 2 * fin:].T
    s = ops.new_rows(n, 129, DEV)
    ops.ea_fwd(_rows(hi), _rows(hj), graph, w1.to(DEV), fin, 129, s)
    _assert_close(s[:, :129], seg_sum(torch.relu(pre).double(), tgt, n), what="S full size")
    s2 = ops.new_rows(n, 129, DEV)
    ops.ea_fwd(_rows(hi), _rows(hj), graph, w1.to(DEV), fin, 129, s2)
    assert torch.equal(s, s2)  # deterministic: one thread per row chunk, fixed edge order


@pytest.mark.parametrize("name,h", [("tiny", 8), ("isolated_and_parallel", 20), ("case118_h33", 33), ("mixed", 129), ("no_edges", 8)])
def test_tag_hop_forward_and_transposed(name, h):
    from poweflownet_b200 import ops
    n, ei, ea, x, add, ym, _, _, graph = _edge_case(name, h, seed=3)
    src, tgt = ei[0], ei[1]
    deg = torch.bincount(tgt, minlength=n).double()
    dis = torch.where(deg > 0, deg.pow(-0.5), torch.zeros_like(deg))
    w = (dis[src] * dis[tgt])[:, None]
    y = ops.new_rows(n, h, DEV)
    ops.spmm_hop(_rows(x), graph, y, h)
    _assert_close(y[:, :h], seg_sum(w * x.double()[src], tgt, n), what="A_hat x")
    ops.spmm_hop(_rows(x), graph, y, h, transpose=True)
    _assert_close(y[:, :h], seg_sum(w * x.double()[tgt], src, n), what="A_hat^T x")
    acc = _rows(add)
    ops.spmm_hop(_rows(x), graph, acc, h, transpose=True, addend=acc, ymask=_rows(ym), scale=1.25)
    want = (seg_sum(w * x.double()[tgt], src, n) + add.double()) * (ym > 0).double() * 1.25
    _assert_close(acc[:, :h], want, what="in-place add + mask")


def test_fused_mse():
    from poweflownet_b200.training import mse_loss_and_grad
    g = torch.Generator().manual_seed(1)
    out, y = torch.randn(15104, 4, generator=g), torch.randn(15104, 4, generator=g)
    loss, dout = mse_loss_and_grad(out.to(DEV), y.to(DEV))
    ref = ((out.double() - y.double()) ** 2).mean()
    assert abs(float(loss) - float(ref)) < 1e-6 * float(ref)
    _assert_close(dout, 2 * (out.double() - y.double()) / out.numel())
    loss2, _ = mse_loss_and_grad(out.to(DEV), y.to(DEV), total_count=2 * out.numel())
    assert abs(float(loss2) - float(ref) / 2) < 1e-6 * float(ref)


def _star_batch(spokes=70):
    """A hub bus with `spokes` branches (70: several rounds of gathered rows in one batch; 300 / 1200: more than a batch
    buffer of the bulk-copy kernel holds, so the hub row is summed from global memory -- 1200 also exceeds the producer's
    window of neighbour ids), a short chain and one isolated bus."""
    n = spokes + 5
    hub = torch.stack([torch.zeros(spokes, dtype=torch.long), torch.arange(1, spokes + 1)])
    chain = torch.tensor([[spokes + 1, spokes + 2], [spokes + 2, spokes + 3]])
    ei = torch.cat([hub, chain], dim=1)
    g = torch.Generator().manual_seed(123)
    z = torch.zeros(n, 4)
    return common.GraphBatch(x=z, y=z, bus_type=torch.zeros(n, dtype=torch.long), pred_mask=torch.zeros(n, 4, dtype=torch.long),
                             edge_index=ei.contiguous(), edge_attr=torch.randn(ei.size(1), 2, generator=g),
                             batch=torch.zeros(n, dtype=torch.long), ptr=torch.tensor([0, n]))


def _tma_vs_cta(batch, h, monkeypatch, env=None, seed=11):
    from poweflownet_b200 import ops
    n, ei, ea, hi, hj, ds, w1, fin, graph = _edge_case(batch, h, seed=seed)
    src, tgt = ei[0], ei[1]
    pre = hi.double()[tgt] + hj.double()[src] + ea.double() @ w1[:, 2 * fin:].double().T
    s_ref = seg_sum(torch.relu(pre), tgt, n)
    outs = {}
    for which in ("cta", "tma"):
        monkeypatch.setenv("PFN_EA_FWD", which)
        for k, v in (env or {}).items():
            monkeypatch.setenv(k, v)
        s = ops.new_rows(n, h, DEV)
        s.fill_(float("nan"))
        ops.ea_fwd(_rows(hi), _rows(hj), graph, w1.to(DEV), fin, h, s)
        outs[which] = s
        _assert_close(s[:, :h], s_ref, what=f"S ({which})")
    assert torch.equal(outs["cta"][:, :h], outs["tma"][:, :h])  # same FMAs, same ascending-edge summation order


@pytest.mark.parametrize("h", [8, 33, 64, 93, 128, 129, 132, 256, 512, 1100, 3300])
@pytest.mark.parametrize("name", ["mixed", "star", "bigstar", "hugestar", "isolated_and_parallel", "case118_h33", "no_edges"])
def test_ea_forward_bulk_copy_kernel_equals_cta_kernel(name, h, monkeypatch):
    """`k_ea_fwd_tma` (a producer warp feeding shared-memory batch buffers with bulk copies, ~24 consumer warps) against
    the double-precision reference and, bit for bit, against the CTA-slab kernel -- narrow and wide rows, hub buses
    inside a batch (star) and beyond a batch buffer / the neighbour window (bigstar, hugestar), isolated buses, parallel
    branches, an empty edge list."""
    batch = {"star": lambda: _star_batch(), "bigstar": lambda: _star_batch(300), "hugestar": lambda: _star_batch(1200)}.get(name, lambda: name)()
    _tma_vs_cta(batch, h, monkeypatch)


@pytest.mark.parametrize("env", [{"PFN_EA_STAGES": "2"}, {"PFN_EA_STAGES": "3", "PFN_EA_THREADS": "256"}, {"PFN_EA_THREADS": "992"},
                                 {"PFN_EA_THREADS": "64", "PFN_EA_PRODUCERS": "1"}, {"PFN_EA_PRODUCERS": "7", "PFN_EA_THREADS": "800"}, {"PFN_EA_PRODUCERS": "12"}, {"PFN_EA_PRODUCERS": "16", "PFN_EA_THREADS": "256"},
                                 {"PFN_EA_BULK": "1"}, {"PFN_EA_BULK": "0", "PFN_EA_PRODUCERS": "2"}, {"PFN_EA_PREFETCH": "1"},
                                 {"PFN_EA_BULK8": "3"}, {"PFN_EA_BULK8": "5", "PFN_EA_STAGES": "4"}, {"PFN_EA_BULK8": "7", "PFN_EA_PRODUCERS": "3"},
                                 {"PFN_EA_CHUNK": "8", "PFN_EA_BULK8": "4"}, {"PFN_EA_ROUND": "1"}, {"PFN_EA_CTAS_PER_SM": "1"},
                                 {"PFN_EA_CTAS_PER_SM": "2", "PFN_EA_STAGES": "1"}, {"PFN_EA_CTAS_PER_SM": "2", "PFN_EA_THREADS": "64", "PFN_EA_PRODUCERS": "1"}])
@pytest.mark.parametrize("name,h", [("mixed", 129), ("star", 64), ("case118_h33", 512)])
def test_ea_forward_bulk_copy_kernel_knobs(name, h, env, monkeypatch):
    """Every ring geometry / copy mechanism the host can pick gives the same bits (2-4 stages, 2 to 31 consumer warps, 1 to
    7 producer warps, gathered rows as bulk copies or as cp.async chunks, L2 prefetch)."""
    _tma_vs_cta(_star_batch() if name == "star" else name, h, monkeypatch, env)


def test_ea_forward_bulk_copy_full_size_and_strided(monkeypatch):
    """Bench size (case118v2 x 128, hidden 129), 6470-bus graphs at hidden 512 and a batch whose CTA row ranges exceed the
    staged row-pointer slice, plus operands that are column blocks of wider matrices (rows not contiguous: per-row copies)."""
    from poweflownet_b200 import ops
    from poweflownet_b200.data import synthetic_batch
    _tma_vs_cta(synthetic_batch("118v2", 128), 129, monkeypatch, seed=5)
    _tma_vs_cta(synthetic_batch("6470rte", 2), 512, monkeypatch, seed=6)
    _tma_vs_cta(synthetic_batch("6470rte", 50), 8, monkeypatch, seed=7)  # 323,500 rows: 2,186 per CTA > 2,048 staged row pointers
    for h in (129, 256):
        n, ei, ea, hi, hj, ds, w1, fin, graph = _edge_case("mixed", h, seed=12)
        ld = ops.round_up4(h)
        wide_i, wide_j = torch.zeros(n, 3 * ld, device=DEV), torch.zeros(n, 3 * ld, device=DEV)
        wide_i[:, ld:ld + h].copy_(hi)
        wide_j[:, 2 * ld:2 * ld + h].copy_(hj)
        outs = {}
        for which in ("cta", "tma"):
            monkeypatch.setenv("PFN_EA_FWD", which)
            s = torch.zeros(n, 2 * ld, device=DEV)
            from poweflownet_b200._lib import check, lib
            check(lib().pfn_ea_fwd(wide_i[:, ld:].data_ptr(), wide_j[:, ld:].data_ptr() + 4 * ld, 3 * ld, graph.ws.data_ptr(), n, graph.e_raw,
                                   w1.to(DEV).data_ptr() + 4 * 2 * fin, 2 * fin + 2, s[:, ld:].data_ptr(), 2 * ld, h,
                                   torch.cuda.current_stream().cuda_stream), "pfn_ea_fwd")
            outs[which] = s
        assert torch.equal(outs["cta"], outs["tma"]) and float(outs["tma"][:, :ld].abs().max()) == 0.0

