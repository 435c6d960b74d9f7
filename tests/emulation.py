"""CPU emulation (plain torch, test-only) of the ALGORITHM libpfn_b200's engine runs, step for step:
per-node Hi/Hj GEMMs, segmented gather/ReLU/sum, second Linear hoisted after the sum, K-segmented
TAGConv GEMM, and the hand-derived backward with the ReLU mask recomputed from Hi/Hj.

It exists to check the algebra of csrc/engine.cu against the oracle's autograd on a box without a GPU
(tests/test_emulation.py); the kernels themselves are checked against the oracle on the B200.
"""
from __future__ import annotations

import torch


def stable_csr(edge_index, n, by_target=True):
    key = edge_index[1] if by_target else edge_index[0]
    order = torch.sort(key, stable=True).indices
    rowptr = torch.zeros(n + 1, dtype=torch.long)
    rowptr[1:] = torch.bincount(key, minlength=n).cumsum(0)
    nbr = (edge_index[0] if by_target else edge_index[1])[order]
    return rowptr, nbr, order


def seg_sum(per_edge, owner, n):
    out = torch.zeros((n, per_edge.size(1)), dtype=per_edge.dtype)
    return out.index_add_(0, owner, per_edge)


class EngineEmulation:
    def __init__(self, state_dict, kw, dtype=torch.float32):
        self.sd = {k: v.to(dtype).clone() for k, v in state_dict.items()}
        self.kw, self.dtype = kw, dtype
        L = kw["n_gnn_layers"]
        self.kinds = ["ea", "tag"] + ["ea", "tag"] * (L - 2) + ["ea"]

    def forward(self, batch, undirected, training, masks=None):
        sd, kw, dt = self.sd, self.kw, self.dtype
        ei, ea = undirected
        ea = ea.to(dt)
        n = batch.x.size(0)
        p = kw["dropout_rate"] if training else 0.0
        scale = 1.0 / (1.0 - p)
        src, tgt = ei[0], ei[1]
        deg = torch.bincount(tgt, minlength=n).to(dt)
        dis = torch.where(deg > 0, deg.pow(-0.5), torch.zeros_like(deg))
        saved = {"n": n, "ei": ei, "ea": ea, "deg": deg, "dis": dis, "scale": scale, "layers": []}
        maskf = batch.pred_mask.to(dt)
        t1 = torch.relu(maskf @ sd["mask_embd.0.weight"].T + sd["mask_embd.0.bias"])
        cur = t1 @ sd["mask_embd.2.weight"].T + sd["mask_embd.2.bias"] + batch.x.to(dt)
        saved.update(maskf=maskf, t1=t1, x0=cur)
        for li, kind in enumerate(self.kinds):
            last = li == len(self.kinds) - 1
            rec = {"in": cur}
            if kind == "ea":
                w1, b1 = sd[f"layers.{li}.edge_aggr.0.weight"], sd[f"layers.{li}.edge_aggr.0.bias"]
                w2, b2 = sd[f"layers.{li}.edge_aggr.2.weight"], sd[f"layers.{li}.edge_aggr.2.bias"]
                fin = cur.size(1)
                wi, wj, we = w1[:, :fin], w1[:, fin:2 * fin], w1[:, 2 * fin:]
                hi, hj = cur @ wi.T + b1, cur @ wj.T
                pre = hi[tgt] + hj[src] + ea @ we.T
                s = seg_sum(torch.relu(pre), tgt, n)
                z = s @ w2.T + deg[:, None] * b2
                rec.update(hi=hi, hj=hj, s=s)
            else:
                K = kw["K"]
                xs = [cur]
                for _ in range(K):
                    xs.append(dis[:, None] * seg_sum(dis[src][:, None] * xs[-1][src], tgt, n))
                z = sum(xs[k] @ sd[f"layers.{li}.lins.{k}.weight"].T for k in range(K + 1)) + sd[f"layers.{li}.bias"]
                rec.update(xs=xs)
            if not last:
                if training and masks is not None:
                    z = z * masks[li].to(dt) * scale
                cur = torch.relu(z)
            else:
                cur = z
            rec["out"] = cur
            saved["layers"].append(rec)
        self.saved = saved
        return cur

    def backward(self, dout):
        sd, kw, sv = self.sd, self.kw, self.saved
        n, ei, ea, deg, dis, scale = sv["n"], sv["ei"], sv["ea"], sv["deg"], sv["dis"], sv["scale"]
        src, tgt = ei[0], ei[1]
        grads = {}
        g = dout.to(self.dtype)
        for li in range(len(self.kinds) - 1, -1, -1):
            rec, kind = sv["layers"][li], self.kinds[li]
            cur = rec["in"]
            cur_has_act = li > 0
            if kind == "ea":
                w1 = sd[f"layers.{li}.edge_aggr.0.weight"]
                w2 = sd[f"layers.{li}.edge_aggr.2.weight"]
                fin = cur.size(1)
                wi, wj, we = w1[:, :fin], w1[:, fin:2 * fin], w1[:, 2 * fin:]
                grads[f"layers.{li}.edge_aggr.2.weight"] = g.T @ rec["s"]
                grads[f"layers.{li}.edge_aggr.2.bias"] = (deg[:, None] * g).sum(0)
                ds = g @ w2
                pre = rec["hi"][tgt] + rec["hj"][src] + ea @ we.T  # mask recomputed, not stored
                ge = ds[tgt] * (pre > 0).to(ds.dtype)
                dhi, dhj = seg_sum(ge, tgt, n), seg_sum(ge, src, n)
                dwe = ge.T @ ea
                grads[f"layers.{li}.edge_aggr.0.weight"] = torch.cat([dhi.T @ cur, dhj.T @ cur, dwe], dim=1)
                grads[f"layers.{li}.edge_aggr.0.bias"] = dhi.sum(0)
                g = dhi @ wi + dhj @ wj
            else:
                K = kw["K"]
                xs = rec["xs"]
                for k in range(K + 1):
                    grads[f"layers.{li}.lins.{k}.weight"] = g.T @ xs[k]
                grads[f"layers.{li}.bias"] = g.sum(0)
                dxs = [g @ sd[f"layers.{li}.lins.{k}.weight"] for k in range(K + 1)]
                for k in range(K, 0, -1):  # A_hat^T: walk the CSR by source
                    dxs[k - 1] = dxs[k - 1] + dis[:, None] * seg_sum(dis[tgt][:, None] * dxs[k][tgt], src, n)
                g = dxs[0]
            if cur_has_act:
                g = g * (cur > 0).to(g.dtype) * scale
        grads["mask_embd.2.weight"] = g.T @ sv["t1"]
        grads["mask_embd.2.bias"] = g.sum(0)
        dt1 = (g @ sd["mask_embd.2.weight"]) * (sv["t1"] > 0).to(g.dtype)
        grads["mask_embd.0.weight"] = dt1.T @ sv["maskf"]
        grads["mask_embd.0.bias"] = dt1.sum(0)
        self.dx0 = g
        return grads


def power_imbalance_emulation(x, undirected, stats):
    """The arithmetic of csrc/loss_optim.cu (k_pi_node / k_pi_final / k_pi_grad) in plain torch: the loss and the
    HAND-DERIVED gradient w.r.t. the normalised predictions.  `undirected` = the (doubled) branch list the loss works
    on; `stats` = (xymean, xystd, edgemean, edgestd).  Returns (loss, dx [N, 4])."""
    import math
    xm, xs, em, es = [s.reshape(-1).to(x.dtype) for s in stats]
    ei, ea = undirected
    n = x.size(0)
    k = 1 / 180. * math.pi
    vm = x[:, 0] * xs[0] + xm[0]
    va = k * (x[:, 1] * xs[1] + xm[1])
    c, s = torch.cos(va), torch.sin(va)
    e, f = vm * c, vm * s
    r, xx = ea[:, 0].to(x.dtype) * es[0] + em[0], ea[:, 1].to(x.dtype) * es[1] + em[1]
    den = r * r + xx * xx
    g, b = r / den, -xx / den
    i, j = ei[0], ei[1]
    ei_, fi_, ej_, fj_ = e[i], f[i], e[j], f[j]
    cross = fi_ * ej_ - ei_ * fj_
    P = g * (ei_ * ej_ - ei_ * ei_ + fi_ * fj_ - fi_ * fi_) + b * cross
    Q = g * cross + b * (-ei_ * ej_ + ei_ * ei_ - fi_ * fj_ + fi_ * fi_)
    agg = seg_sum(torch.stack([P, Q], dim=1), i, n)
    dp = -agg[:, 0] + (x[:, 2] * xs[2] + xm[2])
    dq = -agg[:, 1] + (x[:, 3] * xs[3] + xm[3])
    loss = (dp * dp + dq * dq).sum() / n
    # k_pi_grad: a = d loss / d aggregated of the aggregating bus of each branch
    ap, aq = (-2.0 * dp / n)[i], (-2.0 * dq / n)[i]
    p_e, p_f = g * (ej_ - 2 * ei_) - b * fj_, g * (fj_ - 2 * fi_) + b * ej_
    q_e, q_f = -g * fj_ + b * (2 * ei_ - ej_), g * ej_ + b * (2 * fi_ - fj_)
    own = torch.stack([ap * p_e + aq * q_e, ap * p_f + aq * q_f], dim=1)          # row of the CSR by source
    pj_e, pj_f = g * ei_ + b * fi_, g * fi_ - b * ei_
    far = torch.stack([ap * pj_e + aq * pj_f, ap * pj_f - aq * pj_e], dim=1)        # row of the CSR by target
    d = seg_sum(own, i, n) + seg_sum(far, j, n)
    de, df = d[:, 0], d[:, 1]
    dvm, dva = de * c + df * s, -de * f + df * e
    dx = torch.stack([dvm * xs[0], dva * k * xs[1], 2.0 * dp / n * xs[2], 2.0 * dq / n * xs[3]], dim=1)
    return loss, dx
