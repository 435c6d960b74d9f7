"""Tensor-level wrappers over the C ABI (include/pfn_b200.h) and the autograd Functions of the
stand-alone layers.  torch is used for device memory and stream handles only; every arithmetic
step below is a kernel of libpfn_b200.so.  CUDA tensors only -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from ._lib import GraphLayout, check, lib

ACT_NONE, ACT_RELU, ACT_DROPOUT_RELU = 0, 1, 2


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError(
                "poweflownet_b200 runs on CUDA tensors only (hand-written sm_100a kernels; no CPU fallback). "
                f"Got a tensor on {t.device}.")
        dev = dev or t.device
        if t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} vs {t.device}")
    return dev


def round_up4(v: int) -> int:
    return (v + 3) // 4 * 4


def new_rows(n: int, width: int, device) -> torch.Tensor:
    """[n, round_up4(width)] fp32 node-feature matrix (16-byte rows) in the library's layout."""
    return torch.zeros((n, round_up4(width)), dtype=torch.float32, device=device)


# ------------------------------------------------------------------------------------------------
# graph preparation
# ------------------------------------------------------------------------------------------------
class PreparedGraph:
    """Device-side CSR (by target and by source), degrees and deg^-1/2 of one mini-batch, built by
    `pfn_graph_prep` without a host round trip (replaces networks/MPN.py:498-523 + PyG gather/scatter
    bookkeeping).  `mode=1` applies the reference's undirect rule, `mode=0` takes the edges as given.
    `tile_rows > 0`: the batch is promised to be laid out tile by tile (equal-sized small graphs as PyG's loader
    collates them) and ONE launch (`pfn_graph_prep_tiled`) builds the same arrays; `self.tiled` tells whether that path
    ran (the shape may not qualify).  The promise is validated on the device and reported by `pfn_graph_tile_status`."""

    def __init__(self, edge_index: torch.Tensor, edge_attr: torch.Tensor, n_nodes: int, mode: int,
                 workspace: Optional[torch.Tensor] = None, tile_rows: int = 0):
        dev = require_cuda(edge_index, edge_attr)
        if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.size(0) != 2:
            raise ValueError("edge_index must be int64 [2, E]")
        e_raw = int(edge_index.size(1))
        if edge_attr.dtype != torch.float32 or edge_attr.dim() != 2 or edge_attr.size(0) != e_raw or edge_attr.size(1) != 2:
            raise NotImplementedError(
                f"edge_attr must be float32 [E, 2] (efeature_dim=2, the dataset's edge width); got {tuple(edge_attr.shape)}")
        if edge_index.stride(1) != 1 and e_raw > 0:
            edge_index = edge_index.contiguous()
        self.edge_index, self.edge_attr = edge_index, edge_attr.contiguous()
        self.n_nodes, self.e_raw, self.mode, self.device = int(n_nodes), e_raw, mode, dev
        self.layout = GraphLayout()
        check(lib().pfn_graph_layout_get(self.n_nodes, e_raw, C.byref(self.layout)), "pfn_graph_layout_get")
        need = int(self.layout.total_bytes)
        if workspace is None or workspace.numel() < need:
            workspace = torch.empty(need, dtype=torch.uint8, device=dev)
        self.ws = workspace
        stride = int(self.edge_index.stride(0)) if e_raw > 0 else 0
        self.tile_rows = int(tile_rows)  # the caller's closed-tile promise (0: none), whichever preparation runs
        self.tiled = tile_rows > 0 and bool(lib().pfn_graph_prep_tiled_supported(self.n_nodes, e_raw, int(tile_rows)))
        if self.tiled:
            check(lib().pfn_graph_prep_tiled(_ptr(self.edge_index), max(stride, e_raw), _ptr(self.edge_attr), self.n_nodes,
                                             e_raw, mode, int(tile_rows), self.ws.data_ptr(), _stream()), "pfn_graph_prep_tiled")
        else:
            check(lib().pfn_graph_prep(_ptr(self.edge_index), max(stride, e_raw), _ptr(self.edge_attr), self.n_nodes, e_raw,
                                       mode, self.ws.data_ptr(), _stream()), "pfn_graph_prep")

    # -- views (tests / export) -------------------------------------------------------------------
    def _view(self, off: int, count: int, dtype) -> torch.Tensor:
        nbytes = count * 4
        return self.ws[off:off + nbytes].view(dtype)

    def meta(self) -> Tuple[bool, int, int]:
        """(directed, E, error flag) -- synchronises the stream."""
        host = (C.c_int32 * 3)()
        check(lib().pfn_graph_meta(self.ws.data_ptr(), host, _stream()), "pfn_graph_meta")
        return bool(host[0]), int(host[1]), int(host[2])

    def arrays(self):
        lay, n, cap = self.layout, self.n_nodes, int(self.layout.e_cap)
        return {
            "rowptr_t": self._view(lay.rowptr_t, n + 1, torch.int32), "nbr_t": self._view(lay.nbr_t, cap, torch.int32),
            "eid_t": self._view(lay.eid_t, cap, torch.int32), "ea_t": self._view(lay.ea_t, 2 * cap, torch.float32).view(-1, 2),
            "rowptr_s": self._view(lay.rowptr_s, n + 1, torch.int32), "nbr_s": self._view(lay.nbr_s, cap, torch.int32),
            "eid_s": self._view(lay.eid_s, cap, torch.int32), "ea_s": self._view(lay.ea_s, 2 * cap, torch.float32).view(-1, 2),
            "deg": self._view(lay.deg, n, torch.float32), "dis": self._view(lay.dis, n, torch.float32),
        }

    @property
    def deg(self) -> torch.Tensor:
        return self._view(self.layout.deg, self.n_nodes, torch.float32)

    def export(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """The lists `undirect_graph` returns (networks/MPN.py:506-523)."""
        directed, e, err = self.meta()
        if err:
            raise IndexError("edge_index holds node ids outside [0, num_nodes)")
        if not directed:
            return self.edge_index, self.edge_attr
        ei = torch.empty((2, e), dtype=torch.int64, device=self.device)
        ea = torch.empty((e, 2), dtype=torch.float32, device=self.device)
        check(lib().pfn_graph_export(_ptr(self.edge_index), int(self.edge_index.stride(0)), _ptr(self.edge_attr),
                                     self.e_raw, e, ei.data_ptr(), ea.data_ptr(), _stream()), "pfn_graph_export")
        return ei, ea


# ------------------------------------------------------------------------------------------------
# thin kernel wrappers (tensors carry their own leading dimension = stride(0))
# ------------------------------------------------------------------------------------------------
def _ld(t: torch.Tensor) -> int:
    assert t.dim() == 2 and t.stride(1) == 1, "row-major matrix expected"
    return int(t.stride(0))


def linear_fwd(x, w, ldw, n_in, n_out, bias, out, *, rowscale=None, addend=None, act=ACT_NONE, p=0.0, seed=0,
               inj_mask=None, w_offset=0):
    """out[:, :n_out] = x[:, :n_in] @ W^T (+ rowscale*bias + addend, activation).  `w` is a parameter tensor;
    `w_offset`/`ldw` select a column block of it (W1 = [Wi | Wj | We] of EdgeAggregation)."""
    check(lib().pfn_linear_fwd(x.data_ptr(), _ld(x), w.data_ptr() + 4 * w_offset, ldw, _ptr(bias), _ptr(rowscale),
                               _ptr(addend), _ld(addend) if addend is not None else 0, out.data_ptr(), _ld(out),
                               x.size(0), n_in, n_out, act, float(p), int(seed) & (2**64 - 1), _ptr(inj_mask),
                               _ld(inj_mask) if inj_mask is not None else 0, _stream()), "pfn_linear_fwd")
    return out


def linear_dgrad(dy, w, ldw, n_in, n_out, dx, *, ymask=None, scale=1.0, w_offset=0):
    check(lib().pfn_linear_dgrad(dy.data_ptr(), _ld(dy), w.data_ptr() + 4 * w_offset, ldw, _ptr(ymask),
                                 _ld(ymask) if ymask is not None else 0, float(scale), dx.data_ptr(), _ld(dx),
                                 dy.size(0), n_in, n_out, _stream()), "pfn_linear_dgrad")
    return dx


def linear_wgrad(dy, x, n_in, n_out, dw, lddw, *, dbias=None, rowscale=None, dw_offset=0):
    m = dy.size(0)
    nbytes = int(lib().pfn_linear_wgrad_scratch_bytes(m, n_in, n_out))
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dy.device)
    check(lib().pfn_linear_wgrad(dy.data_ptr(), _ld(dy), x.data_ptr(), _ld(x), _ptr(rowscale),
                                 dw.data_ptr() + 4 * dw_offset, lddw, _ptr(dbias), m, n_in, n_out, scratch.data_ptr(),
                                 _stream()), "pfn_linear_wgrad")
    return dw


def ea_fwd(hi, hj, graph: PreparedGraph, w1, fin, h, s):
    check(lib().pfn_ea_fwd(hi.data_ptr(), hj.data_ptr(), _ld(hi), graph.ws.data_ptr(), graph.n_nodes, graph.e_raw,
                           w1.data_ptr() + 4 * 2 * fin, 2 * fin + 2, s.data_ptr(), _ld(s), h, _stream()), "pfn_ea_fwd")
    return s


def ea_bwd(ds, hi, hj, graph: PreparedGraph, w1, fin, h, dhi, dhj, dw1):
    scratch = torch.empty(int(lib().pfn_ea_bwd_scratch_bytes(h)), dtype=torch.uint8, device=ds.device)
    check(lib().pfn_ea_bwd(ds.data_ptr(), _ld(ds), hi.data_ptr(), hj.data_ptr(), _ld(hi), graph.ws.data_ptr(),
                           graph.n_nodes, graph.e_raw, w1.data_ptr() + 4 * 2 * fin, 2 * fin + 2, dhi.data_ptr(),
                           dhj.data_ptr(), _ld(dhi), dw1.data_ptr() + 4 * 2 * fin, 2 * fin + 2, scratch.data_ptr(), h,
                           _stream()), "pfn_ea_bwd")


def spmm_hop(x, graph: PreparedGraph, y, h, *, transpose=False, addend=None, ymask=None, scale=1.0):
    check(lib().pfn_spmm_hop(x.data_ptr(), _ld(x), graph.ws.data_ptr(), graph.n_nodes, graph.e_raw, int(transpose),
                             _ptr(addend), _ld(addend) if addend is not None else 0, _ptr(ymask),
                             _ld(ymask) if ymask is not None else 0, float(scale), y.data_ptr(), _ld(y), h, _stream()),
          "pfn_spmm_hop")
    return y


# ------------------------------------------------------------------------------------------------
# stand-alone layers (EdgeAggregation.forward / TAGConv.forward called outside the fused model)
# ------------------------------------------------------------------------------------------------
class EdgeAggregationFn(torch.autograd.Function):
    """networks/MPN.py:30-56 as  GEMM(Hi|Hj) -> fused gather/ReLU/sum -> GEMM(W2) (+deg*b2)."""

    @staticmethod
    def forward(ctx, x, edge_index, edge_attr, w1, b1, w2, b2):
        dev = require_cuda(x, edge_index, edge_attr, w1, b1, w2, b2)
        with torch.cuda.device(dev):
            x = x.contiguous().float()
            n, fin = x.shape
            h, fout = w1.size(0), w2.size(0)
            g = PreparedGraph(edge_index, edge_attr, n, mode=0)
            hi, hj, s = new_rows(n, h, dev), new_rows(n, h, dev), new_rows(n, h, dev)
            ldw1 = 2 * fin + 2
            linear_fwd(x, w1, ldw1, fin, h, b1, hi)
            linear_fwd(x, w1, ldw1, fin, h, None, hj, w_offset=fin)
            ea_fwd(hi, hj, g, w1, fin, h, s)
            out = torch.empty((n, fout), dtype=torch.float32, device=dev)
            linear_fwd(s, w2, h, h, fout, b2, out, rowscale=g.deg)
        ctx.save_for_backward(x, w1, w2, hi, hj, s)
        ctx.graph = g
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w1, w2, hi, hj, s = ctx.saved_tensors
        g: PreparedGraph = ctx.graph
        dev = x.device
        with torch.cuda.device(dev):
            dout = dout.contiguous().float()
            n, fin = x.shape
            h, fout = w1.size(0), w2.size(0)
            ldw1 = 2 * fin + 2
            dw1, db1 = torch.empty_like(w1), torch.empty(h, dtype=torch.float32, device=dev)
            dw2, db2 = torch.empty_like(w2), torch.empty(fout, dtype=torch.float32, device=dev)
            linear_wgrad(dout, s, h, fout, dw2, h, dbias=db2, rowscale=g.deg)
            ds = new_rows(n, h, dev)
            linear_dgrad(dout, w2, h, h, fout, ds)
            dhi, dhj = new_rows(n, h, dev), new_rows(n, h, dev)
            ea_bwd(ds, hi, hj, g, w1, fin, h, dhi, dhj, dw1)
            linear_wgrad(dhi, x, fin, h, dw1, ldw1, dbias=db1)
            linear_wgrad(dhj, x, fin, h, dw1, ldw1, dw_offset=fin)
            dx = None
            if ctx.needs_input_grad[0]:
                dx = torch.empty_like(x)
                tmp = torch.empty_like(x)
                linear_dgrad(dhi, w1, ldw1, fin, h, dx)
                linear_dgrad(dhj, w1, ldw1, fin, h, tmp, w_offset=fin)
                dx += tmp  # torch elementwise add: only on this stand-alone path (the fused model accumulates in-kernel)
        return dx, None, None, dw1, db1, dw2, db2


class TAGConvFn(torch.autograd.Function):
    """PyG TAGConv.forward (call site networks/MPN.py:545): K hops + one K-segmented Linear."""

    @staticmethod
    def forward(ctx, x, edge_index, bias, *lin_weights):
        dev = require_cuda(x, edge_index, bias, *lin_weights)
        with torch.cuda.device(dev):
            x = x.contiguous().float()
            n, fin = x.shape
            fout = lin_weights[0].size(0)
            k_hops = len(lin_weights) - 1
            dummy_attr = torch.zeros((edge_index.size(1), 2), dtype=torch.float32, device=dev)
            g = PreparedGraph(edge_index, dummy_attr, n, mode=0)
            ld = round_up4(fin)
            xcat = torch.zeros((n, (k_hops + 1) * ld), dtype=torch.float32, device=dev)
            xcat[:, :fin].copy_(x)
            for k in range(1, k_hops + 1):
                spmm_hop(xcat[:, (k - 1) * ld:k * ld], g, xcat[:, k * ld:(k + 1) * ld], fin)
            out = torch.zeros((n, fout), dtype=torch.float32, device=dev)
            for k in range(k_hops + 1):
                linear_fwd(xcat[:, k * ld:(k + 1) * ld], lin_weights[k], fin, fin, fout, bias if k == 0 else None, out,
                           addend=out if k > 0 else None)
        ctx.save_for_backward(xcat, *lin_weights)
        ctx.graph, ctx.dims = g, (n, fin, fout, k_hops, ld)
        return out

    @staticmethod
    def backward(ctx, dout):
        xcat, *lin_weights = ctx.saved_tensors
        g: PreparedGraph = ctx.graph
        n, fin, fout, k_hops, ld = ctx.dims
        dev = xcat.device
        with torch.cuda.device(dev):
            dout = dout.contiguous().float()
            dbias = torch.empty(fout, dtype=torch.float32, device=dev)
            dws = []
            dxcat = torch.zeros_like(xcat)
            for k in range(k_hops + 1):
                dw = torch.empty_like(lin_weights[k])
                linear_wgrad(dout, xcat[:, k * ld:(k + 1) * ld], fin, fout, dw, fin, dbias=dbias if k == 0 else None)
                dws.append(dw)
                linear_dgrad(dout, lin_weights[k], fin, fin, fout, dxcat[:, k * ld:(k + 1) * ld])
            for k in range(k_hops, 0, -1):
                prev = dxcat[:, (k - 1) * ld:k * ld]
                spmm_hop(dxcat[:, k * ld:(k + 1) * ld], g, prev, fin, transpose=True, addend=prev)
            dx = dxcat[:, :fin].contiguous() if ctx.needs_input_grad[0] else None
        return (dx, None, dbias, *dws)
