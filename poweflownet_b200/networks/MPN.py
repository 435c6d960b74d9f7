"""Drop-in `nn.Module` surface of the reference hot path (reference: networks/MPN.py).

Same class names, constructor signatures, attribute names (`layers`, `mask_embd`, `dropout`) and
therefore the same `state_dict` keys/shapes as the reference (`layers.{i}.edge_aggr.{0,2}.{weight,bias}`,
`layers.{i}.lins.{k}.weight`, `layers.{i}.bias`, `mask_embd.{0,2}.{weight,bias}`), so reference
checkpoints load and `train.py` can construct and drive the model unchanged.
The arithmetic runs in libpfn_b200.so (hand-written sm_100a kernels) -- CUDA tensors only.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import List, Optional, Sequence

import torch
from torch import nn

from .. import ops
from .._lib import MpnDesc, check, lib


class EdgeAggregation(nn.Module):
    """networks/MPN.py:6-56.  Parameters live in `edge_aggr` exactly as in the reference
    (`nn.Sequential(Linear(2*nf+ef, hidden), ReLU, Linear(hidden, out))`, :17-21)."""

    def __init__(self, nfeature_dim, efeature_dim, hidden_dim, output_dim):
        super().__init__()
        self.nfeature_dim, self.efeature_dim, self.output_dim = nfeature_dim, efeature_dim, output_dim
        self.edge_aggr = nn.Sequential(
            nn.Linear(nfeature_dim * 2 + efeature_dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, output_dim))

    def forward(self, x, edge_index, edge_attr):
        """`out[i] = sum_{e: edge_index[1,e]=i} edge_aggr(cat[x_i, x_j, edge_attr_e])` (:23-28,:53).
        The degree norm of :43-47 never reaches `message` in the reference and is not computed."""
        if self.efeature_dim != 2:
            raise NotImplementedError("the sm_100a path implements efeature_dim == 2 (the dataset's edge width)")
        l0, l2 = self.edge_aggr[0], self.edge_aggr[2]
        return ops.EdgeAggregationFn.apply(x, edge_index, edge_attr, l0.weight, l0.bias, l2.weight, l2.bias)


class _Lin(nn.Module):
    """Bias-free linear holder with PyG `Linear`'s state_dict key (`weight` [out, in]) and default init."""

    def __init__(self, fin, fout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(fout, fin))
        self.reset_parameters()

    def reset_parameters(self):
        bound = 1.0 / math.sqrt(self.weight.size(1))
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)


class TAGConv(nn.Module):
    """`torch_geometric.nn.TAGConv(in, out, K)` as the reference uses it (networks/MPN.py:477,480,484,545):
    `out = sum_k lins[k](A_hat^k x) + bias`, `A_hat = D^-1/2 A D^-1/2`, degree over `edge_index[1]`."""

    def __init__(self, in_channels, out_channels, K=3):
        super().__init__()
        self.in_channels, self.out_channels, self.K = in_channels, out_channels, K
        self.lins = nn.ModuleList([_Lin(in_channels, out_channels) for _ in range(K + 1)])
        self.bias = nn.Parameter(torch.zeros(out_channels))

    def forward(self, x, edge_index):
        return ops.TAGConvFn.apply(x, edge_index, self.bias, *[l.weight for l in self.lins])


class _Workspace:
    __slots__ = ("graph", "act", "scratch")

    def __init__(self, graph, act, scratch):
        self.graph, self.act, self.scratch = graph, act, scratch


class _MPNFunction(torch.autograd.Function):
    """Whole-model forward/backward through `pfn_mpn_forward` / `pfn_mpn_backward`."""

    @staticmethod
    def forward(ctx, model, tiling, needs_grad, x, pred_mask, edge_index, edge_attr, *params):
        dev = ops.require_cuda(x, pred_mask, edge_index, edge_attr, *params)
        n = int(x.size(0))
        # tiling: (tile_rows, graph_ptr, alt_ptr).  graph_ptr: device int64 [G+1] for batches that mix graph sizes, else
        # None; alt_ptr: the batch's `ptr` when the uniform tiling was chosen (second attempt if its promise is broken)
        tile_rows, graph_ptr, alt_ptr = tiling
        training = bool(model.training)
        # `needs_grad` comes from the module (grad mode and requires_grad of the inputs): inside Function.forward grad
        # mode is off and ctx.needs_input_grad is True for every parameter even under torch.no_grad()
        with torch.cuda.device(dev):
            x = x.contiguous().float()
            pred_mask = pred_mask.contiguous()
            if pred_mask.dtype != torch.int64:
                pred_mask = pred_mask.long()
            ws = model._take_workspace(n, int(edge_index.size(1)), dev)
            e_raw = int(edge_index.size(1))
            # attempts at the graph-resident route, cheapest first: (tile rows, graph_ptr, one-launch preparation?)
            # * equal-sized tiles with the one-launch preparation (it validates its own guess about the edge layout;
            #   PFN_PREP_TILED=0 keeps the general five-pass preparation);
            # * the same tiles on the general preparation;
            # * whole graphs packed from `ptr` (equal-sized tiles refused, or a batch that mixes graph sizes).
            attempts = []
            if tile_rows > 0:
                if (graph_ptr is None and os.environ.get("PFN_PREP_TILED", "1") != "0"
                        and lib().pfn_graph_prep_tiled_supported(n, e_raw, tile_rows)):
                    attempts.append((tile_rows, None, True))
                attempts.append((tile_rows, graph_ptr, False))
                if graph_ptr is None and alt_ptr is not None:
                    attempts.append((128, alt_ptr, False))

            def sig_of(att):  # key of `_tiling_checked`: what was promised about this batch shape
                if att[2]:
                    return ("prep_tiled", n, e_raw, att[0])
                return (n, e_raw, att[0], int(att[1].numel()) - 1 if att[1] is not None else 0)

            attempts = [att for att in attempts if model._tiling_checked.get(sig_of(att), True) is not False]
            fast = bool(attempts) and attempts[0][2]
            graph = ops.PreparedGraph(edge_index, edge_attr, n, mode=1, workspace=ws.graph, tile_rows=attempts[0][0] if fast else 0)
            ws.graph = graph.ws
            seed_dev = model._seed_device if training else None  # device-resident seed (CUDA-graph replays)
            seed = model._next_seed() if (training and seed_dev is None and model.dropout.p > 0) else 0
            inj = model._inject_dropout_masks if training else None
            inj_table = None
            if inj is not None:
                inj = [m.to(dev, torch.float32).contiguous() for m in inj]
                inj_table = (C.c_void_p * len(model.layers))(*([m.data_ptr() for m in inj] + [None]))
            out = torch.empty((n, model.output_dim), dtype=torch.float32, device=dev)
            ptable = (C.c_void_p * len(params))(*[p.data_ptr() for p in params])
            desc = model._desc()
            common = (C.byref(desc), ptable, x.data_ptr(), pred_mask.data_ptr(), n, e_raw, ws.graph.data_ptr(),
                      ws.act.data_ptr(), ws.scratch.data_ptr(), int(training), seed,
                      None if seed_dev is None else seed_dev.data_ptr(), inj_table, out.data_ptr())
            stream = torch.cuda.current_stream().cuda_stream
            capturing = torch.cuda.is_current_stream_capturing()
            tile_rows, n_graphs, refused = 0, 0, False
            for att in attempts:
                # graph-resident kernel: the whole layer stack in one launch, one tile of whole graphs per CTA
                rows, gptr, fast = att
                sig, ng = sig_of(att), (int(gptr.numel()) - 1 if gptr is not None else 0)
                if refused or graph.tiled != fast:  # a refused attempt leaves its violation flag in the graph workspace
                    graph = ops.PreparedGraph(edge_index, edge_attr, n, mode=1, workspace=ws.graph, tile_rows=rows if fast else 0)
                    refused = False
                check(lib().pfn_mpn_forward_tiled(*common, rows, None if gptr is None else gptr.data_ptr(), ng, stream),
                      "pfn_mpn_forward_tiled")
                # the kernels validate the promises themselves (a violation raises meta[6] and poisons the tile's rows
                # with NaN); the flag is read back for the first batch of a shape and then every `_tiling_recheck`
                # batches of it -- never during stream capture
                calls = model._tiling_calls.get(sig, 0)
                model._tiling_calls[sig] = calls + 1
                if not capturing and (sig not in model._tiling_checked or calls % model._tiling_recheck == 0):
                    violated = C.c_int32(0)
                    check(lib().pfn_graph_tile_status(ws.graph.data_ptr(), C.byref(violated), stream), "pfn_graph_tile_status")
                    model._tiling_checked[sig] = not violated.value
                    if violated.value:
                        refused = True
                        continue
                tile_rows, n_graphs = rows, ng
                break
            if tile_rows <= 0 and graph.tiled:  # the layer-wise route needs the general preparation's arrays
                graph = ops.PreparedGraph(edge_index, edge_attr, n, mode=1, workspace=ws.graph)
            if tile_rows <= 0:
                check(lib().pfn_mpn_forward(*common, stream), "pfn_mpn_forward")
        if needs_grad:
            ctx.model, ctx.ws, ctx.params, ctx.n, ctx.e_raw, ctx.training = model, ws, params, n, graph.e_raw, training
            ctx.tile_rows = tile_rows  # > 0: the closed-tile promise held for this batch (validated by the forward kernel)
            ctx.n_graphs = n_graphs if tile_rows > 0 else 0
            ctx.keep = (x, pred_mask, graph, inj)  # keep inputs alive until backward
        else:
            ctx.ws = None
            model._give_workspace(n, graph.e_raw, dev, ws)
        return out

    @staticmethod
    def backward(ctx, dout):
        if getattr(ctx, "ws", None) is None:
            raise RuntimeError("backward on a MaskEmbdMultiMPN forward whose activations were released (called twice, or "
                               "the forward ran without grad mode)")
        model, ws, params = ctx.model, ctx.ws, ctx.params
        dev = dout.device
        with torch.cuda.device(dev):
            dout = dout.contiguous().float()
            sizes = [p.numel() for p in params]
            mult = int(getattr(model._grad_reducer, "pad_multiple", 4))  # whole float4s (per rank slice) for the peer all-reduce
            gflat = torch.empty((sum(sizes) + mult - 1) // mult * mult, dtype=torch.float32, device=dev)
            if gflat.numel() != sum(sizes):
                gflat[sum(sizes):].zero_()
            views, off = [], 0
            for p, s in zip(params, sizes):
                views.append(gflat[off:off + s].view(p.shape))
                off += s
            ptable = (C.c_void_p * len(params))(*[p.data_ptr() for p in params])
            gtable = (C.c_void_p * len(params))(*[v.data_ptr() for v in views])
            desc = model._desc()
            common = (C.byref(desc), ptable, gtable, dout.data_ptr(), ctx.n, ctx.e_raw, ws.graph.data_ptr(),
                      ws.act.data_ptr(), ws.scratch.data_ptr(), int(ctx.training))
            stream = torch.cuda.current_stream().cuda_stream
            if ctx.tile_rows > 0:
                check(lib().pfn_mpn_backward_tiled(*common, ctx.tile_rows, ctx.n_graphs, stream), "pfn_mpn_backward_tiled")
            else:
                check(lib().pfn_mpn_backward(*common, stream), "pfn_mpn_backward")
            dx = None
            if ctx.needs_input_grad[3]:
                # x enters as `mask_embd(mask) + x` (MPN.py:537): d loss / d x is the gradient w.r.t. that sum
                dx = model._dx0_view(ws, ctx.n).clone()
            if model._grad_reducer is not None:
                model._grad_reducer(gflat)  # data parallel: ONE collective over the flat gradient buffer
        model._give_workspace(ctx.n, ctx.e_raw, dev, ws)
        ctx.ws = None
        return (None, None, None, dx, None, None, None, *views)  # (model, tiling, needs_grad, x, pred_mask, edge_index, edge_attr, *params)


class MaskEmbdMultiMPN(nn.Module):
    """networks/MPN.py:456-559 -- constructor mirrors :462-496 argument for argument."""

    def __init__(self, nfeature_dim, efeature_dim, output_dim, hidden_dim, n_gnn_layers, K, dropout_rate):
        super().__init__()
        self.nfeature_dim, self.efeature_dim, self.output_dim = nfeature_dim, efeature_dim, output_dim
        self.hidden_dim, self.n_gnn_layers, self.K, self.dropout_rate = hidden_dim, n_gnn_layers, K, dropout_rate
        self.layers = nn.ModuleList()
        self.layers.append(EdgeAggregation(nfeature_dim, efeature_dim, hidden_dim, hidden_dim))
        # :475-480 -- with one GNN layer the TAGConv emits output_dim (kept for state_dict fidelity)
        self.layers.append(TAGConv(hidden_dim, output_dim if n_gnn_layers == 1 else hidden_dim, K=K))
        for _ in range(n_gnn_layers - 2):
            self.layers.append(EdgeAggregation(hidden_dim, efeature_dim, hidden_dim, hidden_dim))
            self.layers.append(TAGConv(hidden_dim, hidden_dim, K=K))
        self.layers.append(EdgeAggregation(hidden_dim, efeature_dim, hidden_dim, output_dim))
        self.mask_embd = nn.Sequential(
            nn.Linear(nfeature_dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, nfeature_dim))
        self.dropout = nn.Dropout(self.dropout_rate, inplace=False)  # kept for attribute parity; p is read from it
        self._pool = {}
        self._inject_dropout_masks: Optional[Sequence[torch.Tensor]] = None  # test hook: replay given keep-masks
        self._grad_reducer = None  # set by poweflownet_b200.parallel.attach_gradient_allreduce
        self._seed_device: Optional[torch.Tensor] = None  # int64[1] on the device: dropout seed read by the kernels
        self.fused = True  # use the graph-resident kernel when the batch is made of equal-sized graphs of <= 128 nodes
        self._tiling_checked = {}  # (N, E_raw, tile_rows, n_graphs) -> did the kernel's closed-tile validation pass
        self._tiling_calls = {}    # same key -> forwards taken on the graph-resident route
        self._tiling_recheck = 64  # the validation flag is read back (one stream sync) every this many batches of a shape

    # ---- reference helper methods (networks/MPN.py:498-523) -------------------------------------
    def is_directed(self, edge_index):
        """True iff the reverse of the FIRST edge is absent (the reference reads one edge only); False for
        an empty edge list.  Evaluated on the device; returning a Python bool costs one sync, as in the reference."""
        if edge_index.shape[1] == 0:
            return False
        dev = ops.require_cuda(edge_index)
        with torch.cuda.device(dev):
            n = int(edge_index.max().item()) + 1
            dummy = torch.zeros((edge_index.size(1), 2), dtype=torch.float32, device=dev)
            return ops.PreparedGraph(edge_index, dummy, n, mode=1).meta()[0]

    def undirect_graph(self, edge_index, edge_attr):
        dev = ops.require_cuda(edge_index, edge_attr)
        if edge_index.shape[1] == 0:
            return edge_index, edge_attr
        with torch.cuda.device(dev):
            n = int(edge_index.max().item()) + 1
            return ops.PreparedGraph(edge_index, edge_attr.float(), n, mode=1).export()

    # ---- plumbing ---------------------------------------------------------------------------------
    def _next_seed(self) -> int:
        """Dropout seed of one training forward, drawn from torch's global CPU generator (reproducible under
        `torch.manual_seed`, re-seeding replays the same masks).  Only called when dropout_rate > 0: with dropout off
        the module consumes no random numbers at all, exactly like `nn.Dropout(p=0)` in the reference, so a
        `DataLoader(shuffle=True)` fed from the same global stream shuffles identically for both
        (tests/test_gpu_dropin_train.py compares the two trajectories batch for batch)."""
        return int(torch.empty((), dtype=torch.int64).random_().item())

    def _desc(self) -> MpnDesc:
        return MpnDesc(self.nfeature_dim, self.efeature_dim, self.output_dim, self.hidden_dim, self.n_gnn_layers,
                       self.K, float(self.dropout.p), 0)

    def _engine_params(self) -> List[torch.Tensor]:
        """Parameters in the order of pfn_mpn_num_params (include/pfn_b200.h)."""
        out: List[torch.Tensor] = []
        for layer in self.layers:
            if isinstance(layer, EdgeAggregation):
                l0, l2 = layer.edge_aggr[0], layer.edge_aggr[2]
                out += [l0.weight, l0.bias, l2.weight, l2.bias]
            else:
                out += [l.weight for l in layer.lins] + [layer.bias]
        out += [self.mask_embd[0].weight, self.mask_embd[0].bias, self.mask_embd[2].weight, self.mask_embd[2].bias]
        return out

    def _take_workspace(self, n, e_raw, dev) -> _Workspace:
        free = self._pool.setdefault((n, e_raw, str(dev)), [])
        if free:
            return free.pop()
        act, scratch = C.c_size_t(), C.c_size_t()
        desc = self._desc()
        check(lib().pfn_mpn_workspace(C.byref(desc), n, e_raw, C.byref(act), C.byref(scratch)), "pfn_mpn_workspace")
        return _Workspace(None, torch.zeros(act.value, dtype=torch.uint8, device=dev),
                          torch.zeros(scratch.value, dtype=torch.uint8, device=dev))

    def _give_workspace(self, n, e_raw, dev, ws) -> None:
        free = self._pool.setdefault((n, e_raw, str(dev)), [])
        if len(free) < 2:
            free.append(ws)

    def _swap_pool(self, pool: dict) -> dict:
        """Install `pool` as the workspace pool and return the previous one.  `training.GraphedMSEStep` captures its
        CUDA graph on a PRIVATE pool that it keeps alive itself: the captured kernels have the workspace addresses baked
        in, so those buffers must never be handed to (or dropped by) another forward of the same shape."""
        old, self._pool = self._pool, pool
        return old

    def _dx0_view(self, ws: _Workspace, n: int) -> torch.Tensor:
        # scratch layout of engine.cu: dx0 [n, nfeat] comes first
        return ws.scratch.view(torch.float32)[:n * self.nfeature_dim].view(n, self.nfeature_dim)

    def _tiling(self, data):
        """(tile_rows, graph_ptr, alt_ptr) for the graph-resident kernels, or (0, None, None) for the layer-wise path.
        Uses only host-side shape information (no device read):
        * `num_graphs` equal-sized graphs of n = N / num_graphs <= 128 nodes: floor(128 / n) whole graphs per tile
          (`alt_ptr` = the batch's device `ptr`, if any: the variable tiling below is tried next when the kernel
          refuses the uniform one);
        * otherwise, when the batch carries PyG's `ptr` on the device: graphs of mixed sizes (the reference's
          `--case mixed`), packed greedily into tiles of <= 128 rows by a device-side pass over `ptr`.
        Either way the kernel validates the tiling itself (no edge may leave a tile, no graph may exceed 128 nodes)."""
        none = (0, None, None)
        if not self.fused:
            return none
        n = int(data.x.size(0))
        ptr = getattr(data, "ptr", None)
        g = getattr(data, "num_graphs", None)
        if g is None and ptr is not None:
            g = int(ptr.numel()) - 1
        if not g or g <= 0 or n <= 0:
            return none
        e_raw = int(data.edge_index.size(1))
        desc = self._desc()
        ptr_ok = (ptr is not None and ptr.is_cuda and ptr.dtype == torch.int64 and ptr.numel() == g + 1 and g <= n
                  and self._tiling_checked.get((n, e_raw, 128, g), True) is not False
                  and bool(lib().pfn_mpn_fused_supported(C.byref(desc), 128)))
        if n % g == 0 and n // g <= 128:
            tile = (128 // (n // g)) * (n // g)
            if self._tiling_checked.get((n, e_raw, tile, 0), True) is not False and lib().pfn_mpn_fused_supported(C.byref(desc), tile):
                return tile, None, (ptr.contiguous() if ptr_ok else None)
            if not ptr_ok:
                return none
        if ptr_ok:
            return 128, ptr.contiguous(), None
        return none

    def _tile_rows(self, data) -> int:
        return self._tiling(data)[0]

    def forward(self, data):
        """networks/MPN.py:525-559.  `data` is any object with the PyG `Batch` attributes the reference
        reads: x [N,4], pred_mask [N,4], edge_index [2,E_raw], edge_attr [E_raw,2] (bus_type / batch are
        read but unused by the reference, :531-532)."""
        assert data.x.shape[-1] == 4  # :528
        if self.n_gnn_layers < 2:
            raise NotImplementedError(
                "n_gnn_layers == 1 builds EA(4->h), TAG(h->out), EA(h->out) in the reference (MPN.py:475-477,489), "
                "which cannot run unless hidden_dim == output_dim; the sm_100a path implements n_gnn_layers >= 2")
        if self.efeature_dim != 2:
            raise NotImplementedError("the sm_100a path implements efeature_dim == 2 (the dataset's edge width)")
        params = self._engine_params()
        needs_grad = torch.is_grad_enabled() and (data.x.requires_grad or any(p.requires_grad for p in params))
        return _MPNFunction.apply(self, self._tiling(data), needs_grad, data.x, data.pred_mask, data.edge_index,
                                  data.edge_attr, *params)
