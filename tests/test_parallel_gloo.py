"""Data-parallel host logic on CPU with the gloo backend (world_size 2): graph sharding at `ptr` boundaries, global
loss normalisation and the single flat-buffer gradient all-reduce reproduce the single-process gradient of the whole
batch.  The arithmetic here is the oracle's (CPU); the CUDA path uses the same `poweflownet_b200.parallel` helpers."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import common
from oracle import pfn_oracle as O
from poweflownet_b200 import parallel
from poweflownet_b200.data import shard_batch, synthetic_batch

KW = dict(common.MODEL_DIMS, hidden_dim=16, n_gnn_layers=3, K=2, dropout_rate=0.0)
CASES = ["14", "118v2", "14", (9, 12), "14", "118v2"]  # variable-N batch: ranks end up with different node counts


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _flat_grad(model):
    return torch.cat([p.grad.reshape(-1) for p in model.parameters()])


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    r, w, _ = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    full = synthetic_batch(cases=CASES, seed=3)
    mine = shard_batch(full, rank, world)
    model = O.MaskEmbdMultiMPN(**KW)
    if rank != 0:  # deliberately different weights: broadcast must fix it
        with torch.no_grad():
            for p in model.parameters():
                p.add_(1.0)
    else:
        common.load_seeded(model)
    parallel.broadcast_parameters(model, src=0)
    total = parallel.global_count(mine.num_nodes * KW["output_dim"])
    out = model(mine)
    loss = ((out - mine.y) ** 2).sum() / total  # MSE(mean) over the GLOBAL element count
    loss.backward()
    flat = _flat_grad(model)
    parallel.allreduce_flat_(flat)
    loss_sum = loss.detach().clone()
    dist.all_reduce(loss_sum)
    # Masked_L2_loss (the reference's default loss) under data parallelism: the two means run over the GLOBAL masked /
    # unmasked selections (`parallel.global_mask_counts`: one 2-element all-reduce), which is what the CUDA loss kernel
    # reads through its `global_counts` argument
    model.zero_grad()
    counts = parallel.global_mask_counts(mine.pred_mask)
    out = model(mine)
    d2 = (out - mine.y) ** 2
    m = mine.pred_mask != 0
    loss_m = d2[m].sum() / counts[0] + 0.5 * d2[~m].sum() / counts[1]
    loss_m.backward()
    flat_m = _flat_grad(model)
    parallel.allreduce_flat_(flat_m)
    loss_m_sum = loss_m.detach().clone()
    dist.all_reduce(loss_m_sum)
    if rank == 0:
        torch.save({"flat": flat, "loss": loss_sum, "total": total, "nodes": mine.num_nodes, "flat_masked": flat_m,
                    "loss_masked": loss_m_sum, "counts": counts}, out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gradient_equals_single_process(tmp_path):
    out_path = str(tmp_path / "rank0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out_path), nprocs=2, join=True)
    got = torch.load(out_path, weights_only=True)
    full = synthetic_batch(cases=CASES, seed=3)
    ref = common.load_seeded(O.MaskEmbdMultiMPN(**KW))
    loss_ref, _ = O.forward_loss_backward(ref, full, "mse")
    assert got["total"] == full.num_nodes * KW["output_dim"]
    assert 0 < got["nodes"] < full.num_nodes
    assert abs(float(got["loss"]) - float(loss_ref)) < 1e-6 * float(loss_ref)
    e = common.rel_err(got["flat"], _flat_grad(ref))
    assert max(e) < 1e-5, e
    # masked loss: global counts, summed loss and reduced gradient equal the single-process Masked_L2_loss(regcoeff=0.5)
    ref.zero_grad()
    loss_m = O.masked_l2_loss(ref(full), full.y, full.pred_mask, True, 0.5)
    loss_m.backward()
    n1 = int((full.pred_mask != 0).sum())
    assert got["counts"].tolist() == [float(n1), float(full.pred_mask.numel() - n1)]
    assert abs(float(got["loss_masked"]) - float(loss_m)) < 1e-6 * float(loss_m)
    e = common.rel_err(got["flat_masked"], _flat_grad(ref))
    assert max(e) < 1e-5, e


def test_shards_partition_the_batch():
    full = synthetic_batch(cases=CASES, seed=3)
    for world in (1, 2, 3, 6):
        shards = [shard_batch(full, r, world) for r in range(world)]
        assert sum(s.num_graphs for s in shards) == full.num_graphs
        assert sum(s.num_nodes for s in shards) == full.num_nodes
        assert sum(s.edge_index.size(1) for s in shards) == full.edge_index.size(1)
        assert torch.equal(torch.cat([s.x for s in shards]), full.x)
        for s in shards:
            assert s.num_graphs >= 1 and int(s.ptr[0]) == 0 and int(s.ptr[-1]) == s.num_nodes
            if s.edge_index.numel():
                assert 0 <= int(s.edge_index.min()) and int(s.edge_index.max()) < s.num_nodes
