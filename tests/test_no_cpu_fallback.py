"""The product path has no CPU fallback: every entry above the C ABI refuses CPU tensors loudly (runs without a GPU)."""
import pytest
import torch

import make_golden_dataset as mgd
import make_golden_losses as mgl


def test_losses_refuse_cpu_tensors():
    from poweflownet_b200 import losses
    batch, pred = mgl.loss_predictions("tiny")
    x = pred.clone().requires_grad_(True)
    with pytest.raises(RuntimeError, match="CUDA"):
        losses.PowerImbalance(*mgl.loss_stats())(x, batch.edge_index, batch.edge_attr)
    with pytest.raises(RuntimeError, match="CUDA"):
        losses.MixedMSEPoweImbalance(*mgl.loss_stats(), alpha=0.9)(x, batch.edge_index, batch.edge_attr, batch.y)
    with pytest.raises(RuntimeError, match="CUDA"):
        losses.Masked_L2_loss()(x, batch.y, batch.pred_mask)


def test_optimizer_refuses_cpu_parameters():
    from poweflownet_b200.optim import FusedAdamW
    p = torch.nn.Parameter(torch.zeros(3))
    p.grad = torch.ones(3)
    opt = FusedAdamW([p])
    with pytest.raises(RuntimeError, match="CUDA"):
        opt.step()
    with pytest.raises(ValueError):
        FusedAdamW([p], lr=-1.0)


def test_dataset_refuses_a_cpu_device_and_ragged_splits():
    from poweflownet_b200.datasets import PowerFlowData
    gold = torch.load(mgd.dataset_golden_path("ds_case14"), weights_only=False)
    raws = [(g["edge_features"], g["node_features"]) for g in gold["raw"].values()]
    with pytest.raises(RuntimeError, match="GPU"):
        PowerFlowData(case="14", split=mgd.SPLIT, device="cpu", raw=raws)


def test_train_epoch_refuses_a_cpu_model():
    from poweflownet_b200.networks.MPN import MaskEmbdMultiMPN
    from poweflownet_b200.training import train_epoch
    import common
    model = MaskEmbdMultiMPN(**common.model_kwargs("tiny"))
    batch = common.make_batch("tiny")
    opt = torch.optim.AdamW(model.parameters())
    with pytest.raises(RuntimeError):
        train_epoch(model, [batch], torch.nn.MSELoss(), opt, "cpu")
