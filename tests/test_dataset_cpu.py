"""The oracle's restatement of datasets/PowerFlowData.py (`process`, `_normalize_dataset`) and of PyG's batch collation
against fixtures produced by the reference's own dataset class (tests/golden/make_golden_dataset.py).  Bit-exact."""
import pytest
import torch

import make_golden_dataset as mgd
from oracle import pfn_oracle as O


def _raws(gold):
    return [(gold["raw"][tag]["edge_features"], gold["raw"][tag]["node_features"]) for tag in gold["raw"]]


@pytest.mark.parametrize("name", list(mgd.DATASET_CASES))
def test_fixture_inputs_reproducible(name):
    gold = torch.load(mgd.dataset_golden_path(name), weights_only=False)
    for (tag, e, nf), (gtag, g) in zip(mgd.raw_arrays(name), gold["raw"].items()):
        assert tag == gtag
        assert torch.equal(torch.from_numpy(e), g["edge_features"]) and torch.equal(torch.from_numpy(nf), g["node_features"])


@pytest.mark.parametrize("name", list(mgd.DATASET_CASES))
def test_oracle_dataset_matches_reference(name):
    torch.set_num_threads(1)
    gold = torch.load(mgd.dataset_golden_path(name), weights_only=False)
    for task, sp in gold["splits"].items():
        samples = O.process_split(_raws(gold), mgd.SPLIT, task)
        assert len(samples) == sp["len"]
        stats = O.dataset_stats(samples)
        for a, b in zip(stats, sp["stats"]):
            assert torch.equal(a, b), (task, a, b)
        for bt in sp["batches"]:
            got = O.collate_batch(samples, bt["ids"], stats)
            for f in mgd.FIELDS:
                assert got[f].dtype == bt[f].dtype and torch.equal(got[f], bt[f]), (task, f)
    val = gold["val_with_train_stats"]
    got = O.collate_batch(O.process_split(_raws(gold), mgd.SPLIT, "val"), val["ids"], gold["splits"]["train"]["stats"])
    for f in mgd.FIELDS:
        assert torch.equal(got[f], val[f]), f


def test_ragged_split_raises_like_the_reference():
    e, nf = torch.zeros(7, 3, 4), torch.zeros(7, 2, 6)  # int(7 * .5) + int(7 * .2) + int(7 * .3) = 6 != 7
    with pytest.raises(RuntimeError):
        O.process_split([(e, nf)], mgd.SPLIT, "train")
