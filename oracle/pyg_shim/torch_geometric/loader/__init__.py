"""PyG `loader.DataLoader` subset (test infrastructure): a torch DataLoader whose collate is
`Batch.from_data_list` (reference call sites: train.py:90-92, utils/training.py:6)."""
import torch

from ..data import Batch


class DataLoader(torch.utils.data.DataLoader):
    def __init__(self, dataset, batch_size=1, shuffle=False, **kwargs):
        kwargs.pop("collate_fn", None)
        super().__init__(dataset, batch_size, shuffle, collate_fn=Batch.from_data_list, **kwargs)
