"""PyG `data` subset (test infrastructure): `Data`, `Batch`, `Dataset`, `InMemoryDataset`.

Only what `datasets/PowerFlowData.py` and `train.py` of the reference touch:
`Data(**tensors)`, `.to()`, `len(data)` (= number of stored attributes, which is what
utils/training.py:76-77 multiplies the loss by), `Batch.from_data_list` (concat on dim 0,
`edge_index` on dim 1 with cumulative node offsets, adds `batch` and `ptr`),
`InMemoryDataset` (`process()` on missing processed files, `collate`, `self.data/self.slices`,
indexing with `transform`, slicing, `len()`).
"""
import copy
import os

import torch


def _cat_dim(key):
    return 1 if "index" in key else 0  # PyG `Data.__cat_dim__`: *index* attributes concatenate on the last dim


class Data:
    def __init__(self, **kwargs):
        self.__dict__["_store"] = {}
        for k, v in kwargs.items():
            if v is not None:
                self._store[k] = v

    # attribute protocol ---------------------------------------------------------------------
    def __getattr__(self, key):
        store = self.__dict__.get("_store")
        if store is not None and key in store:
            return store[key]
        raise AttributeError(key)

    def __setattr__(self, key, value):
        if key in ("_store",):
            self.__dict__[key] = value
        elif value is None:
            self._store.pop(key, None)
        else:
            self._store[key] = value

    def __getitem__(self, key):
        return self._store[key]

    def __setitem__(self, key, value):
        self._store[key] = value

    def __contains__(self, key):
        return key in self._store

    def __getstate__(self):
        return {"_store": self._store}

    def __setstate__(self, state):
        self.__dict__["_store"] = state["_store"]

    def keys(self):
        return list(self._store.keys())

    def __len__(self):  # PyG BaseData.__len__: number of attributes, NOT graphs
        return len(self._store)

    def __iter__(self):
        return iter(self._store.items())

    @property
    def num_nodes(self):
        if "x" in self._store:
            return self._store["x"].size(0)
        if "edge_index" in self._store and self._store["edge_index"].numel():
            return int(self._store["edge_index"].max()) + 1
        return 0

    @property
    def num_edges(self):
        return self._store["edge_index"].size(1) if "edge_index" in self._store else 0

    def to(self, device, non_blocking=False):
        out = copy.copy(self)
        out.__dict__["_store"] = {
            k: (v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v) for k, v in self._store.items()
        }
        return out

    def clone(self):
        out = copy.copy(self)
        out.__dict__["_store"] = {k: (v.clone() if torch.is_tensor(v) else copy.deepcopy(v)) for k, v in self._store.items()}
        return out

    def __repr__(self):
        body = ", ".join(f"{k}={list(v.shape) if torch.is_tensor(v) else v}" for k, v in self._store.items())
        return f"{type(self).__name__}({body})"


class Batch(Data):
    @classmethod
    def from_data_list(cls, data_list):
        keys = data_list[0].keys()
        out = cls()
        offsets, n_acc = [], 0
        for d in data_list:
            offsets.append(n_acc)
            n_acc += d.num_nodes
        for k in keys:
            vals = [d[k] for d in data_list]
            if not torch.is_tensor(vals[0]):
                out[k] = vals
                continue
            if k == "edge_index":
                vals = [v + off for v, off in zip(vals, offsets)]
            out[k] = torch.cat(vals, dim=_cat_dim(k))
        sizes = torch.tensor([d.num_nodes for d in data_list], dtype=torch.long)
        out["batch"] = torch.repeat_interleave(torch.arange(len(data_list)), sizes)
        out["ptr"] = torch.cat([torch.zeros(1, dtype=torch.long), sizes.cumsum(0)])
        out.__dict__["_num_graphs"] = len(data_list)
        return out

    @property
    def num_graphs(self):
        return self.__dict__.get("_num_graphs", int(self._store["ptr"].numel()) - 1)


class Dataset(torch.utils.data.Dataset):
    def __init__(self, root=None, transform=None, pre_transform=None, pre_filter=None):
        self.root = root
        self.transform, self.pre_transform, self.pre_filter = transform, pre_transform, pre_filter
        self._indices = None
        if root is not None and not all(os.path.exists(p) for p in self.processed_paths):
            os.makedirs(self.processed_dir, exist_ok=True)
            self.process()

    @property
    def raw_dir(self):
        return os.path.join(self.root, "raw")

    @property
    def processed_dir(self):
        return os.path.join(self.root, "processed")

    @property
    def raw_paths(self):
        return [os.path.join(self.raw_dir, f) for f in self.raw_file_names]

    @property
    def processed_paths(self):
        return [os.path.join(self.processed_dir, f) for f in self.processed_file_names]

    def indices(self):
        return range(self.len()) if self._indices is None else self._indices

    def __len__(self):
        return len(self.indices())

    def __getitem__(self, idx):
        if isinstance(idx, int) or (torch.is_tensor(idx) and idx.dim() == 0):
            idx = int(idx)
            if idx < 0:
                idx += len(self)
            data = self.get(self.indices()[idx])
            return data if self.transform is None else self.transform(data)
        sub = copy.copy(self)
        if isinstance(idx, slice):
            sub._indices = list(self.indices())[idx]
        else:
            base = list(self.indices())
            sub._indices = [base[int(i)] for i in idx]
        return sub

    def shuffle(self):
        perm = torch.randperm(len(self))
        return self[perm]


class InMemoryDataset(Dataset):
    def __init__(self, root=None, transform=None, pre_transform=None, pre_filter=None):
        self.data, self.slices = None, None
        super().__init__(root, transform, pre_transform, pre_filter)

    @staticmethod
    def collate(data_list):
        keys = data_list[0].keys()
        data, slices = Data(), {}
        for k in keys:
            vals = [d[k] for d in data_list]
            dim = _cat_dim(k)
            data[k] = torch.cat(vals, dim=dim)
            lens = torch.tensor([v.size(dim) for v in vals], dtype=torch.long)
            slices[k] = torch.cat([torch.zeros(1, dtype=torch.long), lens.cumsum(0)])
        return data, slices

    def len(self):
        for v in self.slices.values():
            return v.numel() - 1
        return 0

    def get(self, idx):
        out = Data()
        for k in self.data.keys():
            v, s = self.data[k], self.slices[k]
            lo, hi = int(s[idx]), int(s[idx + 1])
            out[k] = v.narrow(_cat_dim(k), lo, hi - lo)
        return out
