"""ORACLE ONLY. Minimal torch_geometric.data 2.0.2 surface used by the reference's
batch.py / dataloader.py / sgcn_data.py: `Data` with keys, item access,
__inc__/__cat_dim__/num_nodes, contiguous(), to()."""
import torch
import torch.utils.data


class Data(object):
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kwargs):
        object.__setattr__(self, "_store", {})
        for k, v in dict(x=x, edge_index=edge_index, edge_attr=edge_attr, y=y, pos=pos).items():
            if v is not None:
                self._store[k] = v
        for k, v in kwargs.items():
            self[k] = v

    # attribute / item protocol --------------------------------------------------
    def __getattr__(self, key):
        store = object.__getattribute__(self, "_store")
        if key in store:
            return store[key]
        if key in ("x", "edge_index", "edge_attr", "y", "pos", "batch"):
            return None
        raise AttributeError(key)

    def __setattr__(self, key, value):
        if key.startswith("_"):
            object.__setattr__(self, key, value)
        else:
            self._store[key] = value

    def __getitem__(self, key):
        return self._store.get(key, None)

    def __setitem__(self, key, value):
        self._store[key] = value

    def __contains__(self, key):
        return key in self.keys

    @property
    def keys(self):
        return [k for k, v in self._store.items() if v is not None]

    def __iter__(self):
        for k in sorted(self.keys):
            yield k, self[k]

    # batching rules (PyG 2.0.2 Data.__inc__ / __cat_dim__) ------------------------
    def __cat_dim__(self, key, value, *args, **kwargs):
        return -1 if ("index" in key or "face" in key) else 0

    def __inc__(self, key, value, *args, **kwargs):
        if "batch" in key:
            return int(value.max()) + 1
        if "index" in key or "face" in key:
            return self.num_nodes
        return 0

    @property
    def num_nodes(self):
        if "num_nodes" in self._store:
            return self._store["num_nodes"]
        for k in ("x", "pos", "batch"):
            v = self._store.get(k, None)
            if torch.is_tensor(v):
                return v.size(0)
        ei = self._store.get("edge_index", None)
        if torch.is_tensor(ei) and ei.numel() > 0:
            return int(ei.max()) + 1
        return None

    @property
    def num_edges(self):
        ei = self._store.get("edge_index", None)
        return None if ei is None else ei.size(1)

    @property
    def num_node_features(self):
        x = self._store.get("x", None)
        return 0 if x is None else (1 if x.dim() == 1 else x.size(1))

    num_features = num_node_features

    def _apply(self, fn):
        for k, v in list(self._store.items()):
            if torch.is_tensor(v):
                self._store[k] = fn(v)
        return self

    def contiguous(self):
        return self._apply(lambda t: t.contiguous())

    def to(self, device, *a, **k):
        return self._apply(lambda t: t.to(device, *a, **k))

    def cpu(self):
        return self.to("cpu")

    def debug(self):
        pass


class Dataset(torch.utils.data.Dataset):
    def __init__(self, root=None, transform=None, pre_transform=None, pre_filter=None):
        self.root, self.transform, self.pre_transform, self.pre_filter = root, transform, pre_transform, pre_filter


class InMemoryDataset(Dataset):  # imported by sgcn_data.py / util_gdc.py; not used by the oracle
    pass


class DenseDataLoader(torch.utils.data.DataLoader):  # imported by the train_eval_* modules only
    pass


class DataLoader(torch.utils.data.DataLoader):
    pass


def download_url(*a, **k):
    raise RuntimeError("offline")


def extract_zip(*a, **k):
    raise RuntimeError("offline")
