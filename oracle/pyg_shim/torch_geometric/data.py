import copy

import torch

from .utils import coalesce as _coalesce


class Data:
    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    @property
    def num_nodes(self):
        for k in ("x", "pos"):
            v = self.__dict__.get(k)
            if v is not None:
                return v.shape[0]
        return int(self.edge_index.max()) + 1

    @property
    def num_edges(self):
        return self.edge_index.shape[1]

    def coalesce(self):
        ea = self.__dict__.get("edge_attr")
        if ea is None:
            self.edge_index = _coalesce(self.edge_index, None, self.num_nodes)
        else:
            self.edge_index, self.edge_attr = _coalesce(self.edge_index, ea, self.num_nodes)
        return self

    def to(self, device):
        for k in self.keys():
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self

    def clone(self):
        return copy.deepcopy(self)


def _cat_dim(key, value):
    if torch.is_tensor(value) and value.is_sparse and "adj" not in key:
        return 0  # row-stack, columns untouched
    if "index" in key or key == "face":
        return -1
    return 0


class Batch(Data):
    @classmethod
    def from_data_list(cls, data_list):
        b = cls()
        ns = [d.num_nodes for d in data_list]
        ptr = [0]
        for n in ns:
            ptr.append(ptr[-1] + n)
        for key in data_list[0].keys():
            vals = [getattr(d, key) for d in data_list]
            v0 = vals[0]
            if v0 is None:
                setattr(b, key, None)
            elif torch.is_tensor(v0) and v0.is_sparse:
                rows, cols, dat = [], [], []
                for v, off in zip(vals, ptr[:-1]):
                    v = v.coalesce()
                    rows.append(v.indices()[0] + off)
                    cols.append(v.indices()[1])
                    dat.append(v.values())
                width = max(int(v.shape[1]) for v in vals)
                setattr(b, key, torch.sparse_coo_tensor(
                    torch.vstack([torch.cat(rows), torch.cat(cols)]), torch.cat(dat), (ptr[-1], width)).coalesce())
            elif torch.is_tensor(v0):
                if _cat_dim(key, v0) == -1:
                    setattr(b, key, torch.cat([v + off for v, off in zip(vals, ptr[:-1])], dim=-1))
                else:
                    setattr(b, key, torch.cat(vals, dim=0))
            elif isinstance(v0, (bool, int, float)) or hasattr(v0, "dtype"):
                setattr(b, key, torch.tensor([float(v) if not isinstance(v, bool) else v for v in vals]))
            else:
                setattr(b, key, vals)
        b.batch = torch.repeat_interleave(torch.arange(len(ns)), torch.tensor(ns))
        b.ptr = torch.tensor(ptr)
        b._num_graphs = len(ns)
        b._ns = ns
        b._ne = [d.edge_index.shape[1] for d in data_list]
        return b

    @property
    def num_nodes(self):
        return int(self.ptr[-1])

    @property
    def batch_size(self):
        return self._num_graphs

    @property
    def num_graphs(self):
        return self._num_graphs

    def __len__(self):
        return self._num_graphs

    def __getitem__(self, i):
        return self.get_example(i)

    def get_example(self, i):
        s, e = int(self.ptr[i]), int(self.ptr[i + 1])
        es = sum(self._ne[:i])
        ee = es + self._ne[i]
        d = Data()
        for key in self.keys():
            if key in ("batch", "ptr"):
                continue
            v = getattr(self, key)
            if torch.is_tensor(v) and v.is_sparse:
                v = v.coalesce()
                idx, val = v.indices(), v.values()
                sel = (idx[0] >= s) & (idx[0] < e)
                setattr(d, key, torch.sparse_coo_tensor(
                    torch.vstack([idx[0][sel] - s, idx[1][sel]]), val[sel], (e - s, v.shape[1])).coalesce())
            elif torch.is_tensor(v) and v.dim() >= 1 and key == "edge_index":
                setattr(d, key, v[:, es:ee] - s)
            elif torch.is_tensor(v) and key == "edge_attr":
                setattr(d, key, v[es:ee])
            elif torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == self.num_nodes:
                setattr(d, key, v[s:e])
            elif torch.is_tensor(v) and v.dim() == 1 and v.shape[0] == self._num_graphs:
                setattr(d, key, v[i])
        return d


class InMemoryDataset:
    def __init__(self, root=None, transform=None, pre_transform=None, pre_filter=None):
        self.transform = transform

    @staticmethod
    def collate(data_list):
        b = Batch.from_data_list(data_list)
        b._data_list = data_list
        return b, None

    def __len__(self):
        return len(self.data._data_list)

    def __getitem__(self, i):
        return self.data._data_list[i]
