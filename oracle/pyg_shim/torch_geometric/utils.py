import torch


def coalesce(edge_index, edge_attr=None, num_nodes=None):
    n = int(num_nodes if num_nodes is not None else edge_index.max() + 1)
    key = edge_index[0].long() * n + edge_index[1].long()
    uniq, inv = torch.unique(key, sorted=True, return_inverse=True)
    ei = torch.stack([uniq // n, uniq % n], 0)
    if edge_attr is None:
        return ei
    out = torch.zeros(uniq.numel(), *edge_attr.shape[1:], dtype=edge_attr.dtype)
    out.index_add_(0, inv, edge_attr)
    return ei, out


def to_undirected(edge_index, num_nodes=None):
    row, col = edge_index
    both = torch.stack([torch.cat([row, col]), torch.cat([col, row])], 0)
    return coalesce(both, None, num_nodes)
