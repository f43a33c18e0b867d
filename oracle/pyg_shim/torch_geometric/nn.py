import inspect

import torch


class LayerNorm(torch.nn.Module):
    """mode='graph'; called with one positional arg => batch=None."""

    def __init__(self, in_channels, eps=1e-5, affine=True, mode="graph"):
        super().__init__()
        self.in_channels, self.eps, self.mode = in_channels, eps, mode
        self.weight = torch.nn.Parameter(torch.ones(in_channels))
        self.bias = torch.nn.Parameter(torch.zeros(in_channels))

    def forward(self, x, batch=None):
        assert batch is None and self.mode == "graph"
        x = x - x.mean()
        out = x / (x.std(unbiased=False) + self.eps)
        return out * self.weight + self.bias


class MessagePassing(torch.nn.Module):
    """aggr='add', flow='source_to_target', node_dim=-2."""

    def __init__(self, aggr="add"):
        super().__init__()
        assert aggr == "add"

    def propagate(self, edge_index, **kwargs):
        row, col = edge_index[0], edge_index[1]
        n = kwargs["x"].shape[0]
        margs = {}
        for name in inspect.signature(self.message).parameters:
            if name.endswith("_i"):
                margs[name] = kwargs[name[:-2]][col]
            elif name.endswith("_j"):
                margs[name] = kwargs[name[:-2]][row]
            else:
                margs[name] = kwargs[name]
        msg = self.message(**margs)
        index = col.view(-1, 1).expand_as(msg)
        aggr = msg.new_zeros(n, msg.shape[1]).scatter_add_(0, index, msg)
        uargs = {k: kwargs[k] for k in inspect.signature(self.update).parameters if k in kwargs}
        return self.update(aggr, **uargs)


def summary(model, *args, **kwargs):
    hooks, names = [], []
    for name, mod in model.named_modules():
        hooks.append(mod.register_forward_hook(lambda m, i, o, name=name: names.append(name)))
    with torch.no_grad():
        model(*args, **kwargs)
    for h in hooks:
        h.remove()
    return "\n".join(names)
