"""Drop-in ``EncodeProcessDecode`` backed by the sm_100a CUDA library.

Mirrors ``gnn_local_stress/models.py`` of the reference: same constructor arguments
(models.py:98-139, 246-259), ``forward(mesh_graph, scale_output, scale_input)``
(:288-326), ``.to(device)`` moving the 8 dataset statistics (:164-179), the exact
``state_dict`` layout (28 tensors / 167 299 elements) and the checkpoint helpers
(:44-95).  The sub-modules exist to own the parameters under the reference's names;
all arithmetic happens in ``libpdivgnn.so`` through :mod:`autograd`.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch.nn import Linear, Sequential

from . import _lib
from .autograd import epd_forward

try:  # the reference returns torch_geometric Data objects; keep that when PyG exists
    from torch_geometric.data import Data as _Data  # type: ignore
except Exception:  # pragma: no cover - PyG is not installed in the build image
    class _Data:  # minimal attribute container with the PyG calling convention
        def __init__(self, **kw):
            for k, v in kw.items():
                setattr(self, k, v)

        def to(self, device):
            for k, v in list(vars(self).items()):
                if torch.is_tensor(v):
                    setattr(self, k, v.to(device))
            return self


class GraphLayerNorm(torch.nn.Module):
    """Parameter holder for PyG ``LayerNorm(mode="graph")`` (models.py:199,207,265,273).

    One mean / one population std over ALL elements of the batched tensor, eps added to
    the std, then per-channel affine.  The normalisation itself is fused into the CUDA
    kernels (lazy application in the consumer); this module is never called.
    """

    def __init__(self, in_channels: int, eps: float = 1e-5):
        super().__init__()
        self.in_channels, self.eps = in_channels, eps
        self.weight = torch.nn.Parameter(torch.ones(in_channels))
        self.bias = torch.nn.Parameter(torch.zeros(in_channels))

    def forward(self, x):  # pragma: no cover
        raise RuntimeError("GraphLayerNorm is fused into libpdivgnn kernels and is not callable on its own")

    def extra_repr(self):
        return f"{self.in_channels}, mode=graph"


def _mlp_ln(fin: int, latent: int) -> Sequential:
    return Sequential(Linear(fin, latent), torch.nn.ReLU(), Linear(latent, latent), torch.nn.ReLU(),
                      GraphLayerNorm(latent))


class Processor(torch.nn.Module):
    """Parameters of the shared message-passing block (models.py:182-208)."""

    def __init__(self, latent_size: int, input_nodes_features_size: int, input_edges_features_size: int):
        super().__init__()
        self.latent_size = latent_size
        self.edge_net = _mlp_ln(input_edges_features_size, latent_size)
        self.node_net = _mlp_ln(input_nodes_features_size, latent_size)

    def forward(self, graph):  # pragma: no cover
        raise RuntimeError("Processor steps run inside libpdivgnn (pdg_forward); call the model instead")


_STAT_ATTRS = ["mean_local_stress", "std_local_stress", "mean_mean_stress", "std_mean_stress", "mean_pos",
               "std_pos", "mean_edge_weight", "std_edge_weight"]


class StressFieldBaseModel(torch.nn.Module):
    """models.py:98-179 -- statistics are plain attributes (not buffers), default ``torch.Tensor(1)``."""

    def __init__(self, latent_size: int, input_nodes_features_size: int, output_nodes_features_size: int,
                 mean_pos=None, std_pos=None, mean_mean_stress=None, std_mean_stress=None, mean_local_stress=None,
                 std_local_stress=None, mean_edge_weight=None, std_edge_weight=None):
        super().__init__()
        self.latent_size = latent_size
        self.input_nodes_features_size = input_nodes_features_size
        self.output_nodes_features_size = output_nodes_features_size
        d = lambda v: torch.Tensor(1) if v is None else v  # noqa: E731  (reference default: uninitialised)
        self.mean_pos, self.std_pos = d(mean_pos), d(std_pos)
        self.mean_mean_stress, self.std_mean_stress = d(mean_mean_stress), d(std_mean_stress)
        self.mean_local_stress, self.std_local_stress = d(mean_local_stress), d(std_local_stress)
        self.mean_edge_weight, self.std_edge_weight = d(mean_edge_weight), d(std_edge_weight)
        self._norm_cache = None

    def to(self, device):
        for attr in _STAT_ATTRS:
            v = getattr(self, attr)
            if v is not None and torch.is_tensor(v):
                setattr(self, attr, v.to(device))
        return super().to(device)

    def _norm_struct(self) -> _lib.PdgNorm:
        """Host copy of the 8 scalars, refreshed only when an attribute object changes."""
        vals = [getattr(self, a) for a in _STAT_ATTRS]
        key = tuple((id(v), v._version) if torch.is_tensor(v) else v for v in vals)
        if self._norm_cache is None or self._norm_cache[0] != key:
            # The reference broadcasts whatever it is given ((pos - mean_pos) / std_pos, models.py:140-162), so a
            # per-channel statistic would silently change meaning here: pdg_norm_t carries SCALARS (what the dataset
            # class computes, datasets.py:283-291).  Anything with more than one element is rejected.
            for a, v in zip(_STAT_ATTRS, vals):
                n = v.numel() if torch.is_tensor(v) else (len(v) if isinstance(v, (tuple, list)) else 1)
                if n != 1:
                    raise NotImplementedError(
                        f"pdivgnn_b200: statistic `{a}` has {n} elements; only the scalar dataset statistics of "
                        "datasets.py:283-291 are supported (per-channel statistics are not)")
            tens = [v for v in vals if torch.is_tensor(v)]
            if len(tens) == len(vals) and len({t.device for t in tens}) == 1:
                # one device->host copy for all 8 scalars instead of 8 synchronising reads
                flat = torch.stack([t.detach().reshape(()).to(torch.float64) for t in tens]).tolist()
            else:
                flat = [float(v.reshape(-1)[0]) if torch.is_tensor(v) else
                        (float(v[0]) if isinstance(v, (tuple, list)) else float(v)) for v in vals]
            s = _lib.PdgNorm(**dict(zip(_STAT_ATTRS, flat)))
            self._norm_cache = (key, s, vals)  # vals kept alive so ids stay unique
        return self._norm_cache[1]


class EncodeProcessDecode(StressFieldBaseModel):
    """models.py:246-326."""

    def __init__(self, input_edges_features_size: int, message_passing_steps: int, *args, precision: str = "fp32",
                 cuda_graphs: bool = False, **kwargs):
        super().__init__(*args, **kwargs)
        self.message_passing_steps = message_passing_steps
        self.input_edges_features_size = input_edges_features_size
        if (self.latent_size, self.input_nodes_features_size, self.input_edges_features_size,
                self.output_nodes_features_size) != (128, 6, 1, 3):
            raise NotImplementedError(
                "libpdivgnn is built for latent 128, node/edge/output features 6/1/3 (every shipped config); got "
                f"{(self.latent_size, self.input_nodes_features_size, self.input_edges_features_size, self.output_nodes_features_size)}")
        # "fp32": the 1e-5 mode (FFMA tiles).  "bf16" (north-star name) == "fp16" == "tc16": the 16-bit tensor-core tile
        # mode (tolerance 2e-2; since round 2 its operand tiles are fp16)
        if precision not in ("fp32", "bf16", "fp16", "tc16"):
            raise ValueError("precision must be 'fp32' or 'bf16' (aliases of the 16-bit tile mode: 'fp16', 'tc16')")
        self.precision = precision
        self.cuda_graphs = cuda_graphs  # replay no-grad forwards on unchanged input tensors as one CUDA-graph launch
        # construction order == reference (same RNG stream => same default init)
        self.node_encoder = _mlp_ln(self.input_nodes_features_size, self.latent_size)
        self.edge_encoder = _mlp_ln(self.input_edges_features_size, self.latent_size)
        self.processor = Processor(self.latent_size, input_nodes_features_size=self.latent_size * 2,
                                   input_edges_features_size=self.latent_size * 3)
        self.node_decoder = Sequential(Linear(self.latent_size, self.latent_size), torch.nn.ReLU(),
                                       Linear(self.latent_size, self.output_nodes_features_size))

    def forward(self, mesh_graph, scale_output: bool = True, scale_input: bool = True):
        # models.py:294-299: an all-zero load case returns zeros.  The reference decides on the host
        # (`torch.any` + a device->host sync per call); here the node encoder raises a device flag while it reads
        # mean_stress and the decoder (and its backward) are predicated on it (PDG_FLAG_ZERO_CHECK): same values,
        # no sync.  `skip_zero_check = True` (caller guarantees a non-zero mean stress) drops the predicate too.
        out = epd_forward(self, mesh_graph, scale_output, scale_input,
                          zero_check=not getattr(self, "skip_zero_check", False))
        return _Data(local_stress=out, edge_index=mesh_graph.edge_index, pos=mesh_graph.pos)


# ---- checkpoint helpers (models.py:33-95), same dict layout --------------------------------

def print_model(model: torch.nn.Module, data_loader, device: str) -> str:
    sample = next(iter(data_loader)).to(device)
    model = model.to(device)
    try:
        import torch_geometric as PyG  # type: ignore
        return PyG.nn.summary(model, sample)
    except ImportError:
        with torch.no_grad():
            model(sample)
        return str(model)


def save_model_checkpoint(model, optimizer, epoch: int, filename: str) -> None:
    checkpoint = {"model_state_dict": model.state_dict(), "optimizer_state_dict": optimizer.state_dict(),
                  "epoch": epoch}
    for a in ("mean_pos", "mean_mean_stress", "std_mean_stress", "mean_local_stress", "std_pos", "std_local_stress",
              "mean_edge_weight", "std_edge_weight"):
        checkpoint[a] = getattr(model, a)
    torch.save(checkpoint, filename)


def load_model_checkpoint(model, filename: str, optimizer: Optional[torch.optim.Optimizer] = None) -> int:
    if not torch.cuda.is_available():
        checkpoint = torch.load(filename, map_location=torch.device("cpu"))
    else:
        checkpoint = torch.load(filename)
    model.load_state_dict(checkpoint["model_state_dict"])
    if optimizer:
        optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
    for a in _STAT_ATTRS:
        setattr(model, a, checkpoint[a])
    return checkpoint["epoch"]


def load_optimizer_checkpoint(optimizer, filename: str):
    checkpoint = torch.load(filename)
    optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
    return optimizer
