"""Shared test helpers: golden loading and oracle-side batch assembly."""
from __future__ import annotations

import os
from types import SimpleNamespace

import numpy as np
import torch

from oracle import pdg_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STAT_KEYS = ("mean_pos", "std_pos", "mean_mean_stress", "std_mean_stress", "mean_local_stress",
             "std_local_stress", "mean_edge_weight", "std_edge_weight")


def load_golden(name):
    return np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)


def golden_params():
    g = np.load(os.path.join(GOLD, "params_seed69.npz"))
    return {k: torch.from_numpy(g[k]) for k in O.STATE_KEYS}


def golden_samples(g):
    out = []
    for i in range(int(g["n_graphs"])):
        out.append({k: g[f"in{i}_{k}"] for k in ("pos", "faces", "labels", "op_div_row", "op_div_col", "op_div_data",
                                                  "op_div_shape", "mean_stress", "stress_field")})
    return out


def oracle_batch_from_samples(samples, periodic=True):
    graphs = [O.build_graph(s, periodic) for s in samples]
    return graphs, O.collate(graphs), O.dataset_stats(graphs)


def rel_err(a, b):
    """Norm-wise relative errors: (L-inf, L2)."""
    a = torch.as_tensor(a, dtype=torch.float64).flatten()
    b = torch.as_tensor(b, dtype=torch.float64).flatten()
    dinf = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)
    d2 = (a - b).norm().item() / max(b.norm().item(), 1e-30)
    return dinf, d2


def synthetic_batch(n_graphs, target_nodes, seed0=69, periodic=True, stress_scale=5.0e3):
    from pdivgnn_b200 import synth
    samples = synth.make_dataset(n_graphs, target_nodes, seed0, stress_scale)
    return samples, *oracle_batch_from_samples(samples, periodic)
