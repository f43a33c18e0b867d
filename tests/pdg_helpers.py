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


# ---- GPU-side helpers -------------------------------------------------------------------
class DeviceBatch:
    """Duck-typed PyG ``Batch`` on the GPU (attribute names of the reference dataset)."""

    def __init__(self, ob, device="cuda"):
        for k in ("pos", "edge_index", "edge_attr", "mean_stress", "local_stress", "nodes_types",
                  "surfaces_nodes_for_div", "batch", "ptr"):
            setattr(self, k, getattr(ob, k).to(device))
        self.op_div_matrix = ob.op_div_matrix.to(device)
        self.batch_size = ob.batch_size
        self.num_nodes = ob.num_nodes

    def to(self, device):
        return self

    def __len__(self):
        return self.batch_size


def make_model(stats, steps=10, params=None, device="cuda", seed=69):
    import pdivgnn_b200
    if params is None:
        torch.manual_seed(seed)
    m = pdivgnn_b200.EncodeProcessDecode(
        input_edges_features_size=1, message_passing_steps=steps, latent_size=128, input_nodes_features_size=6,
        output_nodes_features_size=3, **{k: v.clone() for k, v in stats.items()})
    if params is not None:
        m.load_state_dict(params)
    return m.to(device)

# order of batcher.dataset_stats() keys
STAT_KEYS_ORDERED = ("mean_pos", "std_pos", "mean_mean_stress", "std_mean_stress", "mean_local_stress",
                     "std_local_stress", "mean_edge_weight", "std_edge_weight")


def two_hole_plate(n=8, holes=((2, 2), (5, 4))):
    """(pos [n*n,3], quad faces [4,F], triangle faces [3,2F], hand-derived labels) of an n x n-node plate of unit quads with
    single-cell holes: side nodes 1, the 4 corners of every removed cell -1, everything else 0 (datasets.py:133-179)."""
    pos = np.array([[x, y, 0.0] for y in range(n) for x in range(n)], dtype=np.float64)
    holes = set(holes)
    quads = [(y * n + x, y * n + x + 1, (y + 1) * n + x + 1, (y + 1) * n + x) for y in range(n - 1) for x in range(n - 1)
             if (x, y) not in holes]
    want = np.zeros(n * n, dtype=np.int64)
    for y in range(n):
        for x in range(n):
            if x in (0, n - 1) or y in (0, n - 1):
                want[y * n + x] = 1
    for (hx, hy) in holes:
        for dx in (0, 1):
            for dy in (0, 1):
                want[(hy + dy) * n + hx + dx] = -1
    tris = [t for q in quads for t in ((q[0], q[1], q[2]), (q[0], q[2], q[3]))]
    return pos, np.array(quads, dtype=np.int64).T, np.array(tris, dtype=np.int64).T, want
