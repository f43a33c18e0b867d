"""Generate tests/golden/*.npz from the reference's UNMODIFIED modules.

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference):

    python oracle/make_golden.py

The reference (gnn_local_stress/{models,datasets,data_utils,convert_utils}.py and
scripts/gnn_train.py) is imported as-is; ``torch_geometric`` is provided by the
test-only shim ``oracle/pyg_shim`` and pyvista/fedoo/fire by import-only stubs
(``oracle/stubs``).  The dataset class reads its meshes through a fake
``pyvista.get_reader`` that serves in-memory synthetic meshes, and its ``.npz``
side files from a temp dir, so ``MeshStressFieldDatasetInMemory.__init__`` runs
unmodified (mesh_to_graph, edge weights, periodic edges, op_div, stats, collate).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import tempfile
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pandas as pd
import torch

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT / "oracle" / "pyg_shim"))
sys.path.insert(0, str(ROOT / "oracle" / "stubs"))
sys.path.insert(0, str(REF))
sys.path.insert(0, str(ROOT))

import pyvista as pv  # noqa: E402  (stub)
from gnn_local_stress import data_utils, datasets, models  # noqa: E402  (reference)

spec = importlib.util.spec_from_file_location("ref_gnn_train", REF / "scripts" / "gnn_train.py")
ref_train = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_train)

from pdivgnn_b200 import synth  # noqa: E402

GOLD = ROOT / "tests" / "golden"
_MESHES: dict[str, SimpleNamespace] = {}


class _FakeMesh:
    def __init__(self, pos, faces):
        self.points = pos
        f = faces.T
        self.faces = np.concatenate([np.full((f.shape[0], 1), f.shape[1], dtype=np.int64), f], axis=1).ravel()
        self._quad = f.shape[1] == 4

    def get_cell(self, i):
        return SimpleNamespace(type=pv.CellType.QUAD if self._quad else 5)


pv.get_reader = lambda fn: SimpleNamespace(read=lambda: _MESHES[fn])


def _reference_dataset(samples, periodic, tmp):
    rows = []
    for i, s in enumerate(samples):
        mf = f"mem://mesh_{id(samples)}_{i}.vtk"
        _MESHES[mf] = _FakeMesh(s["pos"], s["faces"])
        df = os.path.join(tmp, f"d_{id(samples)}_{i}.npz")
        np.savez(df, stress_field=s["stress_field"], mean_stress=s["mean_stress"],
                 op_div_matrix_data=s["op_div_data"], op_div_matrix_row_indices=s["op_div_row"],
                 op_div_matrix_col_indices=s["op_div_col"], op_div_matrix_shape=np.array(s["op_div_shape"]),
                 node_labels=s["labels"])
        rows.append(dict(mesh_filename=mf, data_filename=df))
    return datasets.MeshStressFieldDatasetInMemory(pd.DataFrame(rows), periodic_graph=periodic)


def _reference_model(ds, seed=69, steps=10, latent=128):
    torch.manual_seed(seed)
    return models.EncodeProcessDecode(
        input_edges_features_size=1, input_nodes_features_size=6, message_passing_steps=steps,
        latent_size=latent, output_nodes_features_size=3,
        mean_pos=ds.mean_pos, std_pos=ds.std_pos, mean_mean_stress=ds.mean_mean_stress,
        std_mean_stress=ds.std_mean_stress, mean_local_stress=ds.mean_local_stress,
        std_local_stress=ds.std_local_stress, mean_edge_weight=ds.mean_edge_weight,
        std_edge_weight=ds.std_edge_weight)


def _reference_train_loss(model, batch, divergence, penalty):
    """The body of train() (gnn_train.py:154-202) with the reference's own functions."""
    pred = model.forward(batch, scale_output=False, scale_input=True).local_stress
    gt = data_utils.standardize(batch.local_stress, model.mean_local_stress, model.std_local_stress)
    batch.local_stress = gt
    batch_loss, batch_div = 0, 0
    for sample, pred_i in data_utils.slice_batch_gt_and_predictions(batch, pred):
        batch_loss = batch_loss + ref_train.normalized_mse_loss_single(
            ground_truth_local_stress=sample.local_stress, predicted_local_stress=pred_i)
        if divergence:
            batch_div = batch_div + ref_train.compute_divergence(
                pred_i, sample.op_div_matrix, sample.surfaces_nodes_for_div, reduce_strategy="square") * penalty
    batch_loss = batch_loss / batch.batch_size
    nmse = batch_loss.detach().clone()
    div = torch.zeros(())
    if divergence:
        batch_div = batch_div / batch.batch_size
        batch_loss = batch_loss + batch_div
        div = batch_div.detach().clone()
    return batch_loss, nmse, div, pred


def _sample_arrays(prefix, s):
    return {f"{prefix}_{k}": np.asarray(v) for k, v in s.items()}


def case_grid3x3(tmp):
    """SURVEY 2.4 hand-checkable vector, produced by the reference's own code."""
    n = 3
    ix, iy = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    pos = np.stack([ix.ravel(), iy.ravel(), np.zeros(n * n)], axis=1).astype(np.float64)
    tris = []
    for y in range(n - 1):
        for x in range(n - 1):
            a, b, c, d = y * n + x, y * n + x + 1, (y + 1) * n + x, (y + 1) * n + x + 1
            tris += [(a, b, d), (a, d, c)]
    faces = np.array(tris, dtype=np.int64).T
    from gnn_local_stress.convert_utils import mesh_to_graph
    g = mesh_to_graph(_FakeMesh(pos, faces))
    g.edge_attr = datasets._compute_node_distances_as_edge_weights(g).float()
    mesh_ei = g.edge_index.clone()
    mesh_ea = g.edge_attr.clone()
    pg = datasets.compute_periodic_graph(g)
    np.savez(GOLD / "grid3x3.npz", pos=pos, faces=faces, mesh_edge_index=mesh_ei.numpy(),
             mesh_edge_attr=mesh_ea.numpy(), edge_index=pg.edge_index.numpy(), edge_attr=pg.edge_attr.numpy())
    print("grid3x3: mesh edges", mesh_ei.shape[1], "periodic total", pg.edge_index.shape[1])


def case_model(name, samples, periodic, divergence, penalty, tmp, also_scaled=True, store_grads=True):
    ds = _reference_dataset(samples, periodic, tmp)
    from torch_geometric.loader import DataLoader
    batch = next(iter(DataLoader(ds, batch_size=len(samples), shuffle=False)))
    model = _reference_model(ds)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    out = {}
    for i, s in enumerate(samples):
        out.update(_sample_arrays(f"in{i}", s))
    out["n_graphs"] = np.array(len(samples))
    out["periodic"] = np.array(periodic)
    out["divergence"] = np.array(divergence)
    out["penalty"] = np.array(penalty)
    for k in ("mean_pos", "std_pos", "mean_mean_stress", "std_mean_stress", "mean_local_stress",
              "std_local_stress", "mean_edge_weight", "std_edge_weight"):
        out["stat_" + k] = getattr(ds, k).numpy()
    out["edge_index"] = batch.edge_index.numpy()
    out["edge_attr"] = batch.edge_attr.numpy()
    out["ptr"] = batch.ptr.numpy()
    op = batch.op_div_matrix.coalesce()
    out["op_indices"] = op.indices().numpy()
    out["op_values"] = op.values().numpy()
    out["op_shape"] = np.array(op.shape)
    ppath = GOLD / "params_seed69.npz"
    if not ppath.exists():
        np.savez_compressed(ppath, **{k: v.numpy() for k, v in sd.items()})
    else:  # every case uses the same seeded default init
        ref = np.load(ppath)
        assert all(np.array_equal(ref[k], v.numpy()) for k, v in sd.items())
    with torch.no_grad():
        if also_scaled:
            out["pred_scaled"] = model.forward(batch, scale_output=True, scale_input=True).local_stress.numpy()
        out["summary"] = np.array(models.PyG.nn.summary(model, batch))
    model.zero_grad()
    loss, nmse, div, pred = _reference_train_loss(model, batch, divergence, penalty)
    loss.backward()
    out["pred_std"] = pred.detach().numpy()
    out["loss"], out["nmse"], out["div"] = loss.detach().numpy(), nmse.numpy(), div.numpy()
    if store_grads:
        for k, p in model.named_parameters():
            out["grad_" + k] = p.grad.numpy() if p.grad is not None else np.zeros_like(p.detach().numpy())
    np.savez_compressed(GOLD / f"{name}.npz", **out)
    print(f"{name}: N={batch.num_nodes} E={batch.edge_index.shape[1]} loss={float(loss):.6f} "
          f"nmse={float(nmse):.6f} div={float(div):.6f}")


def case_quad_grid(tmp):
    """2 x 3 cells of quads through the reference's own ``mesh_to_graph`` -> ``_quad_face_to_edge``
    (convert_utils.py:52-81), edge weights and periodic edges: a hand-checkable vector for the quad path."""
    nx, ny = 3, 4  # nodes
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    pos = np.stack([ix.ravel() * 1.5, iy.ravel() * 1.0, np.zeros(nx * ny)], axis=1).astype(np.float64)
    quads = []
    for y in range(ny - 1):
        for x in range(nx - 1):
            a = y * nx + x
            quads.append((a, a + 1, a + nx + 1, a + nx))
    faces = np.array(quads, dtype=np.int64).T
    from gnn_local_stress.convert_utils import mesh_to_graph
    g = mesh_to_graph(_FakeMesh(pos, faces))
    g.edge_attr = datasets._compute_node_distances_as_edge_weights(g).float()
    mesh_ei, mesh_ea = g.edge_index.clone(), g.edge_attr.clone()
    pg = datasets.compute_periodic_graph(g)
    np.savez(GOLD / "quad_grid.npz", pos=pos, faces=faces, mesh_edge_index=mesh_ei.numpy(), mesh_edge_attr=mesh_ea.numpy(),
             edge_index=pg.edge_index.numpy(), edge_attr=pg.edge_attr.numpy())
    print("quad_grid: mesh edges", mesh_ei.shape[1], "periodic total", pg.edge_index.shape[1])


def main():
    """``python oracle/make_golden.py [case ...]``: no argument = every case; existing files of other cases stay."""
    GOLD.mkdir(parents=True, exist_ok=True)
    want = set(sys.argv[1:])
    on = lambda name: not want or name in want  # noqa: E731
    with tempfile.TemporaryDirectory() as tmp:
        if on("grid3x3"):
            case_grid3x3(tmp)
        if on("quad_grid"):
            case_quad_grid(tmp)
        s2 = [synth.make_rve_mesh(69, 90), synth.make_rve_mesh(70, 140)]
        if on("train2_div"):
            case_model("train2_div", s2, True, True, 10.0, tmp)
        if on("train2_nodiv"):
            case_model("train2_nodiv", s2, True, False, 10.0, tmp, also_scaled=False)
        if on("train3_noperiodic"):
            case_model("train3_noperiodic", [synth.make_rve_mesh(71 + i, 100, 3.0) for i in range(3)], False, True, 10.0, tmp)
        if on("infer1"):
            case_model("infer1", [synth.make_rve_mesh(80, 260)], True, True, 10.0, tmp, store_grads=False)
        if on("train2_quad"):  # all-quad meshes: _quad_face_to_edge path of mesh_to_graph, then the same pipeline
            case_model("train2_quad", [synth.make_quad_rve_mesh(69, 100), synth.make_quad_rve_mesh(72, 170)], True, True, 10.0, tmp)


if __name__ == "__main__":
    main()
