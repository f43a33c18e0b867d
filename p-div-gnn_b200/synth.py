"""Synthetic periodic 2-D RVE meshes (100 x 100 plate with one circular hole).

Replaces, for benchmarks and tests only, the FEM dataset generators of the
reference (scripts/generate_dataset.py:118-193 mesh, :85-105 divergence operator,
:413-464 hole placement, :584-598 on-disk fields), whose gmsh/fedoo stack is not
available here.  It produces *inputs* only (host numpy arrays):

  pos      [N,3] float64, z = 0        (pyvista ``mesh.points`` analogue)
  faces    [3,F] int64                 (``_format_faces_from_pyvista`` output layout)
  labels   [N]   int64 in {-1,0,1}     (``datasets.NodeType``: -1 hole, 0 interior, 1 sides)
  op_div   COO (row, col, data) of an N x 2N nodal-averaged P1 divergence operator
  mean_stress [3], stress_field [N,3]  (``.npz`` keys of generate_dataset.py:588-597)

Opposite sides carry nodes at identical coordinates and the four corners are
exact, so ``compute_periodic_graph`` (datasets.py:39-119) pairs them by ``==``.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import Delaunay
import scipy.sparse as sp

PLATE = 100.0


def _p1_divergence_operator(pos2: np.ndarray, tris: np.ndarray):
    """Nodal (area-weighted) average of the element-wise P1 divergence.

    div = D @ [[sxx, sxy], [sxy, syy]] stacked as [2N, 2]  (gnn_train.py:68-76), so
    D[:, :N] is d/dx and D[:, N:] is d/dy.
    """
    n = pos2.shape[0]
    i, j, k = tris[:, 0], tris[:, 1], tris[:, 2]
    xi, yi = pos2[i, 0], pos2[i, 1]
    xj, yj = pos2[j, 0], pos2[j, 1]
    xk, yk = pos2[k, 0], pos2[k, 1]
    two_a = (xj - xi) * (yk - yi) - (xk - xi) * (yj - yi)  # > 0 (CCW)
    area = 0.5 * two_a
    # grad phi_a = (b_a, c_a) / (2A)
    b = np.stack([yj - yk, yk - yi, yi - yj], axis=1)
    c = np.stack([xk - xj, xi - xk, xj - xi], axis=1)
    node_area = np.zeros(n)
    for a in range(3):
        np.add.at(node_area, tris[:, a], area)
    rows, cols, vals = [], [], []
    for tgt in range(3):  # node receiving the element value
        w = area / node_area[tris[:, tgt]]
        for a in range(3):  # shape function
            rows.append(tris[:, tgt])
            cols.append(tris[:, a])
            vals.append(w * b[:, a] / two_a)
            rows.append(tris[:, tgt])
            cols.append(tris[:, a] + n)
            vals.append(w * c[:, a] / two_a)
    m = sp.coo_matrix(
        (np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
        shape=(n, 2 * n),
    ).tocsr()
    m.sum_duplicates()
    m = m.tocoo()
    return m.row.astype(np.int64), m.col.astype(np.int64), m.data.astype(np.float64)


def make_rve_mesh(seed: int, target_nodes: int = 1024, stress_scale: float = 5.0e3):
    """One plate-with-hole sample.  ``stress_scale`` 5e3 ~ elastic, 3 ~ hyperelastic."""
    rng = np.random.default_rng(seed)
    # hole: centre and radius first (generate_dataset.py:413-464 draws them the same way),
    # then the grid pitch that lands close to ``target_nodes`` after the hole is cut
    lo = 20.0
    centre = rng.uniform(lo, PLATE - lo, size=2)
    dist_edge = min(centre[0], centre[1], PLATE - centre[0], PLATE - centre[1])
    h0 = PLATE / np.sqrt(target_nodes)
    radius = rng.uniform(6.0, max(6.5, dist_edge - max(8.0, 3.0 * h0)))
    solid = 1.0 - np.pi * radius * radius / (PLATE * PLATE)
    n = max(6, int(round(np.sqrt(target_nodes / solid))))
    xs = np.linspace(0.0, PLATE, n)
    h = PLATE / (n - 1)
    gx, gy = np.meshgrid(xs, xs, indexing="xy")
    pts = np.stack([gx.ravel(), gy.ravel()], axis=1)
    ix, iy = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    on_side = ((ix == 0) | (ix == n - 1) | (iy == 0) | (iy == n - 1)).ravel()
    # jitter the strict interior so the Delaunay triangulation is not degenerate
    jitter = rng.uniform(-0.22 * h, 0.22 * h, size=pts.shape)
    ring = ((ix <= 0) | (ix >= n - 1) | (iy <= 0) | (iy >= n - 1)).ravel()
    pts = np.where(ring[:, None], pts, pts + jitter)
    d = np.linalg.norm(pts - centre, axis=1)
    keep = on_side | (d > radius + 0.6 * h)
    pts = pts[keep]
    on_side = on_side[keep]
    m = max(8, int(round(2.0 * np.pi * radius / h)))
    ang = 2.0 * np.pi * (np.arange(m) + rng.uniform(0, 1)) / m
    circ = centre + radius * np.stack([np.cos(ang), np.sin(ang)], axis=1)
    all_pts = np.concatenate([pts, circ], axis=0)
    labels = np.zeros(all_pts.shape[0], dtype=np.int64)
    labels[: pts.shape[0]][on_side] = 1
    labels[pts.shape[0]:] = -1
    tri = Delaunay(all_pts).simplices.astype(np.int64)
    cen = all_pts[tri].mean(axis=1)
    tri = tri[np.linalg.norm(cen - centre, axis=1) > radius * (1.0 - 1e-9)]
    p = all_pts
    two_a = (p[tri[:, 1], 0] - p[tri[:, 0], 0]) * (p[tri[:, 2], 1] - p[tri[:, 0], 1]) - (
        p[tri[:, 2], 0] - p[tri[:, 0], 0]
    ) * (p[tri[:, 1], 1] - p[tri[:, 0], 1])
    tri = tri[np.abs(two_a) > 1e-9 * h * h]
    two_a = two_a[np.abs(two_a) > 1e-9 * h * h]
    flip = two_a < 0
    tri[flip] = tri[flip][:, [0, 2, 1]]
    used = np.zeros(all_pts.shape[0], dtype=bool)
    used[tri.ravel()] = True
    assert used.all(), "synthetic mesh has unused nodes"
    nn = all_pts.shape[0]
    # shuffle node numbering a little?  No: keep the generator order (row-major grid,
    # then the hole ring) -- it is what a structured mesher would emit.
    pos3 = np.concatenate([all_pts, np.zeros((nn, 1))], axis=1)
    row, col, data = _p1_divergence_operator(all_pts, tri)
    mean_stress = rng.uniform(-1.0, 1.0, size=3) * stress_scale
    stress_field = mean_stress[None, :] + 0.3 * stress_scale * rng.standard_normal((nn, 3))
    return {
        "pos": pos3,
        "faces": np.ascontiguousarray(tri.T),
        "labels": labels,
        "op_div_row": row,
        "op_div_col": col,
        "op_div_data": data,
        "op_div_shape": (nn, 2 * nn),
        "mean_stress": mean_stress,
        "stress_field": stress_field,
    }


def make_quad_rve_mesh(seed: int, target_nodes: int = 1024, stress_scale: float = 5.0e3):
    """Same sample layout on an all-QUAD mesh (``faces [4,F]``, the ``_quad_face_to_edge`` path of
    convert_utils.py:52-81): a structured grid of the plate with a rectangular block of cells cut out as the hole.
    Interior nodes are jittered so edge lengths differ; side and hole-boundary nodes stay exact."""
    rng = np.random.default_rng(seed)
    n = max(8, int(round(np.sqrt(target_nodes * 1.08))))  # nodes per side (the hole removes a few per cent)
    h = PLATE / (n - 1)
    nc = n - 1  # cells per side
    w, hh = int(rng.integers(2, max(3, nc // 3))), int(rng.integers(2, max(3, nc // 3)))
    cx0, cy0 = int(rng.integers(2, nc - w - 1)), int(rng.integers(2, nc - hh - 1))  # >= 2 cells from every side
    ix, iy = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    ix, iy = ix.ravel(), iy.ravel()
    inside = (ix > cx0) & (ix < cx0 + w) & (iy > cy0) & (iy < cy0 + hh)  # strictly inside the hole: removed
    on_hole = (ix >= cx0) & (ix <= cx0 + w) & (iy >= cy0) & (iy <= cy0 + hh) & ~inside
    on_side = (ix == 0) | (ix == n - 1) | (iy == 0) | (iy == n - 1)
    xs = np.linspace(0.0, PLATE, n)
    pts = np.stack([xs[ix], xs[iy]], axis=1)
    free = ~(on_side | on_hole | inside)
    pts = np.where(free[:, None], pts + rng.uniform(-0.2 * h, 0.2 * h, size=pts.shape), pts)
    keep = ~inside
    new_id = np.cumsum(keep) - 1
    quads = []
    for cy in range(nc):
        for cx in range(nc):
            if cx0 <= cx < cx0 + w and cy0 <= cy < cy0 + hh:
                continue
            a, b, c, d = cy * n + cx, cy * n + cx + 1, (cy + 1) * n + cx + 1, (cy + 1) * n + cx  # counter-clockwise
            quads.append((new_id[a], new_id[b], new_id[c], new_id[d]))
    quads = np.asarray(quads, dtype=np.int64)
    all_pts = pts[keep]
    nn = all_pts.shape[0]
    labels = np.zeros(nn, dtype=np.int64)
    labels[on_side[keep]] = 1
    labels[on_hole[keep]] = -1
    # divergence operator: P1 on the two triangles of every quad (it is input data of the loss, any N x 2N operator does)
    tris = np.concatenate([quads[:, [0, 1, 2]], quads[:, [0, 2, 3]]], axis=0)
    row, col, data = _p1_divergence_operator(all_pts, tris)
    mean_stress = rng.uniform(-1.0, 1.0, size=3) * stress_scale
    stress_field = mean_stress[None, :] + 0.3 * stress_scale * rng.standard_normal((nn, 3))
    return {
        "pos": np.concatenate([all_pts, np.zeros((nn, 1))], axis=1),
        "faces": np.ascontiguousarray(quads.T),
        "labels": labels,
        "op_div_row": row,
        "op_div_col": col,
        "op_div_data": data,
        "op_div_shape": (nn, 2 * nn),
        "mean_stress": mean_stress,
        "stress_field": stress_field,
    }


def make_dataset(n_meshes: int, target_nodes: int = 1024, seed0: int = 69, stress_scale: float = 5.0e3, quads: bool = False):
    """``seed0 + i`` per mesh (``default_rng(69 + i)``)."""
    make = make_quad_rve_mesh if quads else make_rve_mesh
    return [make(seed0 + i, target_nodes, stress_scale) for i in range(n_meshes)]
