"""Size-independent properties of the path ("property tests"), driven by hypothesis on the CPU oracle:
symmetry / sortedness / uniqueness of the edge lists, in-degree sums, coalesce as a checksum-preserving merge,
periodic pairing, invariance of the model output under edge re-ordering and node re-labelling, periodicity check
invariances.  The GPU counterparts of the same properties live in tests/test_gpu_forward.py
(test_edge_order_permutation_invariance, test_plan_bit_exact) and tests/test_gpu_batcher.py."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

import pdg_helpers as H
from oracle import pdg_oracle as O

SET = settings(max_examples=30, deadline=None)


def _grid_mesh(nx, ny, rng, quads=False, jitter=0.2):
    """Structured periodic plate: nx x ny nodes on [0, nx-1] x [0, ny-1], interior nodes jittered, triangles or quads."""
    xs, ys = np.meshgrid(np.arange(nx, dtype=np.float64), np.arange(ny, dtype=np.float64), indexing="xy")
    pos = np.stack([xs.ravel(), ys.ravel(), np.zeros(nx * ny)], axis=1)
    interior = (xs.ravel() > 0) & (xs.ravel() < nx - 1) & (ys.ravel() > 0) & (ys.ravel() < ny - 1)
    pos[interior, :2] += rng.uniform(-jitter, jitter, size=(int(interior.sum()), 2))
    idx = lambda i, j: j * nx + i  # noqa: E731
    faces = []
    for j in range(ny - 1):
        for i in range(nx - 1):
            a, b, c, d = idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)
            if quads:
                faces.append((a, b, c, d))
            else:
                faces += [(a, b, c), (a, c, d)]
    return pos, np.asarray(faces, dtype=np.int64).T


@SET
@given(n=st.integers(3, 40), f=st.integers(1, 60), quads=st.booleans(), seed=st.integers(0, 2**31 - 1))
def test_face_to_edge_is_sorted_unique_symmetric(n, f, quads, seed):
    """FaceToEdge / _quad_face_to_edge + to_undirected (convert_utils.py:47-81): for ANY face soup the edge list is
    sorted by (row, col), duplicate-free, symmetric, holds exactly the cell sides, and sum(in-degree) == E."""
    rng = np.random.default_rng(seed)
    k = 4 if quads else 3
    face = torch.from_numpy(rng.integers(0, n, size=(k, f)))
    ei = (O.quad_face_to_edge if quads else O.face_to_edge)(face, n)
    key = ei[0] * n + ei[1]
    assert torch.all(key[1:] > key[:-1])                                  # sorted, unique
    assert set(key.tolist()) == set((ei[1] * n + ei[0]).tolist())         # symmetric
    sides = [(0, 1), (1, 2), (2, 3), (0, 3)] if quads else [(0, 1), (1, 2), (0, 2)]
    want = set()
    for a, b in sides:
        for p, q in zip(face[a].tolist(), face[b].tolist()):
            want.add(p * n + q)
            want.add(q * n + p)
    assert set(key.tolist()) == want
    indeg = torch.bincount(ei[1], minlength=n)
    outdeg = torch.bincount(ei[0], minlength=n)
    assert int(indeg.sum()) == ei.shape[1] and torch.equal(indeg, outdeg)
    assert torch.equal(O.to_undirected(ei, n), ei)                        # idempotent


@SET
@given(n=st.integers(2, 30), e=st.integers(1, 200), seed=st.integers(0, 2**31 - 1))
def test_coalesce_preserves_the_checksum_of_the_weights(n, e, seed):
    """Data.coalesce (datasets.py:119): sorted unique keys, per-key sums, total weight unchanged (integer-valued weights
    so that the sums are exact in fp32)."""
    rng = np.random.default_rng(seed)
    ei = torch.from_numpy(rng.integers(0, n, size=(2, e)))
    w = torch.from_numpy(rng.integers(-8, 9, size=e).astype(np.float32))
    ci, cw = O.coalesce(ei, w, n)
    key = ci[0] * n + ci[1]
    assert torch.all(key[1:] > key[:-1])
    assert float(cw.sum()) == float(w.sum())
    dense = torch.zeros(n, n)
    dense.index_put_((ei[0], ei[1]), w, accumulate=True)
    assert torch.equal(dense[ci[0], ci[1]], cw)
    assert int((dense != 0).sum()) <= ci.shape[1]


@SET
@given(nx=st.integers(3, 9), ny=st.integers(3, 9), quads=st.booleans(), seed=st.integers(0, 2**31 - 1))
def test_periodic_graph_pairs_opposite_sides(nx, ny, quads, seed):
    """compute_periodic_graph (datasets.py:39-119) on a structured periodic plate: the added edges join node (0, j) with
    (nx-1, j), (i, 0) with (i, ny-1) and the diagonal corners, in both directions, with weight 0; mesh edges keep their
    Euclidean length; the result stays sorted, unique and symmetric."""
    rng = np.random.default_rng(seed)
    pos, faces = _grid_mesh(nx, ny, rng, quads)
    n = nx * ny
    p = torch.from_numpy(pos)
    f = torch.from_numpy(faces)
    ei = O.quad_face_to_edge(f, n) if quads else O.face_to_edge(f, n)
    w = O.edge_weights(p, ei).float()
    pi, pw = O.compute_periodic_graph(p, ei, w)
    key = pi[0] * n + pi[1]
    assert torch.all(key[1:] > key[:-1])
    assert set(key.tolist()) == set((pi[1] * n + pi[0]).tolist())
    mesh = set((ei[0] * n + ei[1]).tolist())
    added = {k for k in key.tolist() if k not in mesh}
    idx = lambda i, j: j * nx + i  # noqa: E731
    want = set()
    for j in range(ny):
        want |= {idx(0, j) * n + idx(nx - 1, j), idx(nx - 1, j) * n + idx(0, j)}
    for i in range(nx):
        want |= {idx(i, 0) * n + idx(i, ny - 1), idx(i, ny - 1) * n + idx(i, 0)}
    want |= {idx(0, 0) * n + idx(nx - 1, ny - 1), idx(nx - 1, ny - 1) * n + idx(0, 0),
             idx(0, ny - 1) * n + idx(nx - 1, 0), idx(nx - 1, 0) * n + idx(0, ny - 1)}
    assert added == want - mesh
    wmap = dict(zip(key.tolist(), pw.tolist()))
    assert all(wmap[k] == 0.0 for k in added)
    for k, v in zip((ei[0] * n + ei[1]).tolist(), w.tolist()):
        assert wmap[k] == v
    assert O.is_periodic(pos[:, :2])


@SET
@given(nx=st.integers(3, 8), ny=st.integers(3, 8), seed=st.integers(0, 2**31 - 1))
def test_is_periodic_invariances(nx, ny, seed):
    """is_periodic: unchanged by node re-ordering and by a rigid translation; a side node moved along its side by more
    than tol towards +, or removed, breaks it."""
    rng = np.random.default_rng(seed)
    pos, _ = _grid_mesh(nx, ny, rng)
    p = pos[:, :2]
    assert O.is_periodic(p)
    assert O.is_periodic(p[rng.permutation(len(p))])
    assert O.is_periodic(p + np.array([3.25, -7.5]))  # exactly representable shift: sides stay exact
    right = np.where(p[:, 0] == p[:, 0].max())[0]
    q = p.copy()
    q[right[rng.integers(1, len(right) - 1)], 1] += 1e-3  # an inner right-side node moved up
    assert not O.is_periodic(q)
    assert not O.is_periodic(np.delete(p, right[1], axis=0))


@settings(max_examples=6, deadline=None)
@given(seed=st.integers(0, 2**31 - 1))
def test_model_output_is_invariant_under_edge_reordering_and_node_relabelling(seed):
    """EncodeProcessDecode.forward (models.py:288-326): sum aggregation and graph-wide LayerNorm make the prediction
    independent of the ORDER of the edges, and re-labelling the nodes permutes the prediction rows -- up to fp32
    summation-order noise (measured ~1e-6; bound 2e-5 on the norm-wise relative error)."""
    rng = np.random.default_rng(seed)
    samples, graphs, batch, stats = H.synthetic_batch(2, 120, seed0=int(rng.integers(1, 10_000)))
    sd = O.init_state_dict(seed=69)
    ref = O.forward(sd, batch, stats, 4, True, True)
    e = batch.edge_index.shape[1]
    perm = torch.from_numpy(rng.permutation(e))
    shuffled = O.SimpleNamespace(**vars(batch))
    shuffled.edge_index = batch.edge_index[:, perm]
    shuffled.edge_attr = batch.edge_attr[perm]
    out = O.forward(sd, shuffled, stats, 4, True, True)
    linf, l2 = H.rel_err(out, ref)
    assert linf < 2e-5 and l2 < 2e-5, (linf, l2)
    n = batch.num_nodes
    relabel = torch.from_numpy(rng.permutation(n))          # new id of old node i
    inv = torch.empty_like(relabel)
    inv[relabel] = torch.arange(n)
    moved = O.SimpleNamespace(**vars(batch))
    moved.edge_index = relabel[batch.edge_index]
    for k in ("pos", "mean_stress", "nodes_types", "local_stress", "surfaces_nodes_for_div"):
        setattr(moved, k, getattr(batch, k)[inv])
    out2 = O.forward(sd, moved, stats, 4, True, True)
    linf, l2 = H.rel_err(out2[relabel], ref)
    assert linf < 2e-5 and l2 < 2e-5, (linf, l2)


@SET
@given(sizes=st.lists(st.integers(3, 6), min_size=1, max_size=4), seed=st.integers(0, 2**31 - 1))
def test_collation_offsets_and_ptr(sizes, seed):
    """Batch.from_data_list: ptr = prefix sums of the node counts, batch = graph id per node, every edge
    stays inside its graph's id range and, shifted back, equals the graph's own edge list."""
    rng = np.random.default_rng(seed)
    graphs = []
    for k, nx in enumerate(sizes):
        pos, faces = _grid_mesh(nx, nx, rng)
        n = nx * nx
        sample = dict(pos=pos, faces=faces, stress_field=rng.normal(size=(n, 3)).astype(np.float32),
                      mean_stress=rng.normal(size=3), labels=np.zeros(n, dtype=np.int64),
                      op_div_row=np.array([0]), op_div_col=np.array([0]), op_div_data=np.array([1.0]), op_div_shape=(n, 2 * n))
        graphs.append(O.build_graph(sample, periodic=True))
    b = O.collate(graphs)
    ns = [g.num_nodes for g in graphs]
    assert b.ptr.tolist() == [0] + list(np.cumsum(ns)) and b.num_nodes == sum(ns) and b.batch_size == len(graphs)
    assert b.batch.tolist() == [i for i, m in enumerate(ns) for _ in range(m)]
    off = 0
    for g, lo, hi in zip(graphs, b.ptr[:-1].tolist(), b.ptr[1:].tolist()):
        e = g.edge_index.shape[1]
        part = b.edge_index[:, off:off + e]
        assert int(part.min()) >= lo and int(part.max()) < hi
        assert torch.equal(part - lo, g.edge_index)
        off += e
    assert off == b.edge_index.shape[1]
