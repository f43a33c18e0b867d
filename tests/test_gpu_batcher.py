"""Device graph batcher (pdg_batch_*) vs the oracle's graph construction + collation:
bit-exact edge_index (int64) and edge_attr (fp32), periodic and non-periodic."""
import numpy as np
import pytest
import torch

import pdg_helpers as H
from oracle import pdg_oracle as O

pytestmark = pytest.mark.gpu


def _device_batch(samples, periodic):
    from pdivgnn_b200 import batcher
    return batcher.batch_from_host(batcher.host_arrays(samples), "cuda", periodic)


@pytest.mark.parametrize("name", ["train2_div", "train3_noperiodic", "infer1", "train2_quad"])
def test_golden_edges_bit_exact(name):
    g = H.load_golden(name)
    mb = _device_batch(H.golden_samples(g), bool(g["periodic"]))
    assert np.array_equal(mb.edge_index.cpu().numpy(), g["edge_index"])
    assert np.array_equal(mb.edge_attr.cpu().numpy(), g["edge_attr"])
    assert np.array_equal(mb.ptr.cpu().numpy(), g["ptr"])
    op = mb.op_div_matrix
    assert np.array_equal(op.indices().cpu().numpy(), g["op_indices"])
    assert np.array_equal(op.values().cpu().numpy(), g["op_values"])
    assert tuple(op.shape) == tuple(g["op_shape"])


def test_grid3x3_known_answer():
    g = H.load_golden("grid3x3")
    from pdivgnn_b200 import batcher
    pos = torch.from_numpy(g["pos"][:, :2]).cuda()
    faces = torch.from_numpy(g["faces"]).cuda()
    nptr = torch.tensor([0, 9]).cuda()
    fptr = torch.tensor([0, faces.shape[1]]).cuda()
    ei, ea = batcher.build_edges(pos, faces, nptr, fptr, periodic=True)
    assert ei.shape[1] == 48
    assert np.array_equal(ei.cpu().numpy(), g["edge_index"]) and np.array_equal(ea.cpu().numpy(), g["edge_attr"])
    ei, ea = batcher.build_edges(pos, faces, nptr, fptr, periodic=False)
    assert np.array_equal(ei.cpu().numpy(), g["mesh_edge_index"]) and np.array_equal(ea.cpu().numpy(), g["mesh_edge_attr"])


def test_quad_grid_known_answer():
    """[4,F] faces: the reference's own _quad_face_to_edge output (convert_utils.py:62-81), bit for bit."""
    g = H.load_golden("quad_grid")
    from pdivgnn_b200 import batcher
    pos = torch.from_numpy(np.ascontiguousarray(g["pos"][:, :2])).cuda()
    faces = torch.from_numpy(g["faces"]).cuda()
    nptr, fptr = torch.tensor([0, 12]).cuda(), torch.tensor([0, faces.shape[1]]).cuda()
    ei, ea = batcher.build_edges(pos, faces, nptr, fptr, periodic=False)
    assert ei.shape[1] == 34
    assert np.array_equal(ei.cpu().numpy(), g["mesh_edge_index"]) and np.array_equal(ea.cpu().numpy(), g["mesh_edge_attr"])
    ei, ea = batcher.build_edges(pos, faces, nptr, fptr, periodic=True)
    assert np.array_equal(ei.cpu().numpy(), g["edge_index"]) and np.array_equal(ea.cpu().numpy(), g["edge_attr"])
    lab, reg = batcher.node_labels(pos, faces, nptr, fptr)
    assert lab.tolist() == [1, 1, 1, 1, 0, 1, 1, 0, 1, 1, 1, 1] and reg.tolist() == [1]
    with pytest.raises(NotImplementedError, match="faces must be"):
        batcher.build_edges(pos, faces[:2], nptr, fptr)


@pytest.mark.parametrize("quads", [False, True])
@pytest.mark.parametrize("periodic", [True, False])
def test_synthetic_batch_bit_exact_and_properties(periodic, quads):
    from pdivgnn_b200 import synth
    samples = synth.make_dataset(6, 700, 300, quads=quads)
    graphs, batch, stats = H.oracle_batch_from_samples(samples, periodic)
    mb = _device_batch(samples, periodic)
    ei = mb.edge_index.cpu()
    assert torch.equal(ei, batch.edge_index) and torch.equal(mb.edge_attr.cpu(), batch.edge_attr)
    assert torch.equal(mb.batch.cpu(), batch.batch) and torch.equal(mb.pos.cpu(), batch.pos)
    assert torch.equal(mb.mean_stress.cpu(), batch.mean_stress) and torch.equal(mb.local_stress.cpu(), batch.local_stress)
    assert torch.equal(mb.nodes_types.cpu(), batch.nodes_types)
    n = batch.num_nodes
    key = ei[0] * n + ei[1]
    assert torch.all(key[1:] > key[:-1]), "sorted by (row, col), unique"
    assert torch.equal(torch.sort(ei[1] * n + ei[0]).values, key), "symmetric"
    st = H.rel_err(torch.stack(list(__import__("pdivgnn_b200.batcher", fromlist=["x"]).dataset_stats([mb]).values())).cpu(),
                   torch.stack([stats[k] for k in H.STAT_KEYS_ORDERED]))
    assert st[0] < 1e-6


def test_non_periodic_mesh_is_rejected():
    from pdivgnn_b200 import batcher
    pos = torch.tensor([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.5]], dtype=torch.float64).cuda()
    faces = torch.tensor([[0, 1], [1, 3], [2, 2]]).cuda()
    with pytest.raises(RuntimeError, match="periodic"):
        batcher.build_edges(pos, faces, torch.tensor([0, 4]).cuda(), torch.tensor([0, 2]).cuda(), periodic=True)


def _check_labels(samples):
    from pdivgnn_b200 import batcher
    h = batcher.host_arrays(samples)
    pos, faces = h["pos64"].cuda(), h["faces"].cuda()
    labels, regions = batcher.node_labels(pos, faces, h["node_ptr"].cuda(), h["face_ptr"].cuda(), check_regions=True)
    ref = np.concatenate([O.compute_node_labels(s["pos"], s["faces"])[0] for s in samples])
    assert labels.dtype == torch.int64 and np.array_equal(labels.cpu().numpy(), ref)
    assert np.array_equal(labels.cpu().numpy(), h["labels"].numpy())  # == the generator's ground truth
    assert regions.tolist() == [2] * len(samples)


def test_device_node_labels_match_the_oracle():
    """pdg_node_labels vs the restated datasets.compute_node_labels (datasets.py:133-179): bit-exact labels."""
    from pdivgnn_b200 import batcher, synth
    _check_labels(synth.make_dataset(5, 400, 123) + [synth.make_rve_mesh(9, 2000)])
    _check_labels(synth.make_dataset(5, 400, 123, quads=True) + [synth.make_quad_rve_mesh(9, 2000)])
    # a mesh without a hole has one loop: every boundary node is external, and the reference's assert fires
    g = H.load_golden("grid3x3")
    p = torch.from_numpy(np.ascontiguousarray(g["pos"][:, :2])).cuda()
    f = torch.from_numpy(g["faces"]).cuda()
    nptr, fptr = torch.tensor([0, 9]).cuda(), torch.tensor([0, f.shape[1]]).cuda()
    lab, reg = batcher.node_labels(p, f, nptr, fptr)
    assert lab.tolist() == [1, 1, 1, 1, 0, 1, 1, 1, 1] and reg.tolist() == [1]
    with pytest.raises(AssertionError, match="Expected 2 regions"):
        batcher.node_labels(p, f, nptr, fptr, check_regions=True)


def test_device_node_labels_two_holes_hand_derived():
    """Hand-derived labels of a plate with two holes (tests/pdg_helpers.two_hole_plate), quads and triangles, batched."""
    from pdivgnn_b200 import batcher
    for which in (1, 2):
        pos, *faces, want = H.two_hole_plate()
        f = faces[which - 1]
        p1, _, _, want1 = H.two_hole_plate(holes=((3, 3),))
        f1 = H.two_hole_plate(holes=((3, 3),))[which]
        posb = torch.from_numpy(np.concatenate([pos[:, :2], p1[:, :2]])).cuda()
        fb = torch.from_numpy(np.concatenate([f, f1], axis=1)).cuda()
        nptr = torch.tensor([0, 64, 128]).cuda()
        fptr = torch.tensor([0, f.shape[1], f.shape[1] + f1.shape[1]]).cuda()
        lab, reg = batcher.node_labels(posb, fb, nptr, fptr)
        assert reg.tolist() == [3, 2]
        assert np.array_equal(lab.cpu().numpy(), np.concatenate([want, want1]))


def test_device_is_periodic_matches_the_oracle():
    """pdg_is_periodic vs the restated microgen.mesh.is_periodic (generate_dataset.py:191, benchmark_gnn_fem.py:195) on a
    ragged batch: periodic meshes, a shifted side node, a missing partner, the one-sided tolerance, loose tolerances."""
    from pdivgnn_b200 import batcher, synth
    rng = np.random.default_rng(5)
    meshes = [s["pos"][:, :2].astype(np.float64) for s in synth.make_dataset(4, 400, 31) + [synth.make_rve_mesh(3, 3000)]
              + synth.make_dataset(2, 300, 8, quads=True)]
    variants = []
    for k, p in enumerate(meshes):
        variants.append(p)
        q = p.copy()
        right = np.where(q[:, 0] == q[:, 0].max())[0]
        left = np.where(q[:, 0] == q[:, 0].min())[0]
        top = np.where(q[:, 1] == q[:, 1].max())[0]
        i = right[len(right) // 2]
        if k % 4 == 0:
            q[i, 1] += 1e-3          # partner higher than tol -> not periodic
        elif k % 4 == 1:
            q[i, 1] -= 1e-3          # partner lower: passes the published one-sided test
        elif k % 4 == 2:
            q[left[1], 0] += 1e-6    # leaves the side (|x - xmin| >= tol): side counts differ
        else:
            q[top[2], 0] += 3e-9     # inside tol: still periodic
        variants.append(q)
        variants.append(p[rng.permutation(p.shape[0])])  # node order is irrelevant
    ptr = np.concatenate([[0], np.cumsum([v.shape[0] for v in variants])])
    pos = torch.from_numpy(np.concatenate(variants)).cuda()
    nptr = torch.from_numpy(ptr).cuda()
    for tol in (1e-8, 1e-2):
        got = batcher.is_periodic(pos, nptr, tol=tol)
        want = [O.is_periodic(v, tol=tol) for v in variants]
        assert got.dtype == torch.bool and got.tolist() == want, (tol, got.tolist(), want)
    assert sum(O.is_periodic(v) for v in variants) not in (0, len(variants))  # both answers occur
    # one mesh, default node_ptr, [N,3] points like shape.points
    p3 = torch.from_numpy(np.hstack([meshes[0], np.zeros((meshes[0].shape[0], 1))])).cuda()
    assert batcher.is_periodic(p3).tolist() == [True]


@pytest.mark.parametrize("quads", [False, True])
def test_convert_mesh_to_graph_matches_the_oracle(quads):
    """batcher.convert_mesh_to_graph (benchmark_gnn_fem.py:388-415 on the GPU) vs the oracle's restatement: bit-exact
    edges / weights / labels, and the model accepts the un-batched result like the benchmark's bare forward call."""
    from pdivgnn_b200 import batcher, synth
    s = (synth.make_quad_rve_mesh if quads else synth.make_rve_mesh)(11, 900)
    ms = (1.5, -0.25, 3.0)
    ref = O.convert_mesh_to_graph(s["pos"], s["faces"], ms)
    pts3 = np.hstack([s["pos"][:, :2], np.zeros((s["pos"].shape[0], 1))])  # shape.points is [N,3]
    g = batcher.convert_mesh_to_graph(pts3, s["faces"], ms)
    assert torch.equal(g.edge_index.cpu(), ref.edge_index) and torch.equal(g.edge_attr.cpu(), ref.edge_attr)
    assert torch.equal(g.pos.cpu(), ref.pos) and torch.equal(g.mean_stress.cpu(), ref.mean_stress)
    assert torch.equal(g.nodes_types.cpu(), ref.nodes_types) and torch.equal(g.surfaces_nodes_for_div.cpu(), ref.surfaces_nodes_for_div)
    assert g.nodes_types.data_ptr() != g.surfaces_nodes_for_div.data_ptr() and g.is_periodic is True
    bad = pts3.copy()
    bad[np.where(bad[:, 0] == bad[:, 0].max())[0][3], 1] += 0.5
    with pytest.raises(AssertionError, match="not periodic"):
        batcher.convert_mesh_to_graph(bad, s["faces"], ms)
    samples, graphs, batch, stats = H.synthetic_batch(2, 256, seed0=69)
    model = H.make_model(stats)
    with torch.no_grad():
        out = model(g, scale_output=True, scale_input=True).local_stress
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    want = O.forward(sd, ref, stats, 10, True, True)
    linf, l2 = H.rel_err(out.cpu(), want)
    assert linf < 1e-5 and l2 < 1e-5, (linf, l2)


@pytest.mark.parametrize("quads", [False, True])
def test_ragged_rectangular_plates_bit_exact(quads):
    """Ragged batch of rectangular (nx != ny) periodic plates without a hole, 3 x 3 up to 40 x 17 nodes: device edges /
    weights / labels / periodicity flags vs the oracle, graph by graph (the plate generator of the CPU property tests)."""
    from pdivgnn_b200 import batcher
    from test_properties_cpu import _grid_mesh
    rng = np.random.default_rng(17)
    shapes = [(3, 3), (4, 9), (12, 5), (40, 17), (7, 7), (3, 25)]
    meshes = [_grid_mesh(nx, ny, rng, quads) for nx, ny in shapes]
    nptr = np.concatenate([[0], np.cumsum([p.shape[0] for p, _ in meshes])])
    fptr = np.concatenate([[0], np.cumsum([f.shape[1] for _, f in meshes])])
    pos = torch.from_numpy(np.concatenate([p[:, :2] for p, _ in meshes])).cuda()
    faces = torch.from_numpy(np.concatenate([f for _, f in meshes], axis=1)).cuda()
    ei, ew = batcher.build_edges(pos, faces, torch.from_numpy(nptr).cuda(), torch.from_numpy(fptr).cuda(), periodic=True)
    lab, reg = batcher.node_labels(pos, faces, torch.from_numpy(nptr).cuda(), torch.from_numpy(fptr).cuda())
    per = batcher.is_periodic(pos, torch.from_numpy(nptr).cuda())
    want_ei, want_ew, want_lab = [], [], []
    for (p, f), off in zip(meshes, nptr[:-1]):
        g = O.convert_mesh_to_graph(p, f, (1.0, 2.0, 3.0))
        want_ei.append(g.edge_index + int(off))
        want_ew.append(g.edge_attr)
        want_lab.append(g.nodes_types[:, 0])
    assert torch.equal(ei.cpu(), torch.cat(want_ei, dim=1)) and torch.equal(ew.cpu(), torch.cat(want_ew))
    assert torch.equal(lab.cpu(), torch.cat(want_lab)) and reg.tolist() == [1] * len(shapes)
    assert per.tolist() == [True] * len(shapes)
