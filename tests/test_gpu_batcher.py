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
