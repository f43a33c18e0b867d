"""The oracle (oracle/pdg_oracle.py) replayed against outputs of the reference's own,
unmodified code (tests/golden/*.npz, written by oracle/make_golden.py in the build
container).  CPU only.  Graph construction, collation, stats, default init, forward
fields and losses are bit-exact (same torch ops, same order); gradients agree to 2e-6."""
import numpy as np
import pytest
import torch

from oracle import pdg_oracle as O
import pdg_helpers as H

CASES = ["train2_div", "train2_nodiv", "train3_noperiodic", "infer1", "train2_quad"]


def test_grid3x3_known_answer():
    """32 mesh edges + 16 periodic = 48; node 0 neighbours/weights; in-degree."""
    g = H.load_golden("grid3x3")
    pos = torch.from_numpy(g["pos"])
    face = torch.from_numpy(g["faces"])
    ei = O.face_to_edge(face, 9)
    assert np.array_equal(ei.numpy(), g["mesh_edge_index"]) and ei.shape[1] == 32
    ea = O.edge_weights(pos, ei).float()
    assert np.array_equal(ea.numpy(), g["mesh_edge_attr"])
    pei, pea = O.compute_periodic_graph(pos, ei, ea)
    assert pei.shape[1] == 48
    assert np.array_equal(pei.numpy(), g["edge_index"])
    assert np.array_equal(pea.numpy(), g["edge_attr"])
    out0 = pei[1][pei[0] == 0].tolist()
    assert out0 == [1, 2, 3, 4, 6, 8]
    np.testing.assert_allclose(pea[pei[0] == 0].numpy(), [1, 0, 1, 2 ** 0.5, 0, 0], rtol=1e-6)
    assert torch.bincount(pei[1], minlength=9).tolist() == [6, 5, 5, 5, 6, 5, 5, 5, 6]
    # symmetric, sorted, unique
    key = pei[0] * 9 + pei[1]
    assert torch.all(key[1:] > key[:-1])
    assert set(map(tuple, pei.t().tolist())) == set(map(tuple, pei.flip(0).t().tolist()))


def test_quad_grid_known_answer():
    """convert_utils.py:52-81 (``_quad_face_to_edge``) run by the reference itself on 2 x 3 quads (pitch 1.5 x 1):
    17 undirected sides = 34 directed edges, NO diagonals (a triangulation of the same grid would add 6 x 2);
    periodic pairing adds the left/right (4), lower/upper (3) pairs both ways minus nothing + the 4 corner links."""
    g = H.load_golden("quad_grid")
    pos, face = torch.from_numpy(g["pos"]), torch.from_numpy(g["faces"])
    assert face.shape == (4, 6)
    ei = O.quad_face_to_edge(face, 12)
    assert np.array_equal(ei.numpy(), g["mesh_edge_index"]) and ei.shape[1] == 34
    ea = O.edge_weights(pos, ei).float()
    assert np.array_equal(ea.numpy(), g["mesh_edge_attr"])
    assert set(np.unique(ea.numpy()).tolist()) == {1.0, 1.5}  # sides only: a diagonal would be sqrt(3.25)
    pei, pea = O.compute_periodic_graph(pos, ei, ea)
    assert np.array_equal(pei.numpy(), g["edge_index"]) and np.array_equal(pea.numpy(), g["edge_attr"])
    # node 0 (corner): mesh neighbours 1 (dx 1.5) and 3 (dy 1); periodic partners 2 (right), 9 (upper), 11 (corner)
    assert pei[1][pei[0] == 0].tolist() == [1, 2, 3, 9, 11]
    np.testing.assert_allclose(pea[pei[0] == 0].numpy(), [1.5, 0, 1, 0, 0])
    # build_graph dispatches on the face arity like mesh_to_graph (convert_utils.py:52-58)
    s = dict(pos=g["pos"], faces=g["faces"], stress_field=np.zeros((12, 3)), mean_stress=np.zeros(3), labels=np.zeros(12, np.int64),
             op_div_row=[0], op_div_col=[0], op_div_data=[0.0], op_div_shape=(12, 24))
    assert np.array_equal(O.build_graph(s, True).edge_index.numpy(), g["edge_index"])


@pytest.mark.parametrize("name", CASES)
def test_graph_construction_and_collation_bit_exact(name):
    g = H.load_golden(name)
    graphs, batch, stats = H.oracle_batch_from_samples(H.golden_samples(g), bool(g["periodic"]))
    assert np.array_equal(batch.edge_index.numpy(), g["edge_index"])
    assert np.array_equal(batch.edge_attr.numpy(), g["edge_attr"])
    assert np.array_equal(batch.ptr.numpy(), g["ptr"])
    assert np.array_equal(batch.op_div_matrix.indices().numpy(), g["op_indices"])
    assert np.array_equal(batch.op_div_matrix.values().numpy(), g["op_values"])
    assert tuple(batch.op_div_matrix.shape) == tuple(g["op_shape"])
    for k in H.STAT_KEYS:
        assert np.array_equal(stats[k].numpy(), g["stat_" + k]), k


def test_default_init_matches_reference():
    sd = O.init_state_dict(seed=69)
    ref = H.golden_params()
    assert list(sd.keys()) == O.STATE_KEYS and len(sd) == 28
    assert sum(v.numel() for v in sd.values()) == 167299
    for k in O.STATE_KEYS:
        assert torch.equal(sd[k], ref[k]), k


@pytest.mark.parametrize("name", CASES)
def test_forward_loss_grads_bit_exact(name):
    g = H.load_golden(name)
    graphs, batch, stats = H.oracle_batch_from_samples(H.golden_samples(g), bool(g["periodic"]))
    sd = H.golden_params()
    if "pred_scaled" in g.files:
        out = O.forward(sd, batch, stats, 10, scale_output=True, scale_input=True)
        assert np.array_equal(out.numpy(), g["pred_scaled"])
    total, nmse, div, pred, grads = O.loss_and_grads(sd, batch, stats, 10, bool(g["divergence"]), float(g["penalty"]))
    assert np.array_equal(pred.numpy(), g["pred_std"])
    assert np.array_equal(total.numpy(), g["loss"])
    assert np.array_equal(nmse.numpy(), g["nmse"])
    assert np.array_equal(np.asarray(div, dtype=np.float32), g["div"])
    if "grad_" + O.STATE_KEYS[0] in g.files:
        # torch's CPU backward is not bit-reproducible run to run (threaded reductions) and the
        # fan-in of x / e is summed in graph-topology order: last-ulp differences only
        for k in O.STATE_KEYS:
            linf, l2 = H.rel_err(grads[k], g["grad_" + k])
            assert linf < 2e-6 and l2 < 2e-6, (k, linf, l2)


def test_divergence_densify_equals_spmm():
    """compare_results.py:647-673 (plain op_div @ S) == gnn_train.py:73-76 (densify+slice)."""
    g = H.load_golden("train2_div")
    graphs, batch, stats = H.oracle_batch_from_samples(H.golden_samples(g), True)
    pred = torch.from_numpy(g["pred_std"])
    ptr = batch.ptr.tolist()
    for i, gr in enumerate(graphs):
        s, e = ptr[i], ptr[i + 1]
        blk = O.op_div_row_block(batch.op_div_matrix, s, e)
        a = O.compute_divergence(pred[s:e], blk, batch.surfaces_nodes_for_div[s:e])
        m = gr.op_div_matrix
        b = O.compute_divergence_spmm(pred[s:e].double(), m.indices()[0], m.indices()[1], m.values(),
                                      gr.surfaces_nodes_for_div)
        assert abs(a.item() - b.item()) <= 2e-6 * abs(b.item())


def test_early_exit_on_zero_mean_stress():
    g = H.load_golden("infer1")
    graphs, batch, stats = H.oracle_batch_from_samples(H.golden_samples(g), True)
    batch.mean_stress = torch.zeros_like(batch.mean_stress)
    out = O.forward(H.golden_params(), batch, stats)
    assert out.shape == batch.mean_stress.shape and not out.any()


def test_known_answers_small():
    # LayerNorm graph mode on a 2x128 tensor: one mean / one population std
    torch.manual_seed(0)
    x = torch.randn(2, 128)
    w, b = torch.rand(128), torch.rand(128)
    y = O.graph_layer_norm(x, w, b)
    ref = (x - x.mean()) / (x.flatten().var(unbiased=False).sqrt() + 1e-5) * w + b
    assert torch.allclose(y, ref, atol=1e-6)
    # one-triangle divergence: sigma_xx = x  => d/dx = 1 at every node
    from pdivgnn_b200 import synth
    pos = np.array([[0.0, 0.0], [2.0, 0.0], [0.0, 2.0]])
    r, c, v = synth._p1_divergence_operator(pos, np.array([[0, 1, 2]]))
    s = torch.tensor([[0.0, 0.0, 0.0], [2.0, 0.0, 0.0], [0.0, 0.0, 0.0]])
    d = O.compute_divergence_spmm(s.double(), torch.from_numpy(r), torch.from_numpy(c), torch.from_numpy(v),
                                  torch.zeros(3, 1, dtype=torch.long))
    assert abs(d.item() - 1.0) < 1e-12  # mean over nodes of (1^2 + 0^2)


def test_fp64_noise_floor():
    """fp32 oracle vs fp64 oracle: the budget the 1e-5 CUDA tolerance has to live in."""
    g = H.load_golden("infer1")
    graphs, batch, stats = H.oracle_batch_from_samples(H.golden_samples(g), True)
    sd = H.golden_params()
    a = O.forward(sd, batch, stats, 10, scale_output=False)
    b = O.forward(sd, batch, stats, 10, scale_output=False, dtype=torch.float64)
    linf, l2 = H.rel_err(a, b)
    assert linf < 5e-6 and l2 < 5e-6


def test_node_labels_restatement_matches_the_generator_truth():
    """compute_node_labels (datasets.py:133-179) restated without VTK: the synthetic generator knows by
    construction which nodes are on the plate sides (1) and on the hole (-1); grid3x3 has a single loop."""
    from pdivgnn_b200 import synth
    for seed in (3, 69, 70):
        s = synth.make_rve_mesh(seed, 300)
        labels, nreg = O.compute_node_labels(s["pos"], s["faces"])
        assert nreg == 2 and np.array_equal(labels, np.asarray(s["labels"]))
    g = H.load_golden("grid3x3")
    labels, nreg = O.compute_node_labels(g["pos"], g["faces"])
    assert nreg == 1 and labels.tolist() == [1, 1, 1, 1, 0, 1, 1, 1, 1]
    for seed in (5, 69, 72):  # quad cells: a side used by one quad is a boundary edge
        s = synth.make_quad_rve_mesh(seed, 300)
        labels, nreg = O.compute_node_labels(s["pos"], s["faces"])
        assert nreg == 2 and np.array_equal(labels, np.asarray(s["labels"]))
    g = H.load_golden("quad_grid")
    labels, nreg = O.compute_node_labels(g["pos"], g["faces"])
    assert nreg == 1 and labels.tolist() == [1, 1, 1, 1, 0, 1, 1, 0, 1, 1, 1, 1]


def test_is_periodic_restatement_hand_cases():
    """``oracle.is_periodic`` (restated microgen.mesh.is_periodic, parity unpinned: microgen is absent) on hand-built
    node sets whose answer follows from the published definition: equal side counts and sorted partners within tol
    (one-sided: only an EXCESS of the max side over the min side counts)."""
    sq = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0], [0.0, 0.4], [1.0, 0.4], [0.3, 0.0], [0.3, 1.0], [0.5, 0.5]])
    assert O.is_periodic(sq)
    assert O.is_periodic(np.hstack([sq, np.zeros((9, 1))])[:, :-1])  # the reference's call shape: points[:, :-1]
    assert O.is_periodic(sq[::-1].copy())  # node order is irrelevant (sides are sorted)
    moved = sq.copy(); moved[5, 1] = 0.45  # right partner of (0, 0.4) sits higher by 0.05 > tol
    assert not O.is_periodic(moved)
    lower = sq.copy(); lower[5, 1] = 0.35  # right partner LOWER than the left one: the published one-sided test passes
    assert O.is_periodic(lower)
    assert O.is_periodic(moved, tol=0.1)  # inside a looser tolerance
    missing = np.delete(sq, 5, axis=0)  # a left node without a right partner: side counts differ
    assert not O.is_periodic(missing)
    top = sq.copy(); top[7, 0] = 0.31  # top partner of (0.3, 0) shifted along x
    assert not O.is_periodic(top)
    assert O.is_periodic(top, dim=1)  # dim = 1 tests the x sides only
    near = sq.copy(); near[4, 0] = 5e-9  # within tol of the bounding box: still a side node
    assert O.is_periodic(near)
    off = sq.copy(); off[4, 0] = 5e-8  # outside tol: not a side node any more -> counts differ
    assert not O.is_periodic(off)
    from pdivgnn_b200 import synth
    for s in synth.make_dataset(3, 300, 7) + synth.make_dataset(2, 300, 7, quads=True):
        assert O.is_periodic(s["pos"][:, :2])  # the synthetic RVE generator is periodic by construction
