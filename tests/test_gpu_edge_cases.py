"""Edge cases and full-size properties of the hot path (ragged inputs, degenerate graphs,
BASELINE.json configs[1] size).  Non-mesh graphs exercise the receiver-segment bookkeeping of the edge
kernels (tiles with 128 one-edge segments, receivers without edges, one segment spanning many tiles)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import pdg_helpers as H
from oracle import pdg_oracle as O

pytestmark = pytest.mark.gpu


def _random_graph(n, edges, seed):
    """A batch-like namespace for an arbitrary directed graph (no mesh): the fields EncodeProcessDecode reads."""
    g = torch.Generator().manual_seed(seed)
    ei = torch.as_tensor(edges, dtype=torch.int64)
    return SimpleNamespace(
        pos=torch.rand(n, 2, generator=g) * 100.0, edge_index=ei.contiguous(),
        edge_attr=torch.rand(ei.shape[1], generator=g) * 5.0, mean_stress=(torch.rand(1, 3, generator=g) * 2 - 1).expand(n, 3).contiguous() * 50.0,
        nodes_types=torch.randint(-1, 2, (n, 1), generator=g), local_stress=torch.randn(n, 3, generator=g) * 40.0,
        ptr=torch.tensor([0, n]), batch=torch.zeros(n, dtype=torch.int64), batch_size=1, num_nodes=n)


def _stats():
    t = torch.tensor
    return dict(mean_pos=t(50.), std_pos=t(29.), mean_mean_stress=t(0.), std_mean_stress=t(30.), mean_local_stress=t(0.),
                std_local_stress=t(45.), mean_edge_weight=t(2.5), std_edge_weight=t(1.4))


def _device(b):
    d = SimpleNamespace(**{k: (v.cuda() if torch.is_tensor(v) else v) for k, v in vars(b).items()})
    return d


GRAPHS = {
    # every receiver has exactly one incoming edge: a tile holds 128 one-row segments
    "ring_degree1": lambda: (600, [list(range(600)), [(i + 1) % 600 for i in range(600)]]),
    # one hub receives from everybody (a single segment spanning several tiles), the hub sends to node 1
    "star_hub": lambda: (500, [list(range(1, 500)) + [0], [0] * 499 + [1]]),
    # half of the nodes have no edge at all; the others form a dense random graph with duplicates of (u,v) removed
    "sparse_isolated": lambda: (400, np.unique(np.random.default_rng(5).integers(0, 200, size=(2, 3000)), axis=1).tolist()),
}


@pytest.mark.parametrize("name", list(GRAPHS))
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_non_mesh_graphs_forward_and_gradients(name, precision):
    import pdivgnn_b200
    n, edges = GRAPHS[name]()
    b = _random_graph(n, edges, 7)
    stats = _stats()
    sd = O.init_state_dict(seed=69)
    model = H.make_model(stats, params=sd)
    model.precision = precision
    db = _device(b)
    pred = model(db, scale_output=False).local_stress
    ref = O.forward(sd, b, stats, 10, scale_output=False, dtype=torch.float64)
    # bf16: the north star's 2e-2 is stated (and tested, test_gpu_bf16.py) for mesh graphs; on these degenerate
    # topologies the L2 error stays below it and the L-inf error may touch it (measured up to 2.1e-2)
    tol = 1e-5 if precision == "fp32" else 2e-2
    linf, l2 = H.rel_err(pred.detach().cpu(), ref)
    assert l2 < tol and linf < (tol if precision == "fp32" else 5e-2), (name, precision, linf, l2)
    nmse, div = pdivgnn_b200.nmse_div_loss(pred, db, model, False, 0.0)
    nmse.backward()
    r = O.loss_and_grads(sd, b, stats, 10, False, 0.0, dtype=torch.float64)
    assert abs(nmse.item() - float(r[0])) <= 10 * tol * abs(float(r[0]))
    cat = lambda d: torch.cat([d[k].double().flatten() for k in O.STATE_KEYS])  # noqa: E731
    ours = {k: p.grad.cpu() for k, p in model.named_parameters()}
    # fp32: the reference's own fp32-vs-fp64 gradient noise is ~1e-4.  bf16: the 2e-2 budget of the north star is
    # stated for mesh graphs; these degenerate topologies (most nodes without in-edges, a 499-edge hub) amplify the
    # operand rounding through ten LayerNorm-ed steps (3-9 % measured, independent of the hub size, fp32 kernels
    # exact on the same graphs), so the bf16 leg only guards the segment logic against O(1) errors.
    gtol = 2e-4 if precision == "fp32" else 0.15
    linf, l2 = H.rel_err(cat(ours), cat(r[4]))
    assert l2 < gtol, (name, precision, "grads", linf, l2)


def test_ragged_batch_of_very_different_meshes():
    """One 4 000-node mesh next to 120-node meshes (tiles that mix graphs, ptr segments of very different size)."""
    from pdivgnn_b200 import synth
    import pdivgnn_b200
    samples = [synth.make_rve_mesh(1, 120), synth.make_rve_mesh(2, 4000), synth.make_rve_mesh(3, 120), synth.make_rve_mesh(4, 130)]
    graphs, batch, stats = H.oracle_batch_from_samples(samples, True)
    sd = O.init_state_dict(seed=69)
    model = H.make_model(stats, params=sd)
    db = H.DeviceBatch(batch)
    pred = model(db, scale_output=False).local_stress
    nmse, div = pdivgnn_b200.nmse_div_loss(pred, db, model, True, 10.0)
    tot, o_nmse, o_div, o_pred = O.train_loss(sd, batch, stats, 10, True, 10.0, dtype=torch.float64)
    linf, l2 = H.rel_err(pred.detach().cpu(), o_pred)
    assert linf < 1e-5 and l2 < 1e-5
    assert abs(nmse.item() - float(o_nmse)) < 1e-5 * abs(float(o_nmse)) and abs(div.item() - float(o_div)) < 1e-5 * abs(float(o_div))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_full_size_batch_of_copies_equals_the_single_graph(precision):
    """BASELINE configs[1] size (32 meshes x ~1024 nodes).  Graph LayerNorm statistics are means over the batch, so
    a batch of 32 copies of one mesh has the statistics of the single mesh: every copy must reproduce the B = 1
    prediction (size-independent property; summation order differs, hence the small tolerance), and the
    gradient of the mean loss over the copies must equal the single-graph gradient."""
    import pdivgnn_b200
    from pdivgnn_b200 import synth
    s = synth.make_rve_mesh(69, 1024)
    g1, b1, stats = H.oracle_batch_from_samples([s], True)
    g32, b32, _ = H.oracle_batch_from_samples([s] * 32, True)
    sd = O.init_state_dict(seed=69)
    model = H.make_model(stats, params=sd)
    model.precision = precision
    d1, d32 = H.DeviceBatch(b1), H.DeviceBatch(b32)
    n = b1.num_nodes
    p1 = model(d1, scale_output=False).local_stress
    l1, _ = pdivgnn_b200.nmse_div_loss(p1, d1, model, False, 0.0)
    l1.backward()
    g_single = torch.cat([p.grad.flatten() for p in model.parameters()]).clone()
    model.zero_grad()
    p32 = model(d32, scale_output=False).local_stress
    l32, _ = pdivgnn_b200.nmse_div_loss(p32, d32, model, False, 0.0)
    l32.backward()
    g_batch = torch.cat([p.grad.flatten() for p in model.parameters()])
    tol = 2e-5 if precision == "fp32" else 2e-2
    copies = p32.detach().view(32, n, 3)
    for c in (0, 13, 31):
        linf, l2 = H.rel_err(copies[c].cpu(), p1.detach().cpu())
        assert linf < tol and l2 < tol, (precision, c, linf, l2)
    assert abs(l32.item() - l1.item()) < tol * abs(l1.item())
    linf, l2 = H.rel_err(g_batch.cpu(), g_single.cpu())
    assert l2 < (5e-4 if precision == "fp32" else 2e-2), (precision, "grad", linf, l2)
    # and the whole full-size step is bit-reproducible run to run
    model.zero_grad()
    q32 = model(d32, scale_output=False).local_stress
    m32, _ = pdivgnn_b200.nmse_div_loss(q32, d32, model, False, 0.0)
    m32.backward()
    assert torch.equal(q32, p32) and torch.equal(torch.cat([p.grad.flatten() for p in model.parameters()]), g_batch)


def test_invalid_inputs_raise():
    n, edges = GRAPHS["ring_degree1"]()
    b = _random_graph(n, edges, 1)
    model = H.make_model(_stats(), params=O.init_state_dict(seed=69))
    db = _device(b)
    db.edge_attr = db.edge_attr[:-1]
    with pytest.raises(ValueError):
        model(db)
    db = _device(b)
    db.mean_stress = db.mean_stress.double()
    with pytest.raises(TypeError):
        model(db)
    db = _device(b)
    db.edge_index = db.edge_index[:, :0]
    db.edge_attr = db.edge_attr[:0]
    with pytest.raises(RuntimeError, match="E=0|empty graph"):
        model(db)
