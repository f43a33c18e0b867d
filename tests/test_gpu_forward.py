"""GPU parity of the CUDA forward path (through the C ABI) against the CPU oracle.

Tolerances (BASELINE.json north_star): bit-exact for index work (plan / CSR), 1e-5
norm-wise relative (L-inf and L2 per tensor) for fp32 fields."""
import numpy as np
import pytest
import torch

import pdg_helpers as H
from oracle import pdg_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5
CASES = ["train2_div", "train2_nodiv", "train3_noperiodic", "infer1", "train2_quad"]


@pytest.fixture(scope="module")
def cuda():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _golden(name):
    g = H.load_golden(name)
    graphs, batch, stats = H.oracle_batch_from_samples(H.golden_samples(g), bool(g["periodic"]))
    return g, batch, stats


def test_plan_bit_exact(cuda):
    from pdivgnn_b200.autograd import build_plan
    for name in CASES:
        g, batch, stats = _golden(name)
        ei = batch.edge_index
        n, e = batch.num_nodes, ei.shape[1]
        plan = build_plan(ei.to(cuda), n)
        perm, recv, send, rowptr, sptr, slist = [v.cpu().long() for v in plan.views()]
        order = torch.sort(ei[1], stable=True).indices  # receiver-sorted, ties in input order
        assert torch.equal(perm[:e], order)
        assert torch.equal(recv[:e], ei[1][order]) and torch.equal(send[:e], ei[0][order])
        assert not perm[e:].any() and not recv[e:].any() and not send[e:].any()
        deg = torch.bincount(ei[1], minlength=n)
        assert torch.equal(rowptr, torch.cat([torch.zeros(1, dtype=torch.long), deg.cumsum(0)]))
        s_sorted = torch.sort(send[:e], stable=True)
        assert torch.equal(slist, s_sorted.indices)
        outdeg = torch.bincount(ei[0], minlength=n)
        assert torch.equal(sptr, torch.cat([torch.zeros(1, dtype=torch.long), outdeg.cumsum(0)]))
        assert int(rowptr[-1]) == e  # sum of in-degrees == E


@pytest.mark.parametrize("name", CASES)
def test_forward_matches_golden(cuda, name):
    g, batch, stats = _golden(name)
    model = H.make_model(stats, params=H.golden_params())
    db = H.DeviceBatch(batch)
    with torch.no_grad():
        out = model(db, scale_output=False, scale_input=True).local_stress.cpu()
    linf, l2 = H.rel_err(out, g["pred_std"])
    assert linf < TOL and l2 < TOL, (linf, l2)
    if "pred_scaled" in g.files:
        with torch.no_grad():
            out = model.forward(db, scale_output=True, scale_input=True).local_stress.cpu()
        linf, l2 = H.rel_err(out, g["pred_scaled"])
        assert linf < TOL and l2 < TOL, (linf, l2)


def test_forward_stages_vs_oracle(cuda):
    """Every saved intermediate (x_t, e_t in receiver-sorted order) against the fp64 oracle."""
    from pdg_debug import forward_with_state
    g, batch, stats = _golden("train2_div")
    sd = H.golden_params()
    model = H.make_model(stats, params=sd)
    st = forward_with_state(model, H.DeviceBatch(batch))
    torch.cuda.synchronize()
    out64, lat = O.forward(sd, batch, stats, 10, scale_output=False, dtype=torch.float64, return_latents=True)
    n, e = batch.num_nodes, batch.edge_index.shape[1]
    perm = st.plan.views()[0].cpu().long()[:e]
    worst = 0.0
    for t in range(10):
        x_o, e_o = lat[t]
        linf, l2 = H.rel_err(st.tensor("x", t)[:n].cpu(), x_o)
        assert linf < TOL, ("x", t, linf)
        worst = max(worst, linf)
        linf, l2 = H.rel_err(st.tensor("e", t)[:e].cpu(), e_o[perm])
        assert linf < TOL, ("e", t, linf)
        worst = max(worst, linf)
    linf, l2 = H.rel_err(st.tensor("x", 10)[:n].cpu(), lat[10][0])
    assert linf < TOL
    linf, l2 = H.rel_err(st.out.cpu(), out64)
    assert linf < TOL and l2 < TOL
    print("worst stage rel err", worst)


def test_forward_synthetic_batch_and_determinism(cuda):
    samples, graphs, batch, stats = H.synthetic_batch(4, 1024)
    sd = O.init_state_dict(seed=69)
    model = H.make_model(stats, params=sd)
    db = H.DeviceBatch(batch)
    with torch.no_grad():
        a = model(db, scale_output=True).local_stress
        b = model(db, scale_output=True).local_stress
    assert torch.equal(a, b), "forward must be bit-reproducible run to run"
    ref = O.forward(sd, batch, stats, 10, scale_output=True, dtype=torch.float64)
    linf, l2 = H.rel_err(a.cpu(), ref)
    assert linf < TOL and l2 < TOL, (linf, l2)


def test_unbatched_graph_and_early_exit(cuda):
    """benchmark_gnn_fem.py:97 calls model.forward on a bare Data (no batch/ptr)."""
    samples, graphs, batch, stats = H.synthetic_batch(1, 500, seed0=5)
    sd = O.init_state_dict(seed=69)
    model = H.make_model(stats, params=sd)

    class Bare:
        pass

    d = Bare()
    for k in ("pos", "edge_index", "edge_attr", "mean_stress", "nodes_types"):
        setattr(d, k, getattr(batch, k).cuda())
    with torch.no_grad():
        out = model.forward(d).local_stress.cpu()
    ref = O.forward(sd, batch, stats, 10, dtype=torch.float64)
    linf, l2 = H.rel_err(out, ref)
    assert linf < TOL and l2 < TOL
    d.mean_stress = torch.zeros_like(d.mean_stress)
    z = model.forward(d)
    assert z.local_stress.shape == d.mean_stress.shape and not z.local_stress.any()
    assert z.edge_index is d.edge_index and z.pos is d.pos


def test_edge_order_permutation_invariance(cuda):
    """Outputs do not depend on the order edges are listed in (property)."""
    samples, graphs, batch, stats = H.synthetic_batch(2, 300, seed0=11)
    sd = O.init_state_dict(seed=69)
    model = H.make_model(stats, params=sd)
    db = H.DeviceBatch(batch)
    with torch.no_grad():
        a = model(db).local_stress
        p = torch.randperm(db.edge_index.shape[1], device="cuda", generator=torch.Generator("cuda").manual_seed(1))
        db.edge_index = db.edge_index[:, p].contiguous()
        db.edge_attr = db.edge_attr[p].contiguous()
        b = model(db).local_stress
    linf, l2 = H.rel_err(b.cpu(), a.cpu())
    assert linf < TOL and l2 < TOL


@pytest.mark.parametrize("name", ["train2_div", "train2_nodiv", "train3_noperiodic"])
def test_loss_matches_golden(cuda, name):
    import pdivgnn_b200
    g, batch, stats = _golden(name)
    model = H.make_model(stats, params=H.golden_params())
    db = H.DeviceBatch(batch)
    pred = torch.from_numpy(g["pred_std"]).cuda().requires_grad_(True)
    nmse, div = pdivgnn_b200.nmse_div_loss(pred, db, model, bool(g["divergence"]), float(g["penalty"]))
    assert abs(nmse.item() - float(g["nmse"])) <= 1e-5 * abs(float(g["nmse"]))
    assert abs(div.item() - float(g["div"])) <= 1e-5 * max(abs(float(g["div"])), 1e-12)
    (nmse + div).backward()
    # d loss / d pred from the oracle (fp64)
    p64 = torch.from_numpy(g["pred_std"]).double().requires_grad_(True)
    gt = (batch.local_stress.double() - stats["mean_local_stress"].double()) / stats["std_local_stress"].double()
    tot = 0
    ptr = batch.ptr.tolist()
    for i in range(batch.batch_size):
        s, e = ptr[i], ptr[i + 1]
        tot = tot + O.normalized_mse_loss_single(gt[s:e], p64[s:e]) / batch.batch_size
        if bool(g["divergence"]):
            blk = O.op_div_row_block(batch.op_div_matrix, s, e).double()
            tot = tot + O.compute_divergence(p64[s:e], blk, batch.surfaces_nodes_for_div[s:e]) * float(g["penalty"]) / batch.batch_size
    tot.backward()
    linf, l2 = H.rel_err(pred.grad.cpu(), p64.grad)
    assert linf < TOL and l2 < TOL, (linf, l2)


@pytest.mark.parametrize("steps", [1, 3])
def test_other_step_counts(cuda, steps):
    samples, graphs, batch, stats = H.synthetic_batch(2, 200, seed0=21)
    sd = O.init_state_dict(seed=3)
    model = H.make_model(stats, steps=steps, params=sd)
    with torch.no_grad():
        out = model(H.DeviceBatch(batch), scale_output=False).local_stress.cpu()
    ref = O.forward(sd, batch, stats, steps, scale_output=False, dtype=torch.float64)
    linf, l2 = H.rel_err(out, ref)
    assert linf < TOL and l2 < TOL, (steps, linf, l2)


def test_inference_keeps_no_training_state(cuda):
    """no_grad forward must not allocate the per-step backward state (ping-pong buffers only)."""
    samples, graphs, batch, stats = H.synthetic_batch(8, 1024, seed0=40)
    model = H.make_model(stats, params=O.init_state_dict(seed=69))
    db = H.DeviceBatch(batch)
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    with torch.no_grad():
        model(db)
    torch.cuda.synchronize()
    infer = torch.cuda.max_memory_allocated() - base
    torch.cuda.reset_peak_memory_stats()
    out = model(db).local_stress
    torch.cuda.synchronize()
    train = torch.cuda.max_memory_allocated() - base
    del out
    e_bytes = (batch.edge_index.shape[1] + 127) // 128 * 128 * 512
    assert infer < 4 * e_bytes, (infer, e_bytes)       # ~2 edge-sized buffers + node buffers
    assert train > 10 * e_bytes, (train, e_bytes)      # 10 steps x (e_t, y2_t)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_cuda_graph_replay_of_inference_forward(cuda, precision):
    """`cuda_graphs=True`: a no-grad forward on the same input tensors is captured once and replayed as ONE launch
    (benchmark_gnn_fem.py:81-100 calls the model on the same Data six times).  The replay must be bit-identical to the
    eager launch sequence, must see in-place edits of inputs and parameters (same addresses), and must re-capture for
    other tensors; a forward that needs gradients never takes the graph path."""
    g, batch, stats = _golden("infer1")
    sd = H.golden_params()
    eager = H.make_model(stats, params=sd)
    eager.precision = precision
    graphed = H.make_model(stats, params=sd)
    graphed.precision = precision
    graphed.cuda_graphs = True
    db = H.DeviceBatch(batch)
    with torch.no_grad():
        ref = eager(db).local_stress
        for _ in range(3):  # capture, then two replays
            assert torch.equal(graphed(db).local_stress, ref)
        db.mean_stress.mul_(0.5)  # in-place input edit: same tensor, new values
        ref2 = eager(db).local_stress
        assert not torch.equal(ref2, ref) and torch.equal(graphed(db).local_stress, ref2)
        for pe, pg in zip(eager.parameters(), graphed.parameters()):  # in-place parameter update (what an optimizer does)
            pe.mul_(1.01)
            pg.mul_(1.01)
        ref3 = eager(db).local_stress
        assert not torch.equal(ref3, ref2) and torch.equal(graphed(db).local_stress, ref3)
        db2 = H.DeviceBatch(batch)  # other tensors -> another capture
        assert torch.equal(graphed(db2).local_stress, eager(db2).local_stress)
    out = graphed(db2).local_stress  # grad mode: the autograd path, not the graph
    assert out.requires_grad
    out.sum().backward()
    assert all(p.grad is not None for p in graphed.parameters())
