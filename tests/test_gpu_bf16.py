"""bf16 tensor-core tile mode (PDG_PREC_BF16: tcgen05/TMEM MLP tiles, fp32 latents / LayerNorm / reductions).

Tolerance from BASELINE.json north_star: 2e-2 norm-wise relative (L-inf AND L2) on fields, loss and
gradients.  Gradients are checked PER TENSOR (all 28) as well as concatenated, against the fp64 oracle, on the three
training goldens (divergence on / off, periodic on / off), at the full configs[1] size (32 distinct meshes, the seeds
of bench.py) and along a 50-step training trajectory against the fp32 mode.
"""
import pytest
import torch

import pdg_helpers as H
from oracle import pdg_oracle as O

pytestmark = pytest.mark.gpu
TOL = 2e-2


def _model(stats, sd, precision="bf16"):
    m = H.make_model(stats, params=sd)
    m.precision = precision
    return m


def _cat(d):
    return torch.cat([d[k].double().flatten() for k in O.STATE_KEYS])


def _train_case(batch, stats, sd, divergence, penalty, precision):
    import pdivgnn_b200
    model = _model(stats, sd, precision)
    db = H.DeviceBatch(batch)
    pred = model(db, scale_output=False).local_stress
    nmse, div = pdivgnn_b200.nmse_div_loss(pred, db, model, divergence, penalty)
    (nmse + div).backward()
    grads = {k: p.grad.detach().cpu() for k, p in model.named_parameters()}
    return (nmse + div).item(), pred.detach().cpu(), grads


def _check_per_tensor(grads, ref, tol, what):
    """every one of the 28 gradient tensors and their concatenation: L-inf and L2 <= tol (norm-wise relative)"""
    rep = [(k, *H.rel_err(grads[k], ref[k])) for k in O.STATE_KEYS]
    rep.append(("flat", *H.rel_err(_cat(grads), _cat(ref))))
    worst = max(rep, key=lambda r: max(r[1], r[2]))
    print(f"{what}: worst tensor {worst[0]} Linf {worst[1]:.2e} L2 {worst[2]:.2e}; flat Linf {rep[-1][1]:.2e} L2 {rep[-1][2]:.2e}")
    bad = [(k, f"{a:.2e}", f"{b:.2e}") for k, a, b in rep if a > tol or b > tol]
    assert not bad, (what, bad)
    return rep


@pytest.mark.parametrize("name", ["train2_div", "infer1", "train3_noperiodic"])
def test_bf16_forward_golden(name):
    g = H.load_golden(name)
    graphs, batch, stats = H.oracle_batch_from_samples(H.golden_samples(g), bool(g["periodic"]))
    model = _model(stats, H.golden_params())
    with torch.no_grad():
        out = model(H.DeviceBatch(batch), scale_output=False).local_stress.cpu()
    linf, l2 = H.rel_err(out, g["pred_std"])
    print(f"{name}: bf16 forward rel err Linf {linf:.2e} L2 {l2:.2e}")
    assert linf < TOL and l2 < TOL, (linf, l2)


def test_bf16_forward_batch_deterministic_and_close_to_fp32():
    samples, graphs, batch, stats = H.synthetic_batch(4, 1024)
    sd = O.init_state_dict(seed=69)
    db = H.DeviceBatch(batch)
    m16, m32 = _model(stats, sd), H.make_model(stats, params=sd)
    with torch.no_grad():
        a = m16(db).local_stress
        b = m16(db).local_stress
        c = m32(db).local_stress
    assert torch.equal(a, b)
    linf, l2 = H.rel_err(a.cpu(), c.cpu())
    print(f"bf16 vs fp32 path: Linf {linf:.2e} L2 {l2:.2e}")
    assert linf < TOL and l2 < TOL


@pytest.mark.parametrize("name", ["train2_div", "train2_nodiv", "train3_noperiodic", "train2_quad"])
def test_bf16_training_step_loss_and_per_tensor_grads(name):
    """gnn_train.py:154-207 on the reference-generated goldens: loss and ALL 28 gradient tensors within 2e-2 of the
    fp64 oracle (and of the reference's own fp32 gradients stored in the golden file)."""
    g = H.load_golden(name)
    graphs, batch, stats = H.oracle_batch_from_samples(H.golden_samples(g), bool(g["periodic"]))
    sd = H.golden_params()
    div, pen = bool(g["divergence"]), float(g["penalty"])
    loss, pred, grads = _train_case(batch, stats, sd, div, pen, "bf16")
    assert abs(loss - float(g["loss"])) <= TOL * abs(float(g["loss"])), (loss, float(g["loss"]))
    linf, l2 = H.rel_err(pred, g["pred_std"])
    assert linf < TOL and l2 < TOL, ("fields", linf, l2)
    r64 = O.loss_and_grads(sd, batch, stats, 10, div, pen, dtype=torch.float64)
    _check_per_tensor(grads, r64[4], TOL, f"{name} bf16 grads vs fp64 oracle")
    ref32 = {k: torch.from_numpy(g["grad_" + k]) for k in O.STATE_KEYS}
    _check_per_tensor(grads, ref32, TOL, f"{name} bf16 grads vs reference fp32")


# ---- configs[1] size: 32 DISTINCT meshes (bench.py's seeds), CUDA path vs the CPU oracle -------------------------------
@pytest.fixture(scope="module")
def full_size_case():
    samples, graphs, batch, stats = H.synthetic_batch(32, 1024, seed0=69)
    sd = O.init_state_dict(seed=69)
    r64 = O.loss_and_grads(sd, batch, stats, 10, False, 10.0, dtype=torch.float64)
    r32 = O.loss_and_grads(sd, batch, stats, 10, False, 10.0, dtype=torch.float32)
    return batch, stats, sd, r64, r32


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_full_size_batch_vs_cpu_oracle(full_size_case, precision):
    """BASELINE configs[1] (32 meshes x ~1 024 nodes, divergence off): fields, loss, per-tensor and flat gradients of
    the CUDA path against the CPU oracle (fp64 yard-stick; fp32 oracle = the reference's own arithmetic)."""
    batch, stats, sd, r64, r32 = full_size_case
    loss, pred, grads = _train_case(batch, stats, sd, False, 10.0, precision)
    tol = TOL if precision == "bf16" else 1e-5
    linf, l2 = H.rel_err(pred, r64[3])
    print(f"configs[1] {precision}: N={batch.num_nodes} E={batch.edge_index.shape[1]} fields Linf {linf:.2e} L2 {l2:.2e}; "
          f"loss {loss:.6f} vs {float(r64[0]):.6f}")
    assert linf < tol and l2 < tol, ("fields", linf, l2)
    assert abs(loss - float(r64[0])) <= tol * abs(float(r64[0]))
    if precision == "bf16":
        _check_per_tensor(grads, r64[4], TOL, "configs[1] bf16 grads vs fp64 oracle")
    else:
        # fp32 mode: same bar as tests/test_gpu_backward.py (the reference's own fp32 gradients sit 2e-5..1e-4 from fp64)
        for k in O.STATE_KEYS:
            ours, ref = H.rel_err(grads[k], r64[4][k]), H.rel_err(r32[4][k], r64[4][k])
            assert ours[0] <= max(1e-5, 4 * ref[0]) and ours[1] <= max(1e-5, 3 * ref[1]), (k, ours, ref)
        ours, ref = H.rel_err(_cat(grads), _cat(r64[4])), H.rel_err(_cat(r32[4]), _cat(r64[4]))
        print(f"configs[1] fp32 grads: flat L2 {ours[1]:.2e} (oracle fp32 {ref[1]:.2e})")
        assert ours[1] <= max(1e-5, 2 * ref[1]), ("flat", ours, ref)


def test_bf16_training_trajectory_tracks_fp32_mode():
    """bf16 storage of the saved activations feeds the NEXT step's weights, so rounding could compound over optimizer
    steps: 50 Adam steps (gnn_train.py:154-207, lr 1e-3) from the same seed in both modes must end at the same loss
    (2e-2) and stay close along the way."""
    import pdivgnn_b200
    from pdivgnn_b200.optim import FusedAdam
    samples, graphs, batch, stats = H.synthetic_batch(4, 512, seed0=169)
    db = H.DeviceBatch(batch)
    traj = {}
    for prec in ("fp32", "bf16"):
        model = _model(stats, O.init_state_dict(seed=69), prec)
        opt = FusedAdam(model.parameters(), lr=1e-3)
        losses = []
        for _ in range(50):
            pred = model(db, scale_output=False).local_stress
            nmse, div = pdivgnn_b200.nmse_div_loss(pred, db, model, True, 10.0)
            loss = nmse + div
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            losses.append(loss.detach())
        traj[prec] = torch.stack(losses).cpu().double()
    a, b = traj["bf16"], traj["fp32"]
    rel = ((a - b).abs() / b.abs())
    print(f"trajectory: loss {b[0]:.4f} -> fp32 {b[-1]:.5f} / bf16 {a[-1]:.5f}; final rel diff {rel[-1]:.2e}, "
          f"max over 50 steps {rel.max():.2e}")
    assert b[-1] < 0.5 * b[0], "the fp32 run must actually train"
    assert rel[-1] < TOL, (float(a[-1]), float(b[-1]))
    assert rel.max() < 2.5 * TOL, float(rel.max())
