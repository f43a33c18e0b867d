"""bf16 tensor-core tile mode (PDG_PREC_BF16: tcgen05/TMEM edge-MLP tiles, fp32 everywhere else).
Tolerance from BASELINE.json north_star: 2e-2 norm-wise relative on fields, loss, gradients."""
import pytest
import torch

import pdg_helpers as H
from oracle import pdg_oracle as O

pytestmark = pytest.mark.gpu
TOL = 2e-2


def _model(stats, sd):
    m = H.make_model(stats, params=sd)
    m.precision = "bf16"
    return m


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


def test_bf16_training_step_loss_and_grads():
    import pdivgnn_b200
    g = H.load_golden("train2_div")
    graphs, batch, stats = H.oracle_batch_from_samples(H.golden_samples(g), True)
    sd = H.golden_params()
    model = _model(stats, sd)
    db = H.DeviceBatch(batch)
    pred = model(db, scale_output=False).local_stress
    nmse, div = pdivgnn_b200.nmse_div_loss(pred, db, model, True, 10.0)
    (nmse + div).backward()
    assert abs((nmse + div).item() - float(g["loss"])) <= TOL * abs(float(g["loss"]))
    cat = lambda d: torch.cat([d[k].double().flatten() for k in O.STATE_KEYS])  # noqa: E731
    ours = {k: p.grad.cpu() for k, p in model.named_parameters()}
    ref = {k: torch.from_numpy(g["grad_" + k]) for k in O.STATE_KEYS}
    linf, l2 = H.rel_err(cat(ours), cat(ref))
    print(f"bf16 grads (flat): Linf {linf:.2e} L2 {l2:.2e}")
    assert l2 < TOL and linf < 2 * TOL
