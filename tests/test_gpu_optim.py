"""Fused Adam (pdg_adam_step) vs torch.optim.Adam on the 28 parameter tensors (SURVEY 8f rank 3;
reference: scripts/gnn_train.py:111,118,204-207), and the device-predicated zero-load early exit
in the backward (models.py:294-299)."""
import copy

import pytest
import torch

import pdg_helpers as H
from oracle import pdg_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _two_models():
    import pdivgnn_b200
    torch.manual_seed(69)
    a = pdivgnn_b200.EncodeProcessDecode(1, 10, 128, 6, 3).cuda()
    b = copy.deepcopy(a)
    return a, b


def _set_grads(ma, mb, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    for pa, pb in zip(ma.parameters(), mb.parameters()):
        gr = torch.randn(pa.shape, device="cuda", generator=g) * scale
        pa.grad, pb.grad = gr.clone(), gr.clone()


@pytest.mark.parametrize("wd", [0.0, 1e-2])
def test_fused_adam_matches_torch_adam(cuda, wd):
    from pdivgnn_b200.optim import FusedAdam
    ma, mb = _two_models()
    ref = torch.optim.Adam(ma.parameters(), lr=1e-3, weight_decay=wd)
    fus = FusedAdam(mb.parameters(), lr=1e-3, weight_decay=wd)
    for it in range(6):
        _set_grads(ma, mb, 100 + it, scale=10.0 ** (it - 3))
        ref.step()
        fus.step()
    for (k, pa), pb in zip(ma.named_parameters(), mb.parameters()):
        linf, l2 = H.rel_err(pb.detach().cpu(), pa.detach().cpu())
        assert linf < 2e-6 and l2 < 2e-6, (k, linf, l2)
    sa, sb = ref.state_dict(), fus.state_dict()
    assert sa["param_groups"][0]["params"] == sb["param_groups"][0]["params"]
    for i in sa["state"]:
        assert float(sa["state"][i]["step"]) == float(sb["state"][i]["step"]) == 6.0
        for key in ("exp_avg", "exp_avg_sq"):
            linf, _ = H.rel_err(sb["state"][i][key].cpu(), sa["state"][i][key].cpu())
            assert linf < 2e-6, (i, key, linf)


def test_fused_adam_state_dict_roundtrip_and_torch_interop(cuda):
    from pdivgnn_b200.optim import FusedAdam
    ma, mb = _two_models()
    ref = torch.optim.Adam(ma.parameters(), lr=1e-3)
    for it in range(3):
        _set_grads(ma, mb, 7 + it)
        ref.step()
    mb.load_state_dict(ma.state_dict())
    fus = FusedAdam(mb.parameters(), lr=5e-4)
    fus.load_state_dict(ref.state_dict())  # a torch.optim.Adam checkpoint loads into the fused optimizer
    assert fus.param_groups[0]["lr"] == 1e-3 and fus._steps == 3
    _set_grads(ma, mb, 99)
    ref.step()
    fus.step()
    for pa, pb in zip(ma.parameters(), mb.parameters()):
        linf, _ = H.rel_err(pb.detach().cpu(), pa.detach().cpu())
        assert linf < 2e-6


def test_fused_adam_skips_non_finite_step(cuda):
    from pdivgnn_b200.optim import FusedAdam
    ma, mb = _two_models()
    fus = FusedAdam(mb.parameters(), lr=1e-3, check_finite=True)
    _set_grads(ma, mb, 3)
    list(mb.parameters())[13].grad[5] = float("inf")
    before = [p.detach().clone() for p in mb.parameters()]
    fus.step(inv_scale=1.0 / 1024)
    assert int(fus.found_inf.item()) == 1
    assert all(torch.equal(a, b) for a, b in zip(before, mb.parameters()))
    assert not fus._exp_avg.any() and not fus._exp_avg_sq.any()
    assert fus.applied_steps() == 0  # GradScaler does not call optimizer.step() on overflow (gnn_train.py:205-207)
    _set_grads(ma, mb, 4)
    fus.step()
    assert int(fus.found_inf.item()) == 0
    assert any(not torch.equal(a, b) for a, b in zip(before, mb.parameters()))
    assert fus.applied_steps() == 1 and float(fus.state_dict()["state"][0]["step"]) == 1.0


def test_fused_adam_skipped_step_matches_gradscaler_trajectory(cuda):
    """A skipped step must leave the bias-correction count where GradScaler.step leaves torch.optim.Adam's: after
    [good, inf, good, good] the fused optimizer equals torch Adam stepped 3 times on the finite gradients."""
    from pdivgnn_b200.optim import FusedAdam
    ma, mb = _two_models()
    ref = torch.optim.Adam(ma.parameters(), lr=1e-3)
    fus = FusedAdam(mb.parameters(), lr=1e-3, check_finite=True)
    for it, poison in enumerate([False, True, False, False]):
        _set_grads(ma, mb, 11 + it)
        if poison:
            list(mb.parameters())[0].grad[0, 0] = float("nan")
        else:
            ref.step()
        fus.step()
    assert fus.applied_steps() == 3
    for pa, pb in zip(ma.parameters(), mb.parameters()):
        linf, _ = H.rel_err(pb.detach().cpu(), pa.detach().cpu())
        assert linf < 2e-6, linf


def test_fused_adam_rejects_foreign_parameter_lists(cuda):
    from pdivgnn_b200.optim import FusedAdam
    with pytest.raises(ValueError):
        FusedAdam([torch.nn.Parameter(torch.zeros(4, device="cuda"))])
    ma, _ = _two_models()
    fus = FusedAdam(ma.parameters())
    with pytest.raises(RuntimeError, match="gradient"):
        fus.step()


def test_zero_load_case_is_predicated_on_device(cuda):
    """All-zero mean_stress: zeros out (not un-standardised) and exactly-zero gradients, with no host sync
    (PDG_FLAG_ZERO_CHECK); a single non-zero entry anywhere in the batch switches the model back on."""
    samples, graphs, batch, stats = H.synthetic_batch(2, 300, seed0=21)
    sd = O.init_state_dict(seed=69)
    model = H.make_model(stats, params=sd)
    db = H.DeviceBatch(batch)
    live = model(db, scale_output=True).local_stress
    assert live.abs().max() > 0
    ms = db.mean_stress.clone()
    db.mean_stress = torch.zeros_like(ms)
    out = model(db, scale_output=True).local_stress
    assert out.shape == ms.shape and not out.any()
    (out.sum() * 3.0).backward()
    for k, p in model.named_parameters():
        assert p.grad is not None and not p.grad.any(), k
    db.mean_stress[-1, 2] = 1e-30
    again = model(db, scale_output=True).local_stress
    assert again.abs().max() > 0
