import sys, os
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
import pdg_helpers as H
from oracle import pdg_oracle as O
import pdivgnn_b200
import test_gpu_edge_cases as T
for hub in (60, 130, 260, 499):
    n = hub + 1
    edges = [list(range(1, n)) + [0], [0] * hub + [1]]
    b = T._random_graph(n, edges, 7)
    stats = T._stats(); sd = O.init_state_dict(seed=69)
    out = {}
    for prec in ("fp32", "bf16"):
        model = H.make_model(stats, params=sd); model.precision = prec
        db = T._device(b)
        pred = model(db, scale_output=False).local_stress
        nmse, _ = pdivgnn_b200.nmse_div_loss(pred, db, model, False, 0.0)
        nmse.backward()
        out[prec] = {k: p.grad.cpu() for k, p in model.named_parameters()}
    r = O.loss_and_grads(sd, b, stats, 10, False, 0.0, dtype=torch.float64)
    cat = lambda d: torch.cat([d[k].double().flatten() for k in O.STATE_KEYS])
    print("hub", hub, "fp32 flat", H.rel_err(cat(out["fp32"]), cat(r[4])), "bf16 flat", H.rel_err(cat(out["bf16"]), cat(r[4])))
    if hub == 499:
        for k in O.STATE_KEYS:
            print(f"  {k:34s} bf16 {H.rel_err(out['bf16'][k], r[4][k])[1]:.2e}  fp32 {H.rel_err(out['fp32'][k], r[4][k])[1]:.2e}  |g| {r[4][k].norm():.2e}")
