import sys, os
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
import pdg_helpers as H
from oracle import pdg_oracle as O
import pdivgnn_b200
import test_gpu_edge_cases as T
for name in ("ring_degree1", "star_hub", "sparse_isolated"):
    n, edges = T.GRAPHS[name]()
    b = T._random_graph(n, edges, 7)
    stats = T._stats(); sd = O.init_state_dict(seed=69)
    model = H.make_model(stats, params=sd); model.precision = "bf16"
    db = T._device(b)
    pred = model(db, scale_output=False).local_stress
    nmse, _ = pdivgnn_b200.nmse_div_loss(pred, db, model, False, 0.0)
    nmse.backward()
    r = O.loss_and_grads(sd, b, stats, 10, False, 0.0, dtype=torch.float64)
    print(name, "E", len(edges[0]), "pred finite", bool(torch.isfinite(pred).all()))
    for k, p in model.named_parameters():
        g = p.grad.cpu()
        print(f"  {k:34s} nan {int(torch.isnan(g).sum()):6d} / {g.numel():6d}  l2err {H.rel_err(torch.nan_to_num(g), r[4][k])[1]:.2e}")
