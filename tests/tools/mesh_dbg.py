import sys, os
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
import pdg_helpers as H
from oracle import pdg_oracle as O
import pdivgnn_b200
samples, graphs, batch, stats = H.synthetic_batch(2, 300)
sd = O.init_state_dict(seed=69)
model = H.make_model(stats, params=sd); model.precision = "bf16"
db = H.DeviceBatch(batch)
pred = model(db, scale_output=False).local_stress
nmse, dv = pdivgnn_b200.nmse_div_loss(pred, db, model, False, 10.0)
model.zero_grad(); (nmse + dv).backward()
g1 = {k: p.grad.clone() for k, p in model.named_parameters()}
model32 = H.make_model(stats, params=sd); model32.precision = "fp32"
pred = model32(db, scale_output=False).local_stress
nmse, dv = pdivgnn_b200.nmse_div_loss(pred, db, model32, False, 10.0)
(nmse + dv).backward()
for k, p in model32.named_parameters():
    g = g1[k].cpu(); r = p.grad.cpu()
    print(f"  {k:34s} nan {int(torch.isnan(g).sum()):6d} / {g.numel():6d}  l2err {H.rel_err(torch.nan_to_num(g), r)[1]:.2e}")
