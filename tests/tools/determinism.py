import sys, os
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
import pdg_helpers as H
from oracle import pdg_oracle as O
import pdivgnn_b200
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
samples, graphs, batch, stats = H.synthetic_batch(B, 1024)
sd = O.init_state_dict(seed=69)
model = H.make_model(stats, params=sd); model.precision = sys.argv[2] if len(sys.argv) > 2 else "bf16"
db = H.DeviceBatch(batch)
def run():
    model.zero_grad()
    pred = model(db, scale_output=False).local_stress
    nmse, dv = pdivgnn_b200.nmse_div_loss(pred, db, model, False, 10.0)
    (nmse + dv).backward()
    return pred.detach().clone(), {k: p.grad.clone() for k, p in model.named_parameters()}
p0, g0 = run()
for it in range(3):
    p1, g1 = run()
    bad = [(k, (g0[k] - g1[k]).abs().max().item() / (g0[k].abs().max().item() + 1e-30)) for k in g0 if not torch.equal(g0[k], g1[k])]
    print("run", it, "pred equal", torch.equal(p0, p1), "differing grads:", [(k, f"{v:.1e}") for k, v in bad])
