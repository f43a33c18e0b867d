import sys, os, time
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
import pdg_helpers as H
from oracle import pdg_oracle as O
import pdivgnn_b200
from pdivgnn_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
div = int(sys.argv[2]) if len(sys.argv) > 2 else 0
prec = sys.argv[3] if len(sys.argv) > 3 else "fp32"
samples, graphs, batch, stats = H.synthetic_batch(B, 1024)
sd = O.init_state_dict(seed=69)
model = H.make_model(stats, params=sd); model.precision = prec
model.skip_zero_check = bool(int(os.environ.get("SKIPZ", "0")))
from pdivgnn_b200.optim import FusedAdam
opt = FusedAdam(model.parameters(), lr=1e-3) if os.environ.get('OPT', 'fused') == 'fused' else torch.optim.Adam(model.parameters(), lr=1e-3)
db = H.DeviceBatch(batch)
print("N", batch.num_nodes, "E", batch.edge_index.shape[1], prec)
def step():
    pred = model(db, scale_output=False).local_stress
    nmse, dv = pdivgnn_b200.nmse_div_loss(pred, db, model, bool(div), 10.0)
    opt.zero_grad(set_to_none=True)
    (nmse + dv).backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
L = _lib.lib(); L.pdg_timing_enable(1); _lib.timing_collect()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record(); K = 10
for _ in range(K): step()
ev[1].record(); torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / K
kt = _lib.timing_collect()
print(f"train step {ms:.3f} ms -> {batch.num_nodes/ms*1e3/1e6:.2f} M nodes/s  mem {torch.cuda.max_memory_allocated()/1e9:.2f} GB")
print({k: round(v[0]/v[1]*1e3, 1) for k, v in kt.items()}, "us per launch")
print({k: round(v[0]/K, 3) for k, v in kt.items()}, "ms per step")
