import sys, os, time
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
import pdg_helpers as H
from oracle import pdg_oracle as O
from pdivgnn_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
samples, graphs, batch, stats = H.synthetic_batch(B, 1024)
sd = O.init_state_dict(seed=69)
model = H.make_model(stats, params=sd); model.precision = prec
db = H.DeviceBatch(batch)
print("N", batch.num_nodes, "E", batch.edge_index.shape[1], prec)
L = _lib.lib()
with torch.no_grad():
    for _ in range(3):
        model(db)
    torch.cuda.synchronize()
    L.pdg_timing_enable(1); _lib.timing_collect()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    K = 10
    for _ in range(K):
        model(db)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / K
    kt = _lib.timing_collect()
print(f"forward {ms:.3f} ms -> {batch.num_nodes/ms*1e3/1e6:.2f} M nodes/s")
print({k: round(v[0]/v[1]*1e3, 1) for k, v in kt.items()}, "us per launch")
