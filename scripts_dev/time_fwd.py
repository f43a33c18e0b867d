import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import pdg_helpers as H
from oracle import pdg_oracle as O
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
samples, graphs, batch, stats = H.synthetic_batch(B, 1024)
sd = O.init_state_dict(seed=69)
model = H.make_model(stats, params=sd)
db = H.DeviceBatch(batch)
print("N", batch.num_nodes, "E", batch.edge_index.shape[1])
with torch.no_grad():
    for _ in range(3):
        model(db)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    K = 10
    for _ in range(K):
        model(db)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / K
print(f"forward {ms:.3f} ms -> {batch.num_nodes/ms*1e3/1e6:.2f} M nodes/s")
