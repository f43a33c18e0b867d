import sys, os, ctypes as C
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
import pdg_helpers as H
from oracle import pdg_oracle as O
from pdivgnn_b200 import _lib
samples, graphs, batch, stats = H.synthetic_batch(32, 1024)
model = H.make_model(stats, params=O.init_state_dict(seed=69)); model.precision = "bf16"
db = H.DeviceBatch(batch)
with torch.no_grad():
    for _ in range(3): model(db)
    torch.cuda.synchronize()
    L = _lib.lib(); buf = (C.c_ulonglong * 32)()
    L.pdg_phase_read_fwd(buf); base = list(buf)
    model(db); torch.cuda.synchronize()
    L.pdg_phase_read_fwd(buf); d = [b - a for a, b in zip(base, buf)]
names = ["wait E tile (producer)", "G mma wait", "hidden (both evaluations)", "sync + y1 mma wait", "y1: stage + segsum + stats", "y2: wait + stage + copy-out + stats", "end sync"]
tot = sum(d)
for n, v in zip(names, d): print(f"{n:40s} {v/110:9.0f} cyc/tile {100*v/tot:5.1f}%")
print("total", tot/110)
