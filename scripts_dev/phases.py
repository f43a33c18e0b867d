import sys, os, ctypes as C
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
import pdg_helpers as H
from oracle import pdg_oracle as O
import pdivgnn_b200
from pdivgnn_b200 import _lib
samples, graphs, batch, stats = H.synthetic_batch(32, 1024)
sd = O.init_state_dict(seed=69)
model = H.make_model(stats, params=sd); model.precision = "bf16"
db = H.DeviceBatch(batch)
def step():
    pred = model(db, scale_output=False).local_stress
    nmse, dv = pdivgnn_b200.nmse_div_loss(pred, db, model, False, 10.0)
    model.zero_grad(); (nmse + dv).backward()
for _ in range(3): step()
torch.cuda.synchronize()
L = _lib.lib()
buf = (C.c_ulonglong * 32)()
L.pdg_phase_read(buf); base = list(buf)
step(); torch.cuda.synchronize()
L.pdg_phase_read(buf); d = [b - a for a, b in zip(base, buf)]
names = ["E load+sync","G wait","hidden+sync","y1 wait","dy1+sync","colsum dy1","c3 wait","dhm epi+sync","segsum RA","dy2 build+sync","colsum dy2","c4 wait","dhn epi+sync","segsum RB","dG+sync","colsum dG","c5 wait","de epi+S32 w+sync","colsum1+sync","S32 w2+sync","colsum2+endsync"]
tot = sum(d)
print("tiles handled by block 0 per launch: ~11, launches 10; total cycles", tot)
for n, v in zip(names, d): print(f"{n:22s} {v/110:9.0f} cyc/tile  {100*v/tot:5.1f}%")
