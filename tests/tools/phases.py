import sys, os, ctypes as C
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
L.pdg_phase_read3(buf); base = list(buf)
step(); torch.cuda.synchronize()
L.pdg_phase_read3(buf); d = [b - a for a, b in zip(base, buf)]
names = ["wait E tile (producer)", "G mma wait", "hidden (message)", "sync + y1 mma wait", "dy1 epilogue", "sync + colsum dy1 + mma wait",
         "dhidden (dhm)", "sync + segsum RA + colsum + mma wait", "sync + hidden (update)", "wait dy2 (producer)",
         "sync + colsum dy2 + mma wait", "dhidden (dhn)", "sync + segsum RB + colsum + mma wait", "de staging + sync"]
tot = sum(d[:16])
ntile = 11 * 9 + 10  # CTA 0 owns 11 tiles (1516 tiles / 148 CTAs), 10 steps of which the last skips the update path
for n, v in zip(names, d):
    print(f"{n:40s} {v/110:9.0f} cyc/tile  {100*v/tot:5.1f}%")
print("total cyc/tile", sum(d[:16])/110)
pn = ["other -> before dy2_build", "ids(j+1)", "dy2_build body", "other -> before de-staged wait", "wait de staged", "final_pass body", "wait dy2 tile consumed"]
for n, v in zip(pn, d[16:23]): print(f"  producer: {n:36s} {v/110:9.0f} cyc/tile")
