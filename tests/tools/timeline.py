"""GPU timeline of one training step (torch.profiler / CUPTI): per-kernel totals and idle gaps."""
import sys, os, json
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
import pdg_helpers as H
from oracle import pdg_oracle as O
import pdivgnn_b200
from torch.profiler import profile, ProfilerActivity
samples, graphs, batch, stats = H.synthetic_batch(32, 1024)
model = H.make_model(stats, params=O.init_state_dict(seed=69)); model.precision = "bf16"
from pdivgnn_b200.optim import FusedAdam
opt = FusedAdam(model.parameters(), lr=1e-3) if os.environ.get('OPT', 'fused') == 'fused' else torch.optim.Adam(model.parameters(), lr=1e-3)
db = H.DeviceBatch(batch)
def step():
    pred = model(db, scale_output=False).local_stress
    nmse, dv = pdivgnn_b200.nmse_div_loss(pred, db, model, False, 10.0)
    opt.zero_grad(set_to_none=True)
    (nmse + dv).backward()
    opt.step()
for _ in range(5): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, evs[-1].time_range.end
busy = sum(e.time_range.end - e.time_range.start for e in evs)
print(f"span {(t1-t0)/3:.0f} us/step  busy {busy/3:.0f} us/step  kernels/step {len(evs)/3:.0f}")
agg = {}
gaps = {}
prev = None
for e in evs:
    n = e.name[:60]
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += e.time_range.end - e.time_range.start
    if prev is not None:
        g = e.time_range.start - prev.time_range.end
        if g > 0:
            k = prev.name[:40] + " -> " + e.name[:40]
            gg = gaps.setdefault(k, [0, 0.0]); gg[0] += 1; gg[1] += g
    prev = e
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{t/3:9.1f} us/step  {c/3:6.1f} x {t/c:7.1f} us  {n}")
print("--- gaps")
for n, (c, t) in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{t/3:9.1f} us/step  {c/3:6.1f} x {t/c:7.1f} us  {n}")
