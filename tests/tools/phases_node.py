"""clock64 phase timers of the node backward kernels (library built with -DPDG_PHASE_TIMERS into lib_t)."""
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
buf = (C.c_ulonglong * 64)()
L.pdg_phase_read_node(buf); base = list(buf)
step(); torch.cuda.synchronize()
L.pdg_phase_read_node(buf); d = [b - a for a, b in zip(base, buf)]
nl = 10  # launches per step; CTA 0 runs 2 tiles per launch: slots [0,32) first tile, [32,64) second tile (+ the per-launch flush)
names_u = ["loop top", "load dy3/hq -> T0/T1 + sync", "colsum dy3 + mma wait + sync", "dhq epilogue + sync", "load agg/x_t -> T1/T2 + sync",
           "colsum dhq + mma wait", "g_agg tmem->s32 + sync", "g_agg pass (aggraw, gagg store)", "mma wait + tmem->s32 + sync", "gx RMW pass + sync", "flush (per launch)"]
names_p = ["loop top", "sender lists -> smem + 2 syncs", "RA/RB + sender gather -> T0/T1", "x_t -> T2 + sync", "mma wait", "tmem->s32 + sync",
           "gx RMW pass + RA/RB zero + sync", "flush (per launch)"]
for title, names, o in (("k_node_update_bwd_tc", names_u, 0), ("k_node_pre_bwd_tc", names_p, 16)):
    print(title)
    for j, n in enumerate(names):
        print(f"  {n:44s} tile 1: {d[o + j]/nl:8.0f}   tile 2: {d[32 + o + j]/nl:8.0f} cyc")
    print("  total cyc/launch", (sum(d[o:o + 16]) + sum(d[32 + o:48 + o])) / nl)
