"""Where does an epoch over an on-disk dataset spend its time?  Per-batch build time in the staging thread, time the
training thread waits for a batch, and the step time, for the resident and the host-collation loaders."""
import sys, os, time, tempfile
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R)
import numpy as np, torch
import pdivgnn_b200
from pdivgnn_b200 import batcher, io as pio, synth, engine
from pdivgnn_b200.optim import FusedAdam
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
folder = os.path.join(tempfile.gettempdir(), f"pdg_lt_{n}")
if not os.path.exists(os.path.join(folder, "dataset.csv")):
    base = synth.make_dataset(32, 1024, 69)
    pio.write_dataset([base[i % 32] for i in range(n)], folder)
ds = pio.MeshStressFieldDataset(os.path.join(folder, "dataset.csv"), device="cuda")
torch.manual_seed(69)
model = pdivgnn_b200.EncodeProcessDecode(1, 10, 128, 6, 3, precision="bf16", **ds.stats()).cuda()
opt = FusedAdam(model.parameters(), lr=1e-3)
builds, waits = [], []
ob, og = batcher.DevicePrefetcher._build, batcher.DevicePrefetcher.get
def tb(self, j):
    t = time.perf_counter(); r = ob(self, j); torch.cuda.current_stream().synchronize() if False else None
    builds.append(time.perf_counter() - t); return r
def tg(self):
    t = time.perf_counter(); r = og(self); waits.append(time.perf_counter() - t); return r
batcher.DevicePrefetcher._build, batcher.DevicePrefetcher.get = tb, tg
for resident in (True, False):
    loader = ds.loader(32, shuffle=True, with_op_div=False, resident=resident)
    engine.train_epoch(model, loader, opt, False, 10.0)
    torch.cuda.synchronize(); builds.clear(); waits.clear()
    t0 = time.perf_counter(); engine.train_epoch(model, loader, opt, False, 10.0); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"resident={resident}: {dt/len(loader)*1e3:.2f} ms/step; build mean {np.mean(builds)*1e3:.2f} ms (max {np.max(builds)*1e3:.2f}); "
          f"get() wait mean {np.mean(waits)*1e3:.2f} ms (max {np.max(waits)*1e3:.2f})")
