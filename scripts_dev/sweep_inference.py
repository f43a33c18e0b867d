"""BASELINE.json configs[2]: inference throughput sweep without periodic edges, batch 1..1024 graphs."""
import json, os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch
import pdivgnn_b200
from pdivgnn_b200 import batcher, synth
out = []
pool = synth.make_dataset(64, 1024, 1000)  # 64 distinct meshes, reused cyclically for large batches
for prec in ("bf16", "fp32"):
    for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024):
        samples = [pool[i % len(pool)] for i in range(B)]
        mb = batcher.batch_from_host(batcher.host_arrays(samples), "cuda", periodic=False, with_op_div=False)
        stats = batcher.dataset_stats([mb])
        torch.manual_seed(69)
        model = pdivgnn_b200.EncodeProcessDecode(1, 10, 128, 6, 3, precision=prec, **stats).to("cuda")
        with torch.no_grad():
            for _ in range(3):
                model(mb)
            torch.cuda.synchronize()
            K = 20 if B <= 64 else 5
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(K):
                model(mb)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        rec = dict(precision=prec, graphs=B, nodes=mb.num_nodes, edges=int(mb.edge_index.shape[1]), ms=ms,
                   nodes_per_s=mb.num_nodes / ms * 1e3, peak_mem_gb=torch.cuda.max_memory_allocated() / 1e9)
        print(rec, flush=True)
        out.append(rec)
        del model, mb
        torch.cuda.empty_cache()
json.dump(out, open(os.path.join(R, "gpurun_out", "inference_sweep.json"), "w"), indent=1)
