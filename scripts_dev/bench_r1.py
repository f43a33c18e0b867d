#!/usr/bin/env python
"""Headline benchmark: mesh nodes / second of the P-GNN training step (fwd + loss + bwd +
Adam) on synthetic periodic 2-D RVE meshes -- BASELINE.json configs[1]
("config_train_no_div.yml ... synthetic meshes batch 32, 1 B200").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
  value        whole-job nodes/s, inputs resident in HBM, K steps, CUDA events, max over ranks
  e2e          same metric through the public API from HOST buffers: every step copies one
               pinned host batch (coordinates, triangles, fields) to the device, builds the
               graph there, runs fwd+loss+bwd+Adam and reads the loss back
  roofline     dominant kernel (edge_step_bwd), algorithmic HBM bytes / CUDA-event time
  cpu_baseline the oracle port (pure-torch restatement of the reference) on the host cores
`--impl reference` times that oracle port alone (the reference itself needs torch_geometric
and cannot be imported here or on the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

B_DEFAULT, NODES_DEFAULT, T_STEPS = 32, 1024, 10
CPU_SAMPLE_GRAPHS = 4  # bounded CPU sample: 4 meshes of the same generator (~4.2k nodes)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=B_DEFAULT, help="graphs per GPU")
    ap.add_argument("--nodes", type=int, default=NODES_DEFAULT, help="target nodes per mesh")
    ap.add_argument("--divergence", type=int, default=0, help="1 = config 4 (divergence regulariser on)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"],
                    help="bf16 = tcgen05 bf16 MLP tiles (2e-2 tolerance mode); fp32 = FFMA tiles (1e-5 mode)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------
# CPU arm: the oracle port (only place bench.py touches oracle/)
# ---------------------------------------------------------------------------------------
def cpu_oracle_step_rate(n_graphs, nodes, divergence, steps, warmup):
    from oracle import pdg_oracle as O
    from pdivgnn_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    samples = synth.make_dataset(n_graphs, nodes, 69)
    graphs = [O.build_graph(s, True) for s in samples]
    batch, stats = O.collate(graphs), O.dataset_stats(graphs)
    sd = O.init_state_dict(seed=69)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.Adam(list(params.values()), lr=1e-3)

    def step():
        total, nmse, div, pred = O.train_loss(params, batch, stats, T_STEPS, bool(divergence), 10.0)
        opt.zero_grad()
        total.backward()
        opt.step()
        return float(total)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return batch.num_nodes / dt, dt, batch.num_nodes, torch.get_num_threads()


_JSON_OUT = None


def guard_stdout():
    """The contract is ONE JSON line on stdout: everything else that writes to fd 1 (NCCL prints its version line
    there on the first communicator) goes to stderr; emit() writes the line to the real stdout."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def run_reference(args, rank):
    if rank != 0:
        return
    steps = max(1, min(args.steps, 20))
    warm = max(1, min(args.warmup, 3))
    rate, dt, n, cores = cpu_oracle_step_rate(CPU_SAMPLE_GRAPHS, args.nodes, args.divergence, steps, warm)
    sample = f"{CPU_SAMPLE_GRAPHS} meshes x ~{args.nodes} nodes ({n} nodes) per step, {steps} steps"
    line = {
        "impl": "reference", "metric": "mesh nodes/sec, P-GNN training step (fwd+loss+bwd+Adam)", "value": rate,
        "unit": "nodes/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": rate, "unit": "nodes/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "nodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "oracle port (pure-torch CPU restatement of the reference path); the reference needs torch_geometric",
    }
    emit(line)


def workload_config(args, world):
    return {"workload": f"configs[1]: P-GNN linear-elastic training step, divergence={'on' if args.divergence else 'off'}, "
                        f"{args.batch} synthetic periodic plate-with-hole meshes x ~{args.nodes} nodes per GPU, "
                        f"latent 128, {T_STEPS} message-passing steps, Adam lr 1e-3",
            "graphs_per_gpu": args.batch, "nodes_per_mesh": args.nodes, "gpus": world,
            "l2_policy": "per-step working set (saved state ~2.2 GB at batch 32) exceeds the 126 MB L2; no flush needed"}


# ---------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------
class Clocks:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.samples, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, l in self.samples:
            if ts < t0 - 0.05 or ts > t1 + 0.05:
                continue
            f = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------
def main():
    args = parse()
    guard_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    import torch.distributed as dist
    import pdivgnn_b200
    from pdivgnn_b200 import _lib, batcher, synth
    from pdivgnn_b200 import dist as pdist

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        pdist.init_from_env("nccl")
    L = _lib.lib()

    # ---- data: 4 distinct host batches per rank (e2e rotates through them), first one resident
    n_host = 4
    host = []
    for j in range(n_host):
        seed0 = 69 + (rank * n_host + j) * args.batch
        host.append(batcher.host_arrays(synth.make_dataset(args.batch, args.nodes, seed0)))
    with_op = bool(args.divergence)
    resident = batcher.batch_from_host(host[0], dev, True, with_op)
    stats = batcher.dataset_stats([resident])
    torch.manual_seed(69)
    model = pdivgnn_b200.EncodeProcessDecode(1, T_STEPS, 128, 6, 3, precision=args.precision, **stats).to(dev)
    if world > 1:
        pdist.broadcast_parameters(model)
        pdist.enable_data_parallel(model)
    from pdivgnn_b200.optim import FusedAdam
    opt = FusedAdam(model.parameters(), lr=1e-3)  # torch.optim.Adam semantics, one launch (pdg_adam_step)
    n_nodes, n_edges = resident.num_nodes, resident.edge_index.shape[1]

    def train_step(b):
        pred = model(b, scale_output=False, scale_input=True).local_stress
        nmse, dv = pdivgnn_b200.nmse_div_loss(pred, b, model, with_op, 10.0)
        loss = nmse + dv
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- kernel-resident timing -------------------------------------------------------
    for _ in range(max(3, args.warmup)):
        train_step(resident)
    barrier()
    L.pdg_launch_count(1)
    clocks = Clocks(local)
    time.sleep(0.25)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        train_step(resident)
    ev1.record()
    barrier()
    t1 = time.perf_counter()
    ms = ev0.elapsed_time(ev1) / args.steps
    launches = L.pdg_launch_count(1)
    clk = clocks.stop(t0, t1)

    # ---- per-kernel pass: the same steps again with CUDA events around every kernel class (library hooks).  Kept
    # out of the region above because an event record between two kernels disables their programmatic dependent
    # launch overlap; shares are taken against this pass's own step time.
    ksteps = max(3, min(args.steps, 10))
    L.pdg_timing_enable(1)
    _lib.timing_collect()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    k0.record()
    for _ in range(ksteps):
        train_step(resident)
    k1.record()
    barrier()
    ms_k = k0.elapsed_time(k1) / ksteps
    L.pdg_timing_enable(0)
    ktimes = _lib.timing_collect()

    # ---- end to end from host buffers -----------------------------------------------------
    h2d = batcher.host_bytes(host[0], with_op)

    # every step consumes a batch that starts in pinned HOST memory; its copies, device edge construction and
    # plan build run on a side stream, in a worker thread, two steps ahead (batcher.DevicePrefetcher), like a DataLoader worker
    pf = batcher.DevicePrefetcher(host, dev, True, with_op)

    # The loss of every step IS read back (4 bytes, pinned host buffer), but one step late: the copy of step j is
    # waited for after step j+1 has been enqueued, so the host never drains the GPU queue (a training loop that
    # logs the previous step's loss).
    loss_host = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]
    seen = []

    def e2e_step(j):
        b = pf.get()
        loss = train_step(b)  # enqueue the whole step ...
        loss_host[j & 1].copy_(loss.detach().reshape(1), non_blocking=True)  # D2H read of the step's loss
        loss_ev[j & 1].record()
        pf.prefetch()         # ... then stage the next host batch underneath it
        if j > 0:
            loss_ev[(j - 1) & 1].synchronize()
            seen.append(float(loss_host[(j - 1) & 1][0]))
        return b.num_nodes

    def e2e_drain(j_last):
        loss_ev[j_last & 1].synchronize()
        seen.append(float(loss_host[j_last & 1][0]))

    # warm-up: two full rotations over the host batches, so the caching allocator has seen every batch size
    # (a first-time cudaMalloc / cudaFree inside the timed region would stall the device)
    nw = max(2 * n_host, args.warmup)
    for j in range(nw):
        e2e_step(j)
    e2e_drain(nw - 1)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nodes_done = 0
    for j in range(args.steps):
        nodes_done += e2e_step(j)
    e2e_drain(args.steps - 1)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1) / args.steps
    pf.close()
    assert len(seen) >= args.steps and all(v == v for v in seen[-args.steps:]), "every step's loss must have been read back"

    # ---- max over ranks ---------------------------------------------------------------------
    tt = torch.tensor([ms, e2e_ms, float(n_nodes), float(nodes_done) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_ms = float(mx[0]), float(mx[1])
        tot_nodes, tot_nodes_e2e = float(sm[2]), float(sm[3])
    else:
        tot_nodes, tot_nodes_e2e = float(n_nodes), float(nodes_done) / args.steps
    value = tot_nodes / (ms * 1e-3)
    e2e_value = tot_nodes_e2e / (e2e_ms * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    top = max(ktimes.items(), key=lambda kv: kv[1][0])[0] if ktimes else None
    dom = top if top in ("edge_step_bwd", "edge_step") else ("edge_step_bwd" if "edge_step_bwd" in ktimes else top)
    roof = None
    if dom:
        tot_ms, cnt = ktimes[dom]
        per_launch_ms = tot_ms / cnt
        e_pad = (n_edges + 127) // 128 * 128
        # algorithmic bytes per launch, for the data types each path stores.
        #   fp32 path   edge_step_bwd per edge: read e_t, y2_t, ge_{t+1}, y_prev, write ge_t (5 x 512 B) + dhm, dhn rows written
        #               (2 x 512 B) + 8 B ids; per node: gathered Pa, Pb, g_agg rows (3 x 512 B) + RA, RB written (2 x 512 B)
        #               edge_step per edge: read e_{t-1}, y2_{t-1}, write e_t, y2_t (4 x 512 B) + 8 B ids; per node: Pa, Pb + aggraw
        #   bf16 path   raw edge-MLP outputs (y2, y_prev) are bf16 rows (256 B), the backward reads e_t as a bf16 operand-tile
        #               image (256 B) written by the forward, dhm / dhn / Pa / Pb / g_agg rows are bf16; ge and the e stream fp32:
        #               edge_step_bwd per edge: 256 (e image) + 256 (y2_t) + 256 (y_prev) + 512 + 512 (ge read, write)
        #               + 2 x 256 (dhm, dhn) + 8; per node 3 x 256 + 2 x 512
        #               edge_step (training) per edge: 512 + 256 read, 512 + 256 (image) + 256 (y2) written + 8; per node 2 x 256 + 512
        if args.precision == "bf16":
            alg = {"edge_step_bwd": e_pad * (3 * 256 + 2 * 512 + 2 * 256 + 8) + n_nodes * (3 * 256 + 2 * 512),
                   "edge_step": e_pad * (512 + 256 + 512 + 256 + 256 + 8) + n_nodes * (2 * 256 + 512)}.get(dom)
        else:
            alg = {"edge_step_bwd": e_pad * (5 * 512 + 2 * 512 + 8) + n_nodes * (3 * 512 + 2 * 512),
                   "edge_step": e_pad * (4 * 512 + 8) + n_nodes * (2 * 512 + 512)}.get(dom)
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(
                f"{dom}:{args.precision}")
        except Exception:
            pass
        if alg:
            ach = alg / (per_launch_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg, "us_per_launch": per_launch_ms * 1e3,
                    "share_of_step": tot_ms / ksteps / ms_k, "steps_in_kernel_pass": ksteps, "ms_per_step_kernel_pass": ms_k,
                    "note": ("warp-specialised tcgen05 bf16 tile kernel: bound by L1 gather throughput and epilogue latency, "
                             "not by HBM or the tensor pipe (DESIGN.md section 4b, profiles/r1_tc_kernels_full_final.md)"
                             if args.precision == "bf16" else
                             "fp32 FFMA tile path: compute-bound, far from the HBM roof (see DESIGN.md)")}
    kshare = {k: round(v[0] / ksteps / ms_k, 4) for k, v in sorted(ktimes.items(), key=lambda kv: -kv[1][0])}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        rate, dt, n, cores = cpu_oracle_step_rate(CPU_SAMPLE_GRAPHS, args.nodes, args.divergence, 3, 1)
        cpu = {"value": rate, "unit": "nodes/s", "cores": cores, "kind": "port",
               "sample": f"{CPU_SAMPLE_GRAPHS} meshes x ~{args.nodes} nodes ({n} nodes) per step, 1 warm-up + 3 steps, "
                         f"{dt:.2f} s/step, oracle port (pure torch)"}

    line = {
        "metric": "mesh nodes/sec, P-GNN training step (fwd+loss+bwd+Adam)", "value": value, "unit": "nodes/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {**workload_config(args, world), "nodes_per_gpu": n_nodes, "edges_per_gpu": n_edges,
                   "precision_mode": ("bf16 MLP-tile operands on tcgen05, fp32 accumulate/LayerNorm/latents (tolerance 2e-2)"
                                      if args.precision == "bf16" else "fp32 FFMA tiles (tolerance 1e-5)")},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "nodes/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms},
        "gpu_launches": int(launches),
        "roofline": roof, "cpu_baseline": cpu, "kernel_share_of_step": kshare,
        "published_reference_gpu_forward_nodes_per_s": 63000,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
