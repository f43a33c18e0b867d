"""The reference arm of bench.py (`--impl reference`: the oracle port timed on the host cores) needs no GPU: run it on a
tiny workload and check the one-line JSON contract the driver parses."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--batch", "2", "--nodes", "128",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must hold exactly one JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "nodes/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("mesh nodes/sec") and d["n_gpus"] == 1 and d["steps"] == 1 and d["scaling"] == "weak"
    assert d["value"] > 0 and abs(d["ms_per_step"] * 1e-3 * d["value"] - d["cpu_baseline"]["value"] * d["ms_per_step"] * 1e-3) < 1e-6
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["config"]["graphs_per_gpu"] == 2 and "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "nodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == ""


def test_flop_model_counts():
    """bench.flop_model: SURVEY 8d's algorithmic formula (forward = N 1 050 880 + E 2 654 464 at T = 10, training = 3 x) and the
    executed tile-GEMM count of one 128-node / 128-edge graph, by hand."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    f = m.flop_model(1000, 5600, 10, "bf16")
    assert f["algorithmic_per_step"] == 3 * (1000 * 1050880 + 5600 * 2654464)
    g = 2 * 128 ** 3  # one tile GEMM
    one = m.flop_model(128, 128, 10, "bf16")["executed_per_step"]
    assert one == g * (10 * (15 + 14) - 6 + 3 + 3 + 3)
    assert m.flop_model(128, 128, 10, "fp32")["executed_per_step"] == g * (10 * (15 + 11) - 3 + 3 + 3 + 3)
    assert m.flop_model(129, 1, 1, "bf16")["executed_per_step"] == g * (2 * 15 + 14 - 6 + 2 * 3 + 3 + 2 * 3)  # padding to whole tiles
    r = m.flops_of(5.0, "bf16", 33273, 193946, 10)
    assert 0.05 < r["tensor_frac_of_executed"] < 1.0 and r["executed_tflops"] < r["reference_equivalent_tflops"]
    r = m.flops_of(23.0, "fp32", 33273, 193946, 10)
    assert 0.1 < r["fp32_pipe_frac_of_executed"] < 1.0
