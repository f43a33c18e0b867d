#!/bin/bash
# full 1-GPU check: all GPU tests, smoke, default bench line (timed), reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v Warning | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "smoke|Error|error" | tail -5
SECONDS=0
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
echo "default bench rc=$? wall=${SECONDS}s"
tail -3 gpurun_out/r2_bench_default.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_default.json"))
print("ms", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
print("roofline", {k: d["roofline"][k] for k in ("kernel","frac","us_per_launch","share_of_step")}, "agg", {k: d["roofline"]["aggregation_kernel"][k] for k in ("frac","us_per_launch")})
print("cpu", d["cpu_baseline"]); print("gpu_eager", d["gpu_eager_baseline"]); print("modes", d["modes"])
print("configs", json.dumps(d["configs"])[:3000])
PY
if [ "$1" = "ref" ]; then
SECONDS=0
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_ref.json 2>/dev/null
echo "reference arm rc=$? wall=${SECONDS}s"; cut -c1-400 gpurun_out/r2_bench_ref.json
fi
