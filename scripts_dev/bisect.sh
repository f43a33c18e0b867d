#!/bin/bash
# run each library variant under tight timeouts (a deadlocked kernel must not eat the GPU budget)
cp p-div-gnn_b200/lib/libpdivgnn.so /tmp/cur.so
run() { timeout 80 python bench.py --no-cpu-baseline --steps 30 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k: round(v*d['roofline']['ms_per_step_kernel_pass']*100,1) for k,v in list(d['kernel_share_of_step'].items())[:6]})" || echo "$1 FAILED/TIMEOUT"; }
for v in "$@"; do
  cp p-div-gnn_b200/$v/libpdivgnn.so p-div-gnn_b200/lib/libpdivgnn.so
  timeout 100 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_edge_cases.py tests/test_gpu_backward.py -x -q 2>&1 | tail -1
  run $v
done
cp /tmp/cur.so p-div-gnn_b200/lib/libpdivgnn.so
