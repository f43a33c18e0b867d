#!/bin/bash
# A/B of two library builds on value and end-to-end time for several configs (tight timeouts)
cp p-div-gnn_b200/lib/libpdivgnn.so /tmp/cur.so
run() { timeout 80 python bench.py --no-cpu-baseline $2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 [$2]', 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'bwd_us', round(d['roofline']['us_per_launch'],1))" || echo "$1 $2 FAILED"; }
for v in cur lib_old; do
  if [ $v = cur ]; then cp /tmp/cur.so p-div-gnn_b200/lib/libpdivgnn.so; else cp p-div-gnn_b200/$v/libpdivgnn.so p-div-gnn_b200/lib/libpdivgnn.so; fi
  run $v ""; run $v "--batch 16 --divergence 1"; run $v "--divergence 1"
done
cp /tmp/cur.so p-div-gnn_b200/lib/libpdivgnn.so
timeout 100 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_backward.py tests/test_gpu_edge_cases.py -x -q 2>&1 | tail -1
