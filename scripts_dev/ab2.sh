#!/bin/bash
# A/B on one box: lib (current build) vs lib_old (previous commit), then phase timers from lib_t
cp p-div-gnn_b200/lib/libpdivgnn.so /tmp/cur.so
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run() { python bench.py --no-cpu-baseline --steps 40 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'bwd_us', round(d['roofline']['us_per_launch'],1), {k: round(v*d['roofline']['ms_per_step_kernel_pass']*100,1) for k,v in list(d['kernel_share_of_step'].items())[:6]})"; }
for rep in 1 2; do
  cp /tmp/cur.so p-div-gnn_b200/lib/libpdivgnn.so; run current
  for v in lib_b lib_old; do
    if [ -f p-div-gnn_b200/$v/libpdivgnn.so ]; then cp p-div-gnn_b200/$v/libpdivgnn.so p-div-gnn_b200/lib/libpdivgnn.so; run $v; fi
  done
done
if [ -f p-div-gnn_b200/lib_t/libpdivgnn.so ]; then
  cp p-div-gnn_b200/lib_t/libpdivgnn.so p-div-gnn_b200/lib/libpdivgnn.so
  python tests/tools/phases.py 2>&1 | tail -24; python tests/tools/phases_fwd.py 2>&1 | tail -10
fi
cp /tmp/cur.so p-div-gnn_b200/lib/libpdivgnn.so
