#!/bin/bash
# usage: r2_ab.sh "<pytest args>" [bench args]   -- GPU tests, then A/B bench of lib vs lib_old (if present)
mkdir -p gpurun_out
timeout 1200 python -m pytest $1 -q -s 2>&1 | grep -v Warning > gpurun_out/ab_tests.log
grep -E "worst tensor|trajectory:|configs\[1\]|rel err|bf16 vs|passed|failed|FAILED|Error" gpurun_out/ab_tests.log | cut -c1-400 | tail -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "smoke|Error|error" | tail -5
cp p-div-gnn_b200/lib/libpdivgnn.so /tmp/cur.so
run() { timeout 600 python bench.py --no-cpu-baseline --lean --steps 40 $2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'bwd_us', round(d['roofline']['us_per_launch'],1), {k: round(v*d['roofline']['ms_per_step_kernel_pass']*100,1) for k,v in list(d['kernel_share_of_step'].items())[:8]})"; }
for rep in 1 2; do
  cp /tmp/cur.so p-div-gnn_b200/lib/libpdivgnn.so; run current "$2"
  if [ -f p-div-gnn_b200/lib_old/libpdivgnn.so ]; then cp p-div-gnn_b200/lib_old/libpdivgnn.so p-div-gnn_b200/lib/libpdivgnn.so; run lib_old "$2"; fi
done
cp /tmp/cur.so p-div-gnn_b200/lib/libpdivgnn.so
