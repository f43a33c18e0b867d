#!/bin/bash
# A/B/C bench on one box: current lib, lib_t, lib_old (each copied over lib/)
cp p-div-gnn_b200/lib/libpdivgnn.so /tmp/cur.so
run() { python bench.py --no-cpu-baseline --steps 40 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'ms', round(d['ms_per_step'],4), 'bwd_us', round(d['roofline']['us_per_launch'],1), 'fwd share', d['kernel_share_of_step']['edge_step'])"; }
for rep in 1 2; do
  cp /tmp/cur.so p-div-gnn_b200/lib/libpdivgnn.so; run current
  for v in lib_t lib_old; do
    if [ -f p-div-gnn_b200/$v/libpdivgnn.so ]; then cp p-div-gnn_b200/$v/libpdivgnn.so p-div-gnn_b200/lib/libpdivgnn.so; run $v; fi
  done
done
