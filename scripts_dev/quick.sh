#!/bin/bash
# quick GPU check: bf16 parity tests, short bench, phase timers (timers build in lib_t)
python -m pytest tests/test_gpu_bf16.py tests/test_gpu_edge_cases.py tests/test_gpu_backward.py -x -q 2>&1 | tail -2
python bench.py --no-cpu-baseline --steps 30 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms', d['ms_per_step'], 'nodes/s', d['value'], 'e2e', d['e2e']['value'], 'bwd_us', d['roofline']['us_per_launch']); print(d['kernel_share_of_step'])"
if [ -f p-div-gnn_b200/lib_t/libpdivgnn.so ]; then
  cp p-div-gnn_b200/lib_t/libpdivgnn.so p-div-gnn_b200/lib/libpdivgnn.so
  python tests/tools/phases.py 2>&1 | tail -15; python tests/tools/phases_fwd.py 2>&1 | tail -8
fi
