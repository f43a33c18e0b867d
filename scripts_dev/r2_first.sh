#!/bin/bash
# round 2, first GPU visit: new bf16 parity tests, smoke in both precisions, phase timers, short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py -q -s 2>&1 | grep -v Warning | tail -60 > gpurun_out/r2_bf16_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
cp p-div-gnn_b200/lib/libpdivgnn.so /tmp/cur.so
if [ -f p-div-gnn_b200/lib_t/libpdivgnn.so ]; then
  cp p-div-gnn_b200/lib_t/libpdivgnn.so p-div-gnn_b200/lib/libpdivgnn.so
  timeout 300 python tests/tools/phases.py 2>&1 | tail -24 > gpurun_out/r2_phases.log; timeout 300 python tests/tools/phases_fwd.py 2>&1 | tail -10 >> gpurun_out/r2_phases.log
fi
cp /tmp/cur.so p-div-gnn_b200/lib/libpdivgnn.so
timeout 600 python bench.py --no-cpu-baseline --steps 30 > gpurun_out/r2_bench0.json 2> gpurun_out/r2_bench0.err
tail -30 gpurun_out/r2_bf16_tests.log; cat gpurun_out/r2_smoke.log | tail -5; cat gpurun_out/r2_phases.log
