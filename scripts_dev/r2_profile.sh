#!/bin/bash
# round-2 final-state evidence: default bench line, divergence / batch-16 / fp32 lines, reference arm, ncu launch list of the
# same command, ncu --set full of the six processor kernels.  Everything lands in gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
SECONDS=0
timeout 900 python bench.py > gpurun_out/r2_final_bf16.json 2> gpurun_out/r2_final_bf16.err; echo "default rc=$? t=${SECONDS}s"
timeout 600 python bench.py --divergence 1 --lean --no-cpu-baseline > gpurun_out/r2_final_div.json 2>/dev/null; echo "div rc=$?"
timeout 600 python bench.py --batch 16 --lean --no-cpu-baseline > gpurun_out/r2_final_b16.json 2>/dev/null; echo "b16 rc=$?"
timeout 600 python bench.py --precision fp32 --steps 10 --lean --no-cpu-baseline > gpurun_out/r2_final_fp32.json 2>/dev/null; echo "fp32 rc=$?"
timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2_final_ref.json 2>/dev/null; echo "ref rc=$? t=${SECONDS}s"
timeout 900 python bench.py --config 4 > gpurun_out/r2_final_config4_1gpu.json 2>/dev/null; echo "config4 rc=$? t=${SECONDS}s"
# launch list of the bench command (2 timed steps)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 2 --warmup 3 --lean --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$? t=${SECONDS}s"
# full captures: skip the warm-up launches of each kernel, take one step's worth
K='regex:k_edge_step_bwd_tc3|k_edge_step_tc|k_node_update_bwd_tc|k_node_pre_bwd_tc|k_node_update_tc|k_node_pre_tc'
timeout 1200 ncu --set full --clock-control none --import-source on -k "$K" -s 200 -c 12 -f -o gpurun_out/r2_full \
  python bench.py --steps 2 --warmup 3 --lean --no-cpu-baseline > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu full rc=$? t=${SECONDS}s"
ls -la gpurun_out/r2_* | head -20
