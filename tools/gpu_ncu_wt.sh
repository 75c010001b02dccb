#!/bin/bash
mkdir -p gpurun_out
export WT_ONLY=1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"attn_wt" -s 13 -c 2 -o gpurun_out/prof_wt python tools/bench_winattn.py > gpurun_out/ncu_wt.log 2>&1
echo "ncu exit=$?"; tail -3 gpurun_out/ncu_wt.log
