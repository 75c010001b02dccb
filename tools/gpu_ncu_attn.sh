#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:attn_wp_bwd -s 6 -c 1 -o gpurun_out/prof_wp_bwd python bench.py --workload swin_s --warmup 3 --nvtx-step > gpurun_out/ncu_wp_bwd.log 2>&1
echo "bwd exit=$?"
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:attn_wp_fwd -s 6 -c 1 -o gpurun_out/prof_wp_fwd python bench.py --workload swin_s --warmup 3 --nvtx-step > gpurun_out/ncu_wp_fwd.log 2>&1
echo "fwd exit=$?"
