#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:attn_tc -s 4 -c 14 -o gpurun_out/prof_vit_attn python bench.py --workload vit_b16 --warmup 3 --nvtx-step > gpurun_out/ncu_vit_attn.log 2>&1
echo "exit=$?"
