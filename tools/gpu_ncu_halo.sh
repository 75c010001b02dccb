#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"attn_res_bwd|attn_halo_dkv" -s 4 -c 2 -o gpurun_out/prof_halo_bwd python bench.py --workload halo_t --warmup 3 --nvtx-step > gpurun_out/ncu_halo_bwd.log 2>&1
echo "exit=$?"
