#!/bin/bash
# ncu --set full of the HBM-bound helper kernels inside one ViT-B step (LayerNorm fwd/bwd, column sums, casts)
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"ln_fwd|ln_bwd|colsum|scale_cast" -s 20 -c 8 -o gpurun_out/prof_ew_vit python bench.py --warmup 3 --nvtx-step > gpurun_out/ncu_ew_vit.log 2>&1
echo "vit exit=$?"
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"ln_fwd|ln_bwd|colsum|scale_cast" -s 8 -c 10 -o gpurun_out/prof_ew_swin python bench.py --workload swin_s --warmup 3 --nvtx-step > gpurun_out/ncu_ew_swin.log 2>&1
echo "swin exit=$?"
