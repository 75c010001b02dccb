#!/bin/bash
# ncu --set full of the persistent global-attention kernels on the ViT-B shape (torch-free harness)
#   bash tools/gpu_ncu_attn.sh <tag> <kernel regex>
tag=$1; regex=$2
ATTN_SKIP_CHECK=1 ATTN_ONLY=global TMO=240 bash tools/ncu_cabi.sh $tag $regex 1 tools/cabi_attn_bench.py
ncu -i gpurun_out/ncu_$tag.ncu-rep --page source --csv > gpurun_out/ncu_${tag}_source.csv 2>/dev/null
python tools/ncu_src_top.py gpurun_out/ncu_${tag}_source.csv 60 > gpurun_out/ncu_${tag}_src_top.txt
rm -f gpurun_out/ncu_${tag}_source.csv
tail -70 gpurun_out/ncu_${tag}_src_top.txt
