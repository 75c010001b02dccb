#!/bin/bash
# --set full of the ViT-B fc1 weight gradient without / with the fused bias gradient (a_colsum on CTA-pair tiles), and of the
# QKV one; then the whole GPU suite on the closing tree
mkdir -p gpurun_out
GEMM_NOCHECK=1 GEMM_BLOCK=vitb GEMM_ONLY="fc1 wgrad  " bash tools/ncu_cabi.sh vitb_fc1_wgrad gemm_tc_kernel 1 tools/cabi_gemm_bench.py > /dev/null
GEMM_NOCHECK=1 GEMM_BLOCK=vitb GEMM_ONLY="fc1 wgrad + colsum" bash tools/ncu_cabi.sh vitb_fc1_wgrad_colsum gemm_tc_kernel 1 tools/cabi_gemm_bench.py > /dev/null
GEMM_NOCHECK=1 GEMM_BLOCK=vitb GEMM_ONLY="qkv wgrad + colsum" bash tools/ncu_cabi.sh vitb_qkv_wgrad_colsum gemm_tc_kernel 1 tools/cabi_gemm_bench.py > /dev/null
rm -f gpurun_out/*.ncu-rep
head -12 gpurun_out/ncu_vitb_fc1_wgrad.txt; head -12 gpurun_out/ncu_vitb_fc1_wgrad_colsum.txt; head -5 gpurun_out/ncu_vitb_qkv_wgrad_colsum.txt
timeout 600 python -m pytest tests/ -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -2 | tee gpurun_out/t_all_closing.log
