#!/bin/bash
# round 2, session 3: a_colsum on CTA-pair tiles — parity first, then the ViT-B weight-gradient A/B through the C-ABI
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "wgrad or gemm" 2>&1 | tail -5
echo "== pair tiles (new)"; GEMM_BLOCK=vitb,swin3 GEMM_ONLY="wgrad,colsum" timeout 300 python tools/cabi_gemm_bench.py 2>&1 | tee gpurun_out/cabi_gemm_colsum_pair.log
echo "== 1-CTA tiles (round-1 rule)"; VTB_OPTS=gemm_colsum_pair=0 GEMM_BLOCK=vitb,swin3 GEMM_ONLY="+ colsum" timeout 300 python tools/cabi_gemm_bench.py 2>&1 | tee gpurun_out/cabi_gemm_colsum_1cta.log
