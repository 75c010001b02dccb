#!/bin/bash
# epilogue TMA helper lanes on one warp vs two warps: C-ABI microbench A/B
mkdir -p gpurun_out
for h in 1 2; do
echo "== gemm_helpers=$h"; VTB_OPTS=gemm_helpers=$h GEMM_BLOCK=swin3,vitb,swin1 timeout 300 python tools/cabi_gemm_bench.py 2>&1 | grep -v "colsum" | tee gpurun_out/cabi_gemm_helpers$h.log
done
