#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu --no-header -p no:cacheprovider -x -k "gemm" 2>&1 | tail -2
GEMM_BLOCK=${GEMM_BLOCK:-swin3,vitb} timeout 200 python tools/cabi_gemm_bench.py 2>&1 | grep -E "swin-s3|vit-b|FAIL|PASS" | tee gpurun_out/cabi_gemm.log | cut -c1-120
