#!/bin/bash
# a_colsum ring service as a non-blocking state machine: parity, then the weight-gradient microbench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "wgrad or gemm" 2>&1 | tail -2
GEMM_BLOCK=vitb,swin3,swin1 GEMM_ONLY="wgrad" timeout 200 python tools/cabi_gemm_bench.py 2>&1 | grep -v "self-check" | tee gpurun_out/cabi_gemm_colsum_poll.log
