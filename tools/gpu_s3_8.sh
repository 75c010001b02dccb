#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_kernels_gpu.py -q -m gpu --no-header -p no:cacheprovider -x -k "gemm" 2>&1 | tail -3
timeout 100 python tools/trace_gemm.py 2>&1 | grep -v Warn
timeout 120 python tools/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1; cat gpurun_out/bench_gemm.log
