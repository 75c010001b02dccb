#!/bin/bash
mkdir -p gpurun_out
python tools/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1; cat gpurun_out/bench_gemm.log
GEMM_BLOCK=swin3 GEMM_ONLY="proj fwd" timeout 300 ncu --set full --import-source on --clock-control none -k regex:gemm_tc -s 5 -c 1 -o gpurun_out/prof_gemm_proj python tools/bench_gemm.py > gpurun_out/ncu_gemm1.log 2>&1; echo "ncu1 $?"
GEMM_BLOCK=swin3 GEMM_ONLY="fc1  fwd" timeout 300 ncu --set full --import-source on --clock-control none -k regex:gemm_tc -s 5 -c 1 -o gpurun_out/prof_gemm_fc1 python tools/bench_gemm.py > gpurun_out/ncu_gemm2.log 2>&1; echo "ncu2 $?"
