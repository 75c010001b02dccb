#!/bin/bash
# tile-level cycle trace of the staged GEMM epilogue (trace build), torch-free
mkdir -p gpurun_out
VTB_LIB=libvtb200_trace.so GEMM_TRACE=1 GEMM_BLOCK=swin3 GEMM_ONLY="fwd,dgrad" timeout 300 python tools/cabi_gemm_bench.py 2>&1 | tee gpurun_out/cabi_gemm_trace_swin3.log
