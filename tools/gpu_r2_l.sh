#!/bin/bash
# knock-out probes of the staged GEMM epilogue (debug build, -DVTB_GEMM_DBG): which part of a sub-tile step costs the time
mkdir -p gpurun_out; : > gpurun_out/cabi_gemm_dbg_swin3.log
for d in 0 1 2 4 8 16 24 28 6; do
  echo "== gemm_dbg=$d" | tee -a gpurun_out/cabi_gemm_dbg_swin3.log
  GEMM_NOCHECK=1 VTB_LIB=libvtb200_dbg.so GEMM_OPTS=gemm_dbg=$d GEMM_BLOCK=swin3 GEMM_ONLY="${ONLY:-fwd}" timeout 120 python tools/cabi_gemm_bench.py 2>&1 | grep "swin-s3" | grep -v total | tee -a gpurun_out/cabi_gemm_dbg_swin3.log
done
