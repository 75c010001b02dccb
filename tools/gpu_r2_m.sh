#!/bin/bash
# operand-traffic knock-outs on the ViT-B GEMMs (debug build): is the mainloop bound by the tensor pipe or by L2 -> SM loads?
mkdir -p gpurun_out; : > gpurun_out/cabi_gemm_dbg_vitb.log
for d in 0 32 64 96 28 60 124; do
  echo "== gemm_dbg=$d" | tee -a gpurun_out/cabi_gemm_dbg_vitb.log
  GEMM_NOCHECK=1 VTB_LIB=libvtb200_dbg.so GEMM_OPTS=gemm_dbg=$d GEMM_BLOCK=vitb GEMM_ONLY="fwd,dgrad bf16" timeout 120 python tools/cabi_gemm_bench.py 2>&1 | grep "vit-b" | grep -v total | tee -a gpurun_out/cabi_gemm_dbg_vitb.log
done
