#!/bin/bash
# split-K factor / CTA-pair sweep of the ViT-B weight gradients (auto: fc1 / fc2 1-CTA tiles x 2 splits, QKV pairs x 8 splits)
mkdir -p gpurun_out; : > gpurun_out/cabi_gemm_wgrad_splits.log
for cl in 1 2; do for sp in 0 2 3 4 5 6 8; do
  echo "== gemm_cluster=$cl splits=$sp" | tee -a gpurun_out/cabi_gemm_wgrad_splits.log
  GEMM_NOCHECK=1 GEMM_OPTS=gemm_cluster=$cl GEMM_SPLITS=$sp GEMM_BLOCK=vitb GEMM_ONLY="wgrad" timeout 60 python tools/cabi_gemm_bench.py 2>&1 | grep "vit-b" | grep -v total | cut -c1-60 | tee -a gpurun_out/cabi_gemm_wgrad_splits.log
done; done
