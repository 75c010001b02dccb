#!/bin/bash
# attention harness A/B: old and new tcgen05 global-attention kernels through the C-ABI (torch-free, seconds each)
mkdir -p gpurun_out
for opts in ${ATTN_CONFIGS:-"attn_tc_fwd_version=1,attn_tc_bwd_version=1" "attn_tc_fwd_version=2,attn_tc_bwd_version=1"}; do
  echo "=== VTB_OPTS=$opts" | tee -a gpurun_out/attn_ab.log
  VTB_OPTS=$opts ATTN_ONLY=global timeout 180 python tools/cabi_attn_bench.py 2>&1 | tee -a gpurun_out/attn_ab.log | tail -n 20
done
