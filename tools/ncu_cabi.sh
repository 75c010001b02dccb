#!/bin/bash
# One `--set full` capture of a few launches of one kernel, driven by a torch-free harness (starts in a second, so the
# whole call costs well under a GPU-minute), condensed by tools/ncu_key_metrics.py.
#   bash tools/ncu_cabi.sh <tag> <kernel regex> <launch count> <harness.py> [harness args...]     (env vars pass through)
# e.g. GEMM_BLOCK=swin3 GEMM_ONLY="fc2  fwd" bash tools/ncu_cabi.sh swin3_fc2 gemm_tc_kernel 2 tools/cabi_gemm_bench.py
tag=$1; regex=$2; count=$3; shift 3
mkdir -p gpurun_out
timeout ${TMO:-180} ncu --set full --clock-control none --import-source on -k regex:$regex -c $count -f \
  -o gpurun_out/ncu_$tag python "$@" > gpurun_out/ncu_$tag.log 2>&1
echo "ncu exit=$?"; tail -3 gpurun_out/ncu_$tag.log
ncu -i gpurun_out/ncu_$tag.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_key_metrics.py > gpurun_out/ncu_$tag.txt
cat gpurun_out/ncu_$tag.txt
