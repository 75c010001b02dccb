#!/bin/bash
# timings of the Swin stage-3 GEMM cases (torch-free harness), then one --set full launch of fc1 fwd and fc2 fwd each
mkdir -p gpurun_out
GEMM_BLOCK=swin3 timeout 120 python tools/cabi_gemm_bench.py 2>&1 | grep "swin-s3" | tee gpurun_out/cabi_gemm_swin3.log
for c in "fc1  fwd" "fc2  fwd"; do
  tag=$(echo $c | tr -d ' ')
  GEMM_BLOCK=swin3 GEMM_ONLY="$c" timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 8 -c 1 -f \
    -o gpurun_out/ncu_gemm_$tag python tools/cabi_gemm_bench.py > gpurun_out/ncu_gemm_$tag.log 2>&1
  echo "ncu $c exit=$?"
  ncu -i gpurun_out/ncu_gemm_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_gemm_${tag}_raw.csv 2>/dev/null
  python tools/ncu_key_metrics.py < gpurun_out/ncu_gemm_${tag}_raw.csv | head -40
done
