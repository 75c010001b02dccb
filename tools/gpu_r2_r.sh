#!/bin/bash
# tile-N choice: N = 384 as 2 x 256 (a third of the second tile's MMA columns wasted, operand re-reads 2x instead of 3x) vs 3 x 128
mkdir -p gpurun_out; : > gpurun_out/ab_bn_waste.log
for w in 120 140; do
  echo "== gemm_bn_waste_pct=$w" | tee -a gpurun_out/ab_bn_waste.log
  GEMM_OPTS=gemm_bn_waste_pct=$w GEMM_BLOCK=swin3 timeout 60 python tools/cabi_gemm_bench.py 2>&1 | grep "swin-s3\|self-check" | grep -v colsum | cut -c1-75 | tee -a gpurun_out/ab_bn_waste.log
done
for w in 120 140 120 140; do
  echo "=== swin_s gemm_bn_waste_pct=$w" >> gpurun_out/ab_bn_waste.log
  VTB_OPTS=gemm_bn_waste_pct=$w timeout 200 python bench.py --workload swin_s --only --no-cpu-baseline --no-optimizer-leg --no-e2e --steps 12 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms_per_step %.3f  img/s %.0f  clocks %s  gemm_ms %.2f' % (d['ms_per_step'], d['value'], d['clocks']['sm_mhz'], d['roofline']['gemm_ms_per_step']))" >> gpurun_out/ab_bn_waste.log
done
tail -8 gpurun_out/ab_bn_waste.log
