#!/bin/bash
# same-box A/B of two library builds over whole steps:  bash tools/gpu_ab_lib.sh <libA.so> <libB.so> [workloads...]
A=$1; B=$2; shift 2
mkdir -p gpurun_out; : > gpurun_out/ab_lib.log
for rep in 1 2; do for wl in ${@:-vit_b16 swin_s}; do for lib in $A $B; do
  echo -n "rep $rep $wl $lib: " >> gpurun_out/ab_lib.log
  VTB_LIB=$lib timeout 300 python bench.py --only --workload $wl --no-cpu-baseline --no-optimizer-leg --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%.3f ms  %.0f img/s  clocks %s  gemm_ms %.2f' % (d['ms_per_step'], d['value'], d['clocks']['sm_mhz'], d['roofline']['gemm_ms_per_step']))" >> gpurun_out/ab_lib.log
done; done; done
cat gpurun_out/ab_lib.log
