#!/bin/bash
# same-box A/B of the whole ViT-B step: VTB_OPTS variants back to back, twice (the pool's boxes differ by +-2 %, and the
# step runs at the power cap, so only same-box numbers compare)
mkdir -p gpurun_out; : > gpurun_out/ab_bench.log
for rep in 1 2; do
  for opts in "${@}"; do
    echo "=== rep $rep VTB_OPTS=$opts" >> gpurun_out/ab_bench.log
    VTB_OPTS=$opts timeout 300 python bench.py --only --no-cpu-baseline --no-optimizer-leg --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
a=d['roofline'].get('attention',{})
print('ms_per_step %.3f  img/s %.0f  clocks %s  gemm_ms %.2f  attn fwd %.2f bwd %.2f' % (d['ms_per_step'], d['value'], d['clocks']['sm_mhz'], d['roofline']['gemm_ms_per_step'], a.get('attention_fwd',{}).get('ms_per_step',0), a.get('attention_bwd',{}).get('ms_per_step',0)))" >> gpurun_out/ab_bench.log
  done
done
cat gpurun_out/ab_bench.log
