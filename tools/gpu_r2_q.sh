#!/bin/bash
# same-box A/B of the whole step: CTA pairs also where pairs x splits < pair slots (gemm_cluster=2) vs the round-1 rule (1)
mkdir -p gpurun_out; : > gpurun_out/ab_pairs_rule.log
one() {  # workload, opts
  echo "=== $1 VTB_OPTS=$2" >> gpurun_out/ab_pairs_rule.log
  VTB_OPTS=$2 timeout 200 python bench.py --workload $1 --only --no-cpu-baseline --no-optimizer-leg --no-e2e --steps 12 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms_per_step %.3f  img/s %.0f  clocks %s  gemm_ms %.2f' % (d['ms_per_step'], d['value'], d['clocks']['sm_mhz'], d['roofline']['gemm_ms_per_step']))" >> gpurun_out/ab_pairs_rule.log
}
for wl in vit_b16 swin_s pvt_small halo_t; do
  one $wl gemm_cluster=1
  one $wl gemm_cluster=2
done
cat gpurun_out/ab_pairs_rule.log
