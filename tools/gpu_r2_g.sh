#!/bin/bash
# same-box A/B of the whole step: bias gradients on the weight-gradient launches (VTB_WGRAD_COLSUM) / CTA-pair a_colsum
mkdir -p gpurun_out; : > gpurun_out/ab_colsum.log
one() {  # workload, env...
  wl=$1; shift
  echo "=== $wl $*" >> gpurun_out/ab_colsum.log
  env "$@" timeout 300 python bench.py --workload $wl --only --no-cpu-baseline --no-optimizer-leg --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms_per_step %.3f  img/s %.0f  clocks %s  gemm_ms %.2f' % (d['ms_per_step'], d['value'], d['clocks']['sm_mhz'], d['roofline']['gemm_ms_per_step']))" >> gpurun_out/ab_colsum.log
}
for rep in 1 2; do
  one vit_b16 VTB_WGRAD_COLSUM=narrow
  one vit_b16 VTB_WGRAD_COLSUM=always
  one vit_b16 VTB_WGRAD_COLSUM=always VTB_OPTS=gemm_colsum_pair=0
done
one swin_s VTB_WGRAD_COLSUM=narrow VTB_OPTS=gemm_colsum_pair=0
one swin_s VTB_WGRAD_COLSUM=always
one swin_s VTB_WGRAD_COLSUM=narrow VTB_OPTS=gemm_colsum_pair=0
one swin_s VTB_WGRAD_COLSUM=always
cat gpurun_out/ab_colsum.log
