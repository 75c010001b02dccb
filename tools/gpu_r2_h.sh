#!/bin/bash
# a_colsum reader on warp 2: parity, C-ABI microbench (pair / 1-CTA), step A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "wgrad or gemm" 2>&1 | tail -3
echo "== pair tiles"; GEMM_BLOCK=vitb,swin3,swin1 GEMM_ONLY="wgrad,colsum" timeout 300 python tools/cabi_gemm_bench.py 2>&1 | tee gpurun_out/cabi_gemm_colsum_pair_w2.log
echo "== 1-CTA tiles"; VTB_OPTS=gemm_colsum_pair=0 GEMM_BLOCK=vitb,swin3,swin1 GEMM_ONLY="+ colsum" timeout 300 python tools/cabi_gemm_bench.py 2>&1 | tee gpurun_out/cabi_gemm_colsum_1cta_w2.log
: > gpurun_out/ab_colsum2.log
one() {  # workload, env...
  wl=$1; shift
  echo "=== $wl $*" >> gpurun_out/ab_colsum2.log
  env "$@" timeout 300 python bench.py --workload $wl --only --no-cpu-baseline --no-optimizer-leg --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms_per_step %.3f  img/s %.0f  clocks %s  gemm_ms %.2f' % (d['ms_per_step'], d['value'], d['clocks']['sm_mhz'], d['roofline']['gemm_ms_per_step']))" >> gpurun_out/ab_colsum2.log
}
one vit_b16 VTB_WGRAD_COLSUM=narrow
one vit_b16 VTB_WGRAD_COLSUM=always
one swin_s VTB_WGRAD_COLSUM=narrow
one swin_s VTB_WGRAD_COLSUM=always
one swin_s VTB_WGRAD_COLSUM=never
cat gpurun_out/ab_colsum2.log
