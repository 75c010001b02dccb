#!/bin/bash
# full GPU suite + ViT-B / Swin-S step with the bias gradients on the CTA-pair weight-gradient launches
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/gputests_colsum.log
GEMM_BLOCK=vitb GEMM_ONLY="wgrad,colsum" timeout 300 python tools/cabi_gemm_bench.py 2>&1 | tail -9
for wl in vit_b16 swin_s; do
timeout 300 python bench.py --workload $wl --only --no-cpu-baseline --no-optimizer-leg --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$wl ms_per_step %.3f  img/s %.0f  clocks %s  gemm_ms %.2f model_frac %s' % (d['ms_per_step'], d['value'], d['clocks']['sm_mhz'], d['roofline']['gemm_ms_per_step'], d['roofline'].get('model_frac')))"
done
