#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-${CUT:-300}; }
: > gpurun_out/summary.txt
TMO=900 TAILN=3 run t_all python -m pytest tests/ -q -m gpu --no-header -p no:cacheprovider
for wl in vit_b16 swin_s; do
  timeout 300 python bench.py --only --workload $wl --no-cpu-baseline --no-optimizer-leg --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$wl: %.3f ms  %.0f img/s  clocks %s  gemm_ms %.2f' % (d['ms_per_step'], d['value'], d['clocks']['sm_mhz'], d['roofline']['gemm_ms_per_step']))"
done
