#!/bin/bash
# Step-side multi-tensor kernels: parity tests, then the bench lines with the weight arena + optimizer leg.
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-${CUT:-400}; }
TAILN=25 run t_step python -m pytest tests/test_step_ops.py -x -q -m gpu --no-header -p no:cacheprovider
for wl in ${WLS:-vit_b16 swin_s pvt_small}; do
  TAILN=1 CUT=3000 run bench_${wl}_arena python bench.py --workload $wl --no-cpu-baseline --no-e2e --steps 10
done
