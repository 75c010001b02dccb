#!/bin/bash
# Step-side multi-tensor kernels + DINO head rows: the whole GPU suite, then bench lines with the weight arena / optimizer leg.
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-${CUT:-400}; }
TAILN=12 run t_all python -m pytest tests/ -x -q -m gpu --no-header -p no:cacheprovider
for wl in ${WLS:-dino_deit_s halo_t vit_b16}; do
  TAILN=1 CUT=3000 run bench_${wl}_arena python bench.py --workload $wl --no-cpu-baseline --no-e2e --steps ${STEPS:-8}
done
