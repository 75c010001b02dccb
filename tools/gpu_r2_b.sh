#!/bin/bash
# kernel + model GPU tests, then the ViT-B bench line only (fast check of a kernel change inside the whole step)
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-8} gpurun_out/$name.log | cut -c1-${CUT:-300}; }
: > gpurun_out/summary.txt
TMO=900 TAILN=6 run t_gpu python -m pytest tests/ -q -m gpu --no-header -p no:cacheprovider -x
TMO=600 TAILN=1 CUT=3000 run bench_vit python bench.py --only --no-cpu-baseline --no-optimizer-leg ${BENCH_ARGS}
cat gpurun_out/breakdown_vit_b16_n1.txt | head -24
