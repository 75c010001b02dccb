#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-400} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-4} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run t_all python -m pytest tests/ -x -q -m gpu --no-header -p no:cacheprovider
TAILN=2 run bench_vit python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e
TAILN=2 run bench_swin python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload swin_s --no-e2e
grep -E "colsum|cast" gpurun_out/breakdown_swin_s_n1.txt gpurun_out/breakdown_vit_b16_n1.txt
