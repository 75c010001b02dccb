#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-12} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run t_full python -m pytest tests/test_fullsize_gpu.py -q -m gpu --no-header -p no:cacheprovider
TAILN=3 run bench_dino python bench.py --steps 5 --warmup 3 --workload dino_deit_s
