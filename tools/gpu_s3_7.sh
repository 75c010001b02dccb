#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-4} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
TMO=600 run t_attn python -m pytest tests/test_kernels_gpu.py -q -m gpu --no-header -p no:cacheprovider -x -k "attention"
TAILN=2 run bench_vit python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e
head -4 gpurun_out/breakdown_vit_b16_n1.txt
