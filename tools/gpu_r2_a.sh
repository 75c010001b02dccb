#!/bin/bash
# Round 2, call A: the whole -m gpu suite WITHOUT -x (every file must run), then the default bench line
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-8} gpurun_out/$name.log | cut -c1-${CUT:-300}; }
: > gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader | tee -a gpurun_out/summary.txt
nproc | tee -a gpurun_out/summary.txt
TMO=1200 TAILN=30 run t_all env VTB_TEST_INPUT_V2=1 python -m pytest tests/ -q -m gpu --no-header -p no:cacheprovider
TMO=900 TAILN=1 CUT=12000 run bench_default python bench.py

