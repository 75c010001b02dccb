#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-15} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
TMO=300 run t_win python -m pytest tests/test_kernels_gpu.py -q -m gpu --no-header -p no:cacheprovider -k "window or masked" -x
TMO=300 TAILN=12 run bench_win python tools/bench_winattn.py
