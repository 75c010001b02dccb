#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-8} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run t_new python -m pytest tests/test_kernels_gpu.py -q -m gpu --no-header -p no:cacheprovider -k "transpose or dwconv or patch"
run t_models python -m pytest tests/test_models_gpu.py -q -m gpu --no-header -p no:cacheprovider
