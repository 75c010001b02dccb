#!/bin/bash
# validation mode: golden models + full-size models at rtol 1e-3
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-15} gpurun_out/$name.log | cut -c1-${CUT:-300}; }
: > gpurun_out/summary.txt
TMO=600 TAILN=40 run t_valid python -m pytest tests/test_models_gpu.py tests/test_fullsize_gpu.py -q -m gpu --no-header -p no:cacheprovider -k "validation" -s
