#!/bin/bash
# window-attention LDS change: parity + micro-benchmark + Swin-S step, then the whole GPU suite
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-15} gpurun_out/$name.log | cut -c1-${CUT:-300}; }
: > gpurun_out/summary.txt
TMO=300 TAILN=3 run t_win python -m pytest tests/test_kernels_gpu.py -q -m gpu --no-header -p no:cacheprovider -k "window or masked" -x
TMO=300 TAILN=14 run bench_win python tools/bench_winattn.py
TAILN=1 CUT=1500 run bench_swin python bench.py --only --workload swin_s --no-cpu-baseline --no-optimizer-leg --no-e2e
grep -E "attention|layernorm" gpurun_out/breakdown_swin_s_n1.txt
TMO=900 TAILN=6 run t_all python -m pytest tests/ -q -m gpu --no-header -p no:cacheprovider
