#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-5} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
TMO=120 run t_gemm python -m pytest tests/test_kernels_gpu.py -q -m gpu --no-header -p no:cacheprovider -x -k "gemm"
TMO=200 run t_models python -m pytest tests/test_models_gpu.py tests/test_fullsize_gpu.py -q -m gpu --no-header -p no:cacheprovider -x
TAILN=2 run bench_vit python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e
TAILN=2 run bench_swin python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload swin_s --no-e2e
grep -E "colsum|wgrad" gpurun_out/breakdown_vit_b16_n1.txt | head -8
