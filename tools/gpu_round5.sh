#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-6} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run t_gemm python -m pytest tests/test_kernels_gpu.py -q -m gpu --no-header -p no:cacheprovider -k "gemm"
run t_models python -m pytest tests/test_models_gpu.py -q -m gpu --no-header -p no:cacheprovider
TAILN=2 run bench_vit python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e
TAILN=2 run bench_swin python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload swin_s --no-e2e
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:gemm_tc_kernel -c 8 -o gpurun_out/prof_gemm_vit python bench.py --warmup 3 --nvtx-step > gpurun_out/ncu_full_gemm.log 2>&1
echo "ncu full exit=$?"
