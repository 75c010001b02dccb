#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-12} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
TMO=300 run t_gemm python -m pytest tests/test_kernels_gpu.py -q -m gpu --no-header -p no:cacheprovider -k "gemm" -x
TMO=300 run t_models python -m pytest tests/test_models_gpu.py tests/test_fullsize_gpu.py -q -m gpu --no-header -p no:cacheprovider -x
TAILN=2 run bench_vit python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e
cp gpurun_out/breakdown_vit_b16_n1.txt gpurun_out/breakdown_vit_b16_pair.txt
TAILN=2 VTB_GEMM_CLUSTER=0 run bench_vit_1cta python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e
TAILN=2 run bench_swin python bench.py --steps 8 --warmup 3 --no-cpu-baseline --workload swin_s --no-e2e
grep gemm gpurun_out/breakdown_vit_b16_pair.txt | head -14
grep gemm gpurun_out/breakdown_vit_b16_n1.txt | head -14
