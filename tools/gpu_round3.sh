#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-6} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run t_gemm   python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm or assemble" --no-header -p no:cacheprovider
run t_models python -m pytest tests/test_models_gpu.py -q -m gpu --no-header -p no:cacheprovider
TAILN=2 run bench_vit python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e
TAILN=2 run bench_swin python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload swin_s --no-e2e
for wl in vit_b16 swin_s; do
  timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
     --log-file gpurun_out/launches_${wl}.csv python bench.py --workload $wl --warmup 3 --nvtx-step > gpurun_out/ncu_${wl}.log 2>&1
  echo "$wl exit=$?"; grep -c gpu__time_duration gpurun_out/launches_${wl}.csv
done
