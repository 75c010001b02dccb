#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py tests/test_fullsize_gpu.py -q -m gpu --no-header -p no:cacheprovider -x -k "halo or attention" 2>&1 | tail -3
timeout 200 python bench.py --workload halo_t --no-cpu-baseline --no-e2e --steps 10 > gpurun_out/bench_halo_t_n1.log 2>&1; grep '^{' gpurun_out/bench_halo_t_n1.log | cut -c1-200; head -4 gpurun_out/breakdown_halo_t_n1.txt
