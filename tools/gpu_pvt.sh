#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py tests/test_fullsize_gpu.py -q -m gpu --no-header -p no:cacheprovider -x -k "patch or pvt or twins or halo or swin" 2>&1 | tail -4
timeout 200 python bench.py --workload pvt_small --no-cpu-baseline --no-e2e --steps 10 > gpurun_out/bench_pvt_small_n1.log 2>&1; grep '^{' gpurun_out/bench_pvt_small_n1.log | cut -c1-200; grep -E "patch_|cast" gpurun_out/breakdown_pvt_small_n1.txt
