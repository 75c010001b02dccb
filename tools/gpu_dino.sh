#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py -q -m gpu --no-header -p no:cacheprovider -x -k "dino" 2>&1 | tail -4
timeout 300 python bench.py --workload dino_deit_s --no-cpu-baseline --no-e2e --steps 5 --warmup 3 > gpurun_out/bench_dino_n1.log 2>&1; grep '^{' gpurun_out/bench_dino_n1.log | cut -c1-200; grep -i "error\|capture" gpurun_out/bench_dino_n1.log | head -3
