#!/bin/bash
# Device input path: parity tests, micro-benchmark, the opt-in bench leg, one ncu capture of the kernel.
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-8} gpurun_out/$name.log | cut -c1-${CUT:-400}; }
: > gpurun_out/summary.txt
python tools/input_selftest.py --prepare > /dev/null 2>&1 || true   # needs tests/golden only; writes tools/_selftest_input.npz
TAILN=20 run selftest_v1 python tools/input_selftest.py
TAILN=20 run selftest_v2 python tools/input_selftest.py --variant2
TAILN=20 run selftest_v3 python tools/input_selftest.py --variant3
TAILN=15 VTB_TEST_INPUT_V2=1 run t_input env VTB_TEST_INPUT_V2=1 python -m pytest tests/test_input_path.py -x -q -m gpu --no-header -p no:cacheprovider
run bench_input python tools/bench_input.py
TAILN=1 CUT=4000 run bench_e2e_u8 python bench.py --workload swin_s --e2e-u8 --no-cpu-baseline --no-optimizer-leg --steps 10
TMO=600 run ncu_input ncu --set full --clock-control none --import-source on -k regex:input_batch_kernel -c 3 \
  -f -o gpurun_out/r02_input_kernel python tools/input_selftest.py --bench-only --variant2
ncu -i gpurun_out/r02_input_kernel.ncu-rep --page raw --csv > gpurun_out/r02_input_kernel_raw.csv 2>/dev/null || true
