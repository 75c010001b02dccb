#!/bin/bash
# First GPU call of round 2: everything written in round 1 after the GPU budget ran out, cheapest first.
#   1. torch-free C-ABI harnesses (seconds each): input kernel variants 1 / 2, GEMM and attention self-checks + timings
#   2. the whole `-m gpu` suite, with the variant-2 input test enabled
#   3. default bench line + the opt-in uint8 input leg
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-8} gpurun_out/$name.log | cut -c1-${CUT:-300}; }
: > gpurun_out/summary.txt
[ -f tools/_selftest_input.npz ] || python tools/input_selftest.py --prepare > gpurun_out/prepare.log 2>&1
TMO=60 TAILN=5 run selftest_v1 python tools/input_selftest.py
TMO=60 TAILN=5 run selftest_v2 python tools/input_selftest.py --variant2
TMO=60 TAILN=5 run selftest_v3 python tools/input_selftest.py --variant3
TMO=120 TAILN=26 run cabi_gemm python tools/cabi_gemm_bench.py
TMO=120 TAILN=16 run cabi_attn python tools/cabi_attn_bench.py
TMO=120 TAILN=12 run cabi_step python tools/cabi_step_bench.py
TMO=120 TAILN=10 run cabi_ln python tools/cabi_ln_bench.py
TMO=1500 TAILN=12 run t_all env VTB_TEST_INPUT_V2=1 python -m pytest tests/ -x -q -m gpu --no-header -p no:cacheprovider
TMO=600 TAILN=1 CUT=5000 run bench_default python bench.py --e2e-u8 --eager-baseline
