#!/bin/bash
# tcgen05 halo attention: parity tests, micro-benchmark + cross-check against the mma.sync kernels, Halo-T step
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-15} gpurun_out/$name.log | cut -c1-${CUT:-400}; }
: > gpurun_out/summary.txt
TMO=300 TAILN=25 run t_halo python -m pytest tests/test_kernels_gpu.py -q -m gpu --no-header -p no:cacheprovider -k "halo" -x
TMO=300 TAILN=8 run bench_halo python tools/bench_haloattn.py
TMO=300 TAILN=6 run t_models python -m pytest tests/test_models_gpu.py tests/test_fullsize_gpu.py -q -m gpu --no-header -p no:cacheprovider -k "halo"
TAILN=1 CUT=600 run bench_halo_t python bench.py --only --workload halo_t --no-cpu-baseline --no-optimizer-leg --no-e2e
grep -E "attention|layernorm" gpurun_out/breakdown_halo_t_n1.txt
