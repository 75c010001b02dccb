#!/bin/bash
# One gpurun call: kernel parity groups in separate processes, model parity, smoke, short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
ls /root/reference > gpurun_out/ref_ls.txt 2>&1
run() { name=$1; shift; timeout ${TMO:-420} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-6} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run t_gemm   python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm" -x --no-header -p no:cacheprovider
run t_ln     python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "layernorm" --no-header -p no:cacheprovider
run t_attn   python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" --no-header -p no:cacheprovider
run t_misc   python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "cast or patch or pool" --no-header -p no:cacheprovider
run t_models python -m pytest tests/test_models_gpu.py -q -m gpu --no-header -p no:cacheprovider
run smoke    python __graft_entry__.py smoke
if [ "${BENCH:-1}" = "1" ]; then
  TAILN=3 run bench python bench.py --steps ${STEPS:-5} --warmup 3 ${BENCH_ARGS}
fi
