#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_models_gpu.py tests/test_kernels_gpu.py -q -m gpu --no-header -p no:cacheprovider -x \
  -k "golden or (gemm_layouts and 128-64-64) or (gemm_layouts and 300) or epilogue or splitk or unaligned or assemble or layernorm_patchify or (layernorm_fwd_bwd and 64-32) or attention_halo or attention_window or (attention_global and 37) or (attention_global and 50) or tcgen05 or masked or cast or patch or pool or transpose or dwconv" > gpurun_out/sanitize.log 2>&1
echo "memcheck exit=$?"; tail -n 15 gpurun_out/sanitize.log
