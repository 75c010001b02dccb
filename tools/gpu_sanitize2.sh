#!/bin/bash
mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -q -m gpu --no-header -p no:cacheprovider -x \
  -k "attention_window or masked or wgrad_with_fused or (layernorm_fwd_bwd and 1000-96) or (layernorm_fwd_bwd and 64-32) or layernorm_patchify or (attention_global and 197) or epilogue or (golden and swin)" > gpurun_out/sanitize2.log 2>&1
echo "memcheck exit=$?"; tail -n 12 gpurun_out/sanitize2.log | cut -c1-200
