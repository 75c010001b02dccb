#!/bin/bash
# compute-sanitizer memcheck of the round-2 kernels: tcgen05 halo attention (forward, backward, per-token sum), the
# validation-mode kernels (split3, fp32 attention, fp32 gather, exact SiLU), the persistent global attention kernels
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -q -m gpu --no-header -p no:cacheprovider -x \
  -k "(attention_halo and 1-) or (validation_mode and (vit_tiny or swin_w7 or halo_w7 or pvt_tiny)) or (attention_global and 197) or (golden and halo_w7)" > gpurun_out/sanitize3.log 2>&1
echo "memcheck exit=$?"; tail -n 12 gpurun_out/sanitize3.log | cut -c1-200
