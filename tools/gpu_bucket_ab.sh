#!/bin/bash
# bucket size of the overlapped flat reducer at N GPUs (ViT-B step)
N=${1:-8}
mkdir -p gpurun_out; : > gpurun_out/bucket_ab_n$N.log
for mb in 128 32 16; do
  VTB_BUCKET_MB=$mb timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --only --workload vit_b16 --no-cpu-baseline --no-e2e --no-optimizer-leg --steps 20 --warmup 5 2>/dev/null | grep '^{' | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bucket_mb $mb: %.3f ms  %.0f img/s  clocks %s' % (d['ms_per_step'], d['value'], d['clocks']['sm_mhz']))" | tee -a gpurun_out/bucket_ab_n$N.log
done
