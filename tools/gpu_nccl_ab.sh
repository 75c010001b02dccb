#!/bin/bash
# NCCL knobs for the overlapped flat reducer at N GPUs (ViT-B step): CTA budget of the collectives, no overlap
N=${1:-8}
mkdir -p gpurun_out; : > gpurun_out/nccl_ab_n$N.log
runone() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --only --workload vit_b16 --no-cpu-baseline --no-e2e --no-optimizer-leg --steps 20 --warmup 5 $EXTRA 2>/dev/null | grep '^{' | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$tag: %.3f ms  %.0f img/s  clocks %s' % (d['ms_per_step'], d['value'], d['clocks']['sm_mhz']))" | tee -a gpurun_out/nccl_ab_n$N.log; }
runone "max_ctas 4" NCCL_MAX_CTAS=4
runone "max_ctas 16" NCCL_MAX_CTAS=16
runone "default" X=1
EXTRA=--no-overlap runone "no overlap" X=1
