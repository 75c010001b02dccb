#!/bin/bash
# gradient buckets in NCCL-registered memory (VTB_NCCL_POOL) vs plain device memory, ViT-B step at N ranks, same box
N=${1:-2}
mkdir -p gpurun_out; : > gpurun_out/multi_pool_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
run() { tag=$1; shift; timeout 240 env "$@" > gpurun_out/_mp.out 2> gpurun_out/_mp.err; rc=$?; line=$(tail -n 1 gpurun_out/_mp.out); echo "$tag rc=$rc $(echo "$line" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.0f img/s %.2f ms registered=%s' % (d['value'], d['ms_per_step'], d['config'].get('nccl_registered_buckets')))" 2>&1 | tail -1)" | tee -a gpurun_out/multi_pool_n$N.log; grep -i "warn\|error\|NVLS" gpurun_out/_mp.err | sort | uniq -c | sort -rn | head -5 | tee -a gpurun_out/multi_pool_n$N.log; }
B="bench.py --gpus $N --only --no-optimizer-leg --no-e2e --no-cpu-baseline"
for rep in 1 ${REPS:-2}; do
run "pool=1" VTB_NCCL_POOL=1 $TR $B
run "pool=0" VTB_NCCL_POOL=0 $TR $B
done
