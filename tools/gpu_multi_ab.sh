#!/bin/bash
# all-reduce variants of the ViT-B step at N ranks, same box:  bash tools/gpu_multi_ab.sh <N>
N=${1:-2}
mkdir -p gpurun_out; : > gpurun_out/multi_ab_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
run() { tag=$1; shift; line=$(timeout 300 env "$@" 2>/dev/null | tail -n 1); echo "$tag $(echo "$line" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.0f img/s %.2f ms' % (d['value'], d['ms_per_step']))")" | tee -a gpurun_out/multi_ab_n$N.log; }
B="bench.py --gpus $N --only --no-optimizer-leg --no-e2e --no-cpu-baseline"
for rep in 1 2; do
run "no-overlap-128" VTB_BUCKET_MB=128 $TR $B --no-overlap
run "overlap-32    " VTB_BUCKET_MB=32 $TR $B
run "overlap-128   " VTB_BUCKET_MB=128 $TR $B
run "overlap-64    " VTB_BUCKET_MB=64 $TR $B
run "ddp-nograph   " VTB_BUCKET_MB=32 $TR $B --no-graph
done
