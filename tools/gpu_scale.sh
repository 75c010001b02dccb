#!/bin/bash
# usage: gpu_scale.sh N   (run under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
for wl in vit_b16 swin_s; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 --workload $wl --no-cpu-baseline --no-e2e > gpurun_out/scale_${wl}_n$N.log 2>&1
echo "$wl n=$N exit=$?"; grep '^{' gpurun_out/scale_${wl}_n$N.log | tail -1 | cut -c1-160
done
# DINO DeiT-S multi-crop (BASELINE config 5): eager step at N > 1, then the opt-in graph capture with the deferred centre update
if [ "${DINO:-1}" = 1 ]; then
for flag in "" "--dino-graph"; do
tag=dino${flag:+_graph}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 8 --warmup 3 --workload dino_deit_s --no-cpu-baseline --no-e2e $flag > gpurun_out/scale_${tag}_n$N.log 2>&1
echo "$tag n=$N exit=$?"; grep '^{' gpurun_out/scale_${tag}_n$N.log | tail -1 | cut -c1-160
done
fi
