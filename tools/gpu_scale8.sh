#!/bin/bash
# the driver's launch line at N GPUs (default bench: ViT-B + no_graph + Swin-S + DINO legs), then the reference arm at N
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N > gpurun_out/bench_default_n$N.log 2>&1
echo "default n=$N exit=$?"; grep '^{' gpurun_out/bench_default_n$N.log | tail -1 | cut -c1-3000
