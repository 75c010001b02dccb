#!/bin/bash
mkdir -p gpurun_out
for red in ddp flat; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e --reducer $red > gpurun_out/bench_n2_$red.log 2>&1
echo "n2 $red exit=$?"; tail -n 2 gpurun_out/bench_n2_$red.log | cut -c1-400
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 --workload swin_s --no-e2e > gpurun_out/bench_n2_swin.log 2>&1
echo "n2 swin exit=$?"; tail -n 2 gpurun_out/bench_n2_swin.log | cut -c1-300
