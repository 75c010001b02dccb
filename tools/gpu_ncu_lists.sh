#!/bin/bash
mkdir -p gpurun_out
for wl in ${WLS:-vit_b16 swin_s}; do
  timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
     --log-file gpurun_out/launches_${wl}.csv python bench.py --workload $wl --warmup 3 --nvtx-step > gpurun_out/ncu_${wl}.log 2>&1
  echo "$wl exit=$?"; grep -c gpu__time_duration gpurun_out/launches_${wl}.csv
done
