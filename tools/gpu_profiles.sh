#!/bin/bash
mkdir -p gpurun_out
# 1. launch lists (one full step each)
for wl in vit_b16 swin_s; do
  timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
     --log-file gpurun_out/launches_${wl}.csv python bench.py --workload $wl --warmup 3 --nvtx-step > gpurun_out/ncu_${wl}.log 2>&1
  echo "$wl list exit=$?"
done
# 2. DRAM traffic + duration of every GEMM launch of one ViT-B step
timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
   --clock-control none -k regex:gemm_tc_kernel --csv --log-file gpurun_out/gemm_traffic_vit_b16.csv python bench.py --workload vit_b16 --warmup 3 --nvtx-step > gpurun_out/ncu_gemm_traffic.log 2>&1
echo "traffic exit=$?"
# 3. full capture of the tcgen05 attention kernels and the first GEMMs (source-level)
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"attn_tc|gemm_tc" -c 10 -o gpurun_out/prof_vit_b16_top python bench.py --workload vit_b16 --warmup 3 --nvtx-step > gpurun_out/ncu_full_top.log 2>&1
echo "full exit=$?"
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"attn_tc_bwd" -c 1 -o gpurun_out/prof_attn_tc_bwd python bench.py --workload vit_b16 --warmup 3 --nvtx-step > gpurun_out/ncu_full_bwd.log 2>&1
echo "full bwd exit=$?"
