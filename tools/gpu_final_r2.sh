#!/bin/bash
# Round-2 evidence run (1 GPU): tests, smoke, the default bench line (all legs), reference arm, the other BASELINE workloads,
# ncu launch lists, per-launch GEMM DRAM traffic, attention tensor-pipe / DRAM metrics, --set full captures of the attention
# kernels.  Everything lands in gpurun_out/; the summaries are copied to profiles/r02_* by hand.
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-300; }
: > gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader | tee -a gpurun_out/summary.txt
TMO=900 run t_all env VTB_TEST_INPUT_V2=1 python -m pytest tests/ -q -m gpu --no-header -p no:cacheprovider
run smoke python -c "import __graft_entry__ as g; g.smoke()"
TAILN=1 run bench_default_n1 python bench.py
TAILN=1 run bench_reference_arm python bench.py --impl reference --steps 3 --warmup 1
TAILN=1 run bench_pvt_small_n1 python bench.py --only --workload pvt_small --steps 60
TAILN=1 run bench_halo_t_n1 python bench.py --only --workload halo_t --steps 40
for wl in vit_b16 swin_s pvt_small; do
  timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
     --log-file gpurun_out/launches_${wl}.csv python bench.py --only --workload $wl --warmup 3 --nvtx-step > gpurun_out/ncu_${wl}.log 2>&1
  echo "launch list $wl exit=$?"
  python tools/ncu_agg.py gpurun_out/launches_${wl}.csv 30 > gpurun_out/agg_${wl}.txt 2>&1
done
for wl in vit_b16 swin_s; do
  timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
     --clock-control none -k regex:gemm_tc_kernel --csv --log-file gpurun_out/gemm_traffic_$wl.csv python bench.py --only --workload $wl --warmup 3 --nvtx-step > gpurun_out/ncu_gemm_traffic_$wl.log 2>&1
  echo "traffic $wl exit=$?"
  python tools/ncu_gemm_traffic.py gpurun_out/gemm_traffic_$wl.csv gpurun_out/gemm_traffic_$wl.json "$wl (B=256) one fwd+bwd step" > /dev/null
done
# every attention kernel of every workload: duration, DRAM bytes, tensor-pipe activity per launch
for wl in vit_b16 swin_s pvt_small halo_t; do
  timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
     --clock-control none -k regex:"attn_" --csv --log-file gpurun_out/attn_metrics_$wl.csv python bench.py --only --workload $wl --warmup 3 --nvtx-step > gpurun_out/ncu_attn_$wl.log 2>&1
  echo "attention metrics $wl exit=$?"
done
python tools/ncu_attn_pipe.py gpurun_out > gpurun_out/attn_tensor_pipe.json; cat gpurun_out/attn_tensor_pipe.json
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"attn_tc_fwd2|attn_tc_bwd2" -c 2 -f -o gpurun_out/prof_vit_attn python bench.py --only --workload vit_b16 --warmup 3 --nvtx-step > gpurun_out/ncu_full_vit.log 2>&1
echo "full vit attention exit=$?"
ncu -i gpurun_out/prof_vit_attn.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_key_metrics.py > gpurun_out/ncu_full_vit_attn.txt
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"attn_wt" -s 20 -c 2 -f -o gpurun_out/prof_swin_wt python bench.py --only --workload swin_s --warmup 3 --nvtx-step > gpurun_out/ncu_full_swin.log 2>&1
echo "full swin window exit=$?"
ncu -i gpurun_out/prof_swin_wt.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_key_metrics.py > gpurun_out/ncu_full_swin_wt.txt
rm -f gpurun_out/*.ncu-rep
