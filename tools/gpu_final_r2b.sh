#!/bin/bash
# Round-2 closing evidence (1 GPU): the whole GPU suite, smoke, default bench line (all legs), reference arm, the other
# BASELINE workloads, launch lists (Halo-T* with the tcgen05 halo kernels), attention metrics per workload, --set full summary
# of the halo kernels.  Everything lands in gpurun_out/; summaries are copied to profiles/r02_* afterwards.
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-300; }
: > gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader | tee -a gpurun_out/summary.txt
TMO=900 run t_all python -m pytest tests/ -q -m gpu --no-header -p no:cacheprovider
run smoke python -c "import __graft_entry__ as g; g.smoke()"
TAILN=1 run bench_default_n1 python bench.py
TAILN=1 run bench_reference_arm python bench.py --impl reference --steps 3 --warmup 1
TAILN=1 run bench_pvt_small_n1 python bench.py --only --workload pvt_small --steps 60
TAILN=1 run bench_halo_t_n1 python bench.py --only --workload halo_t --steps 40
for wl in halo_t; do
  timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
     --log-file gpurun_out/launches_${wl}.csv python bench.py --only --workload $wl --warmup 3 --nvtx-step > gpurun_out/ncu_${wl}.log 2>&1
  echo "launch list $wl exit=$?"
  python tools/ncu_agg.py gpurun_out/launches_${wl}.csv 30 > gpurun_out/agg_${wl}.txt 2>&1
done
for wl in halo_t swin_s; do
  timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
     --clock-control none -k regex:"attn_" --csv --log-file gpurun_out/attn_metrics_$wl.csv python bench.py --only --workload $wl --warmup 3 --nvtx-step > gpurun_out/ncu_attn_$wl.log 2>&1
  echo "attention metrics $wl exit=$?"
done
python tools/ncu_attn_pipe.py gpurun_out > gpurun_out/attn_tensor_pipe.json; cat gpurun_out/attn_tensor_pipe.json | head -c 1500
HT_ONLY=56 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"attn_ht_(fwd|bwd|dkv)" -s 20 -c 3 -f -o gpurun_out/prof_ht python tools/bench_haloattn.py > gpurun_out/ncu_ht.log 2>&1
echo "full halo exit=$?"
ncu -i gpurun_out/prof_ht.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_key_metrics.py > gpurun_out/ncu_full_halo.txt
rm -f gpurun_out/*.ncu-rep
