#!/bin/bash
# Round-end evidence run (1 GPU): tests, smoke, bench lines for every BASELINE workload, reference arm, ncu launch lists,
# per-launch GEMM DRAM traffic, --set full captures of the top kernels.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-300; }
: > gpurun_out/summary.txt
run t_all python -m pytest tests/ -x -q -m gpu --no-header -p no:cacheprovider
run smoke python -c "import __graft_entry__ as g; g.smoke()"
TAILN=1 run bench_vit_b16_n1 python bench.py
TAILN=1 run bench_reference_arm python bench.py --impl reference --steps 3 --warmup 1
TAILN=1 run bench_swin_s_n1 python bench.py --workload swin_s --no-cpu-baseline
TAILN=1 run bench_pvt_small_n1 python bench.py --workload pvt_small --no-cpu-baseline --no-e2e --steps 10
TAILN=1 run bench_halo_t_n1 python bench.py --workload halo_t --no-cpu-baseline --no-e2e --steps 10
TAILN=1 run bench_dino_n1 python bench.py --workload dino_deit_s --no-cpu-baseline --no-e2e --steps 5 --warmup 3
WLS="vit_b16 swin_s" bash tools/gpu_ncu_lists.sh
python tools/ncu_agg.py gpurun_out/launches_vit_b16.csv > gpurun_out/agg_vit.txt 2>&1
python tools/ncu_agg.py gpurun_out/launches_swin_s.csv > gpurun_out/agg_swin.txt 2>&1
for wl in vit_b16 swin_s; do
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
   --clock-control none -k regex:gemm_tc_kernel --csv --log-file gpurun_out/gemm_traffic_$wl.csv python bench.py --workload $wl --warmup 3 --nvtx-step > gpurun_out/ncu_gemm_traffic_$wl.log 2>&1
echo "traffic $wl exit=$?"
python tools/ncu_gemm_traffic.py gpurun_out/gemm_traffic_$wl.csv gpurun_out/gemm_traffic_$wl.json "$wl (B=256) one fwd+bwd step" | head -12
done
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"attn_tc|gemm_tc" -c 10 -o gpurun_out/prof_vit_b16_top python bench.py --workload vit_b16 --warmup 3 --nvtx-step > gpurun_out/ncu_full_vit.log 2>&1
echo "full vit exit=$?"
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"attn_wt" -s 20 -c 4 -o gpurun_out/prof_swin_wt python bench.py --workload swin_s --warmup 3 --nvtx-step > gpurun_out/ncu_full_swin.log 2>&1
echo "full swin exit=$?"
