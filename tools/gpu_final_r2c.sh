#!/bin/bash
# Round-2 closing evidence after the last kernel changes (bias gradients on the CTA-pair weight gradients, element dropout):
# the whole GPU suite, smoke, the default bench line (all legs), the ViT-B launch list.  Outputs -> gpurun_out/, copied to profiles/r02_*.
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-400; }
: > gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader | tee -a gpurun_out/summary.txt
TMO=600 run t_all python -m pytest tests/ -q -m gpu --no-header -p no:cacheprovider
TMO=300 run smoke python -c "import __graft_entry__ as g; g.smoke()"
TMO=900 TAILN=1 run bench_default_n1 python bench.py
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches_vit_b16.csv python bench.py --only --workload vit_b16 --warmup 3 --nvtx-step > gpurun_out/ncu_vit_b16.log 2>&1
echo "launch list vit_b16 exit=$?" | tee -a gpurun_out/summary.txt
python tools/ncu_agg.py gpurun_out/launches_vit_b16.csv 30 > gpurun_out/agg_vit_b16.txt 2>&1; head -12 gpurun_out/agg_vit_b16.txt
