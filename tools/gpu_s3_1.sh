#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-8} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run t_all python -m pytest tests/ -x -q -m gpu --no-header -p no:cacheprovider
run smoke python -c "import __graft_entry__ as g; g.smoke()"
TAILN=2 run bench_vit python bench.py --steps 10 --warmup 3 --no-cpu-baseline
TAILN=2 run bench_swin python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload swin_s --no-e2e
WLS="vit_b16 swin_s" bash tools/gpu_ncu_lists.sh
python tools/ncu_agg.py gpurun_out/launches_vit_b16.csv > gpurun_out/agg_vit.txt 2>&1
python tools/ncu_agg.py gpurun_out/launches_swin_s.csv > gpurun_out/agg_swin.txt 2>&1
cat gpurun_out/agg_vit.txt gpurun_out/agg_swin.txt
