#!/bin/bash
# weight-arena guard test, Twins-SVT-S* bench line, then the whole GPU suite
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-${CUT:-300}; }
: > gpurun_out/summary.txt
TMO=300 TAILN=4 run t_arena python -m pytest tests/test_step_ops.py -q -m gpu --no-header -p no:cacheprovider -k "arena" -x
TAILN=1 CUT=1200 run bench_twins_s_n1 python bench.py --only --workload twins_s --steps 40 --eager-baseline
grep -E "attention|layernorm" gpurun_out/breakdown_twins_s_n1.txt | head -6
TMO=900 TAILN=4 run t_all python -m pytest tests/ -q -m gpu --no-header -p no:cacheprovider
