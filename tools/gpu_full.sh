#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-1200} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-8} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run t_all python -m pytest tests/ -x -q -m gpu --no-header -p no:cacheprovider
run smoke python -c "import __graft_entry__ as g; g.smoke()"
TAILN=2 run bench_default python bench.py
TAILN=2 run bench_ref python bench.py --impl reference --steps 3 --warmup 1
