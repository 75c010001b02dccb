#!/bin/bash
# multi-GPU check: the driver's launch line at N ranks (default run with every leg), then the same step without the
# overlapped all-reduce (A/B).   bash tools/gpu_multi.sh <N>
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N > gpurun_out/bench_default_n$N.log 2> gpurun_out/bench_default_n$N.err; echo "default exit=$?"
tail -n 1 gpurun_out/bench_default_n$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
def f(x): return None if x is None else round(x,1)
print('vit_b16 N=%d value %.0f img/s  %.2f ms  e2e %.0f' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value']))
print('  allreduce:', d['config'].get('allreduce'))
print('  no_graph  %.0f img/s %.2f ms (%s)' % (d['no_graph']['value'], d['no_graph']['ms_per_step'], d['no_graph']['reducer']))
print('  eager_gpu %.0f img/s per GPU;  speedup_vs_eager_gpu %.2f' % (d['eager_baseline']['value'], d['speedup_vs_eager_gpu']))
s=d['swin_s']; print('  swin_s    %.0f img/s %.2f ms  e2e %.0f' % (s['value'], s['ms_per_step'], s['e2e']['value']))
n=d['dino_deit_s']; print('  dino      ', n if 'error' in n else '%.0f src img/s %.2f ms' % (n['value'], n['ms_per_step']))
"
tail -n 5 gpurun_out/bench_default_n$N.err
timeout 600 $TR bench.py --gpus $N --only --no-overlap --no-optimizer-leg --no-e2e 2> gpurun_out/bench_nooverlap_n$N.err | tail -n 1 > gpurun_out/bench_nooverlap_n$N.log; echo "no-overlap exit=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_nooverlap_n$N.log').read())
print('no-overlap: value %.0f img/s  %.2f ms' % (d['value'], d['ms_per_step']))"
