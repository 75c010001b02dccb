#!/bin/bash
# per-launch duration + DRAM bytes of the LayerNorm / column-sum / cast kernels of one step
mkdir -p gpurun_out
W=${1:-vit_b16}
timeout 900 ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active -k regex:"ln_|colsum|cast" -c 120 --csv --log-file gpurun_out/ln_launches_$W.csv python bench.py --workload $W --warmup 3 --nvtx-step > gpurun_out/ncu_ln_$W.log 2>&1
echo "exit=$?"
python - <<PY
import csv,collections
rows=list(csv.reader(open("gpurun_out/ln_launches_$W.csv")))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[hi]; agg=collections.OrderedDict()
for r in rows[hi+1:]:
    d=dict(zip(h,r)); k=(d["ID"],d["Kernel Name"][:50]); agg.setdefault(k,{})[d["Metric Name"]]=float(d["Metric Value"].replace(",",""))
summ=collections.defaultdict(lambda:[0,0.0,0.0,0.0,0.0])
for (i,k),m in agg.items():
    s=summ[k]; s[0]+=1; s[1]+=m.get("gpu__time_duration.sum",0); s[2]+=m.get("dram__bytes_read.sum",0); s[3]+=m.get("dram__bytes_write.sum",0); s[4]+=m.get("smsp__issue_active.avg.pct_of_peak_sustained_active",0)
for k,s in summ.items():
    print(f"{k:52s} n={s[0]:3d} avg {s[1]/s[0]/1e3:8.1f} us  rd {s[2]/s[0]/1e6:8.1f} MB wr {s[3]/s[0]/1e6:8.1f} MB  -> {(s[2]+s[3])/max(s[1],1):7.2f} GB/s*1e0 issue {s[4]/s[0]:.0f}%")
PY
