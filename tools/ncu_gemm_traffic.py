#!/usr/bin/env python
"""Aggregate an ncu per-launch metrics CSV (duration, DRAM bytes, tensor-pipe %) of the GEMM launches of one step
into the JSON that bench.py reads for `roofline.traffic`."""
import collections, csv, json, sys
src, dst, workload = sys.argv[1], sys.argv[2], sys.argv[3]
lines = [l for l in open(src) if not l.startswith("==")]
per = collections.defaultdict(dict)
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    name = row["Metric Name"]
    if name.startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    elif name.startswith("gpu__time"):
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    per[row["ID"]][name] = v
n = len(per)
rd = sum(d.get("dram__bytes_read.sum", 0) for d in per.values())
wr = sum(d.get("dram__bytes_write.sum", 0) for d in per.values())
ms = sum(d.get("gpu__time_duration.sum", 0) for d in per.values())
tp = [k for k in next(iter(per.values())) if "pipe_tensor" in k]
tw = sum(d.get(tp[0], 0) * d.get("gpu__time_duration.sum", 0) for d in per.values()) / ms if tp and ms else None
json.dump({"workload": workload, "launches": n, "dram_bytes_read_total": rd, "dram_bytes_write_total": wr,
           "traffic_bytes_per_launch": (rd + wr) / n, "duration_ms_total_under_ncu": ms,
           "tensor_pipe_pct_time_weighted": tw,
           "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active... -k regex:gemm_tc_kernel"},
          open(dst, "w"), indent=1)
print(open(dst).read())
