#!/usr/bin/env python
"""gpurun_out/attn_metrics_<workload>.csv (ncu per-launch metrics of the attention kernels of one step) -> JSON:
per workload and kernel: launches, time-weighted tensor-pipe %, total ms, DRAM bytes.  bench.py reads the committed copy
(profiles/rNN_attn_tensor_pipe.json) for `roofline.attention.tensor_pipe_pct_ncu`."""
import collections, csv, glob, json, os, re, sys
out = {}
for path in sorted(glob.glob(os.path.join(sys.argv[1], "attn_metrics_*.csv"))):
    wl = re.search(r"attn_metrics_(.+)\.csv", path).group(1)
    lines = [l for l in open(path, errors="replace") if not l.startswith("==")]
    per = collections.defaultdict(dict)
    try:
        rows = list(csv.DictReader(lines))
    except Exception:
        continue
    for row in rows:
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        u, name = row["Metric Unit"], row["Metric Name"]
        if name.startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        elif name.startswith("gpu__time"):
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
        per[row["ID"]]["kernel"] = re.sub(r"^.*::", "", row["Kernel Name"].split("(")[0])
        per[row["ID"]][name] = v
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for d in per.values():
        t = d.get("gpu__time_duration.sum", 0.0)
        tp = [v for k, v in d.items() if "pipe_tensor" in k]
        a = agg[d["kernel"]]
        a[0] += 1; a[1] += t; a[2] += (tp[0] if tp else 0.0) * t
        a[3] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    out[wl] = {k: {"launches": n, "ms_under_ncu": round(t, 4), "tensor_pipe_pct": round(w / t, 2) if t else None,
                   "dram_gb": round(b / 1e9, 3), "dram_gbs": round(b / 1e9 / (t * 1e-3), 1) if t else None}
               for k, (n, t, w, b) in agg.items()}
out["source"] = ("ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
                 "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active -k regex:attn_ , one fwd+bwd step per workload")
print(json.dumps(out, indent=1))
