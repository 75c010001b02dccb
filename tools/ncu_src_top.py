#!/usr/bin/env python
"""Top stall sites of an `ncu --page source --csv` dump (SASS view): samples, dominant stall reason, instruction."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
data = []
tot = 0
for k, r in enumerate(rows[2:]):
    try:
        s = int(r[ix["# Samples"]])
    except Exception:
        continue
    tot += s
    data.append((s, k, r))
print("total samples", tot)
top = sorted(data, reverse=True)[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]
for s, k, r in sorted(top, key=lambda t: t[1]):
    why = sorted(((int(r[ix[n]] or 0), n) for n in stalls), reverse=True)[:2]
    print(f"{k:5d} {s:6d} {100*s/tot:5.1f}%  {why[0][1][6:]:>14s}:{why[0][0]:<5d} {why[1][1][6:]:>12s}:{why[1][0]:<5d} ex={r[ix['Instructions Executed']]:>8s} | {r[ix['Source']].strip()[:90]}")
