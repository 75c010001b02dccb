#!/usr/bin/env python
"""Per-kernel SASS instruction summary of libvtb200.so (what proves a Blackwell-native kernel, B200_PROFILING.md):
counts of UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), UTMALDG / UTMASTG / UTMAREDG / UBLKCP
(TMA), HMMA (mma.sync: the legacy tensor path), LDGSTS (cp.async), and the total instruction count, for every kernel.
    python tools/sass_summary.py [path/to/libvtb200.so] > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "vision-transformers-pytorch_b200", "vtb200", "libvtb200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCCP", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "HMMA", "LDGSTS", "MUFU"]
kern, cnt, tot = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        tot[kern] += 1
        for k in KEYS:
            if op.startswith(k):
                cnt[kern][k] += 1


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "")
    except Exception:  # noqa: BLE001
        return n


print(f"# SASS summary of {os.path.relpath(so, ROOT)} (cuobjdump -sass, sm_100a), one line per kernel: instruction mnemonics that")
print("# identify the execution path.  UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, UTMA* / UBLKCP = TMA,")
print("# HMMA = mma.sync (legacy tensor path), LDGSTS = cp.async.")
rows = []
for k in tot:
    name = re.sub(r"\(.*", "", demangle(k))
    name = re.sub(r"^void ", "", name)
    rows.append((name, tot[k], cnt[k]))
for name, n, c in sorted(rows):
    tags = " ".join(f"{k}={c[k]}" for k in KEYS if c[k])
    path = "tcgen05" if (c["UTCHMMA"] or c["UTCQMMA"]) else ("mma.sync" if c["HMMA"] else "-")
    print(f"{name[:70]:70s} {n:6d} instr  [{path:8s}] {tags}")
