#!/usr/bin/env python
"""Aggregate an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv --print-source=cuda,sass) by CUDA
source line: instructions executed, thread-level SIMD efficiency and stall samples. Usage: source_hotspots.py file.csv [topN]"""
import csv
import sys
import collections

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
# find header rows ("Line No" first cell); sections repeat per file
agg = collections.OrderedDict()
cur_file = ""
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 5:
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    def col(name, idx=0):
        idxs = [i for i, h in enumerate(hdr) if h == name]
        return r[idxs[idx]] if idxs else "0"
    def num(s):
        try:
            return float(s.replace(",", ""))
        except ValueError:
            return 0.0
    key = (cur_file, line)
    a = agg.setdefault(key, dict(src=col("Source", 0)[:110], inst=0.0, tinst=0.0, samples=0.0, long_sb=0.0, wait=0.0, branch=0.0))
    a["inst"] += num(col("Instructions Executed")); a["tinst"] += num(col("Thread Instructions Executed"))
    a["samples"] += num(col("# Samples")); a["long_sb"] += num(col("stall_long_sb")); a["wait"] += num(col("stall_wait"))
    a["branch"] += num(col("stall_branch_resolving"))
tot_i = sum(a["inst"] for a in agg.values()) or 1
tot_s = sum(a["samples"] for a in agg.values()) or 1
tot_t = sum(a["tinst"] for a in agg.values())
print(f"total warp-inst {tot_i:.3e}  thread-inst {tot_t:.3e}  avg threads/inst {tot_t / tot_i:.2f}  samples {tot_s:.0f}")
print(f"{'file:line':28s} {'inst%':>6s} {'thr/inst':>8s} {'smp%':>6s} {'longsb%':>7s} {'wait%':>6s}  source")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    eff = a["tinst"] / a["inst"] if a["inst"] else 0
    print(f"{f + ':' + str(l):28s} {100 * a['inst'] / tot_i:6.2f} {eff:8.1f} {100 * a['samples'] / tot_s:6.2f} {100 * a['long_sb'] / tot_s:7.2f} {100 * a['wait'] / tot_s:6.2f}  {a['src'].strip()}")
