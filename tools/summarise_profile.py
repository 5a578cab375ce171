#!/usr/bin/env python
"""Summarise gpurun_out ncu artefacts into profiles/ (tracked): launch-list shares and the key metrics of
full captures. Usage: tools/summarise_profile.py TAG [kernel-report.ncu-rep ...]"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "gpurun_out")
PR = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]


def launches(tag):
    path = os.path.join(GO, f"launches_{tag}.csv")
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(PR, f"launches_{tag}.md"), "w") as f:
        f.write(f"# ncu launch list `{tag}` (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n")
        f.write("command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline`\n\n")
        f.write("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, (n, t) in agg.items():
            f.write(f"| `{k}` | {n} | {t / 1e6:.3f} | {100 * t / tot:.1f}% | {t / n / 1e3:.1f} |\n")
    print(open(os.path.join(PR, f"launches_{tag}.md")).read())


def full(rep, tag):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {}
        for i, h in enumerate(hdr):
            if h in KEYS or h in ("Kernel Name", "ID") or ("pcsamp_warps_issue_stalled" in h and not h.endswith("_not_issued")):
                d[h] = vals[i] + (" " + units[i] if units[i] else "")
        try:
            inst = float(vals[hdr.index("smsp__inst_executed.sum")].replace(",", ""))
            cyc = float(vals[hdr.index("smsp__cycles_active.avg")].replace(",", ""))
            d["derived.issue_slot_utilisation_pct"] = f"{100 * inst / (cyc * 148 * 4):.1f}"
        except Exception:
            pass
        res.append(d)
    name = os.path.splitext(os.path.basename(rep))[0]
    with open(os.path.join(PR, f"{name}.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


def _bytes(v):
    x, unit = v.split()[0].replace(",", ""), v.split()[1].lower()
    return float(x) * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]


def traffic(tag):
    """profiles/ncu_traffic.json: measured DRAM bytes (read + write) per launch of the dominant kernels, taken from the full
    captures of this tag; bench.py copies the matching entry into roofline.traffic."""
    out = {}
    try:                                     # keep the entries of earlier captures that this tag does not replace
        out = json.load(open(os.path.join(PR, "ncu_traffic.json")))
    except Exception:
        pass
    out.update({"tag": tag, "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, one `ncu --set full` capture each (inst10m, N = 1); "
                                    "each entry names the capture it comes from (`source`)"})
    tp = os.path.join(PR, f"prof_trace_{tag}.json")
    if os.path.exists(tp):
        ks = json.load(open(tp))
        per = {("stage%d" % i): _bytes(k["dram__bytes_read.sum"]) + _bytes(k["dram__bytes_write.sum"]) for i, k in enumerate(ks)}
        def _f(k, key):
            try:
                return float(str(k.get(key, "nan")).split()[0].replace(",", ""))
            except ValueError:
                return None
        out["k_trace"] = {"workload": "inst10m", "dram_bytes_per_frame": sum(per.values()), "per_stage": per, "source": f"profiles/prof_trace_{tag}.json",
                          "issue_slot_utilisation_pct": {("stage%d" % i): _f(k, "derived.issue_slot_utilisation_pct") for i, k in enumerate(ks)},
                          "lanes_per_instruction": {("stage%d" % i): _f(k, "smsp__thread_inst_executed_per_inst_executed.ratio") for i, k in enumerate(ks)},
                          "l1_hit_pct": {("stage%d" % i): _f(k, "l1tex__t_sector_hit_rate.pct") for i, k in enumerate(ks)},
                          "l2_hit_pct": {("stage%d" % i): _f(k, "lts__t_sector_hit_rate.pct") for i, k in enumerate(ks)}}
    bp = os.path.join(PR, f"prof_build_{tag}.json")
    if os.path.exists(bp):
        ks = json.load(open(bp))
        per = collections.OrderedDict()
        for k in ks:
            name = k["Kernel Name"].split("(")[0].split("::")[-1]
            per[name] = per.get(name, 0.0) + _bytes(k["dram__bytes_read.sum"]) + _bytes(k["dram__bytes_write.sum"])
        out["build"] = {"workload": "inst10m", "dram_bytes_per_build": sum(per.values()), "per_kernel": per, "source": f"profiles/prof_build_{tag}.json"}
    json.dump(out, open(os.path.join(PR, "ncu_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    os.makedirs(PR, exist_ok=True)
    tag = sys.argv[1]
    launches(tag)
    for rep in sys.argv[2:]:
        full(rep, tag)
    traffic(tag)
