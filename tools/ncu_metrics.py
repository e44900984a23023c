#!/usr/bin/env python
"""Aggregates an `ncu --metrics a,b,c --csv` launch log into one row per kernel (markdown): launches, total time, DRAM read /
write bytes, L2 hit rate, SM and DRAM throughput.  usage: ncu_metrics.py <log.csv> [--json out.json]"""
import collections
import csv
import json
import re
import sys


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"^void ", "", r[ix["Kernel Name"]])
        name = re.sub(r"\(.*", "", name)
        name = re.sub(r"<.*", "", name).replace("vcb::", "").replace("prims::", "")
        key = (r[ix["ID"]], name)
        per.setdefault(key, {})[r[ix["Metric Name"]]] = (float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]])
    unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    agg = collections.OrderedDict()
    for (_, name), m in per.items():
        a = agg.setdefault(name, {"n": 0, "ms": 0.0, "rd": 0.0, "wr": 0.0, "l2": 0.0, "sm": 0.0, "dram": 0.0, "regs": 0, "occ": 0.0})
        t = m["gpu__time_duration.sum"]
        ms = t[0] * unit[t[1]]
        a["n"] += 1
        a["ms"] += ms
        for k, f in (("rd", "dram__bytes_read.sum"), ("wr", "dram__bytes_write.sum")):
            if f in m:
                a[k] += m[f][0] * unit[m[f][1]]
        for k, f in (("l2", "lts__t_sector_hit_rate.pct"), ("sm", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                     ("dram", "dram__throughput.avg.pct_of_peak_sustained_elapsed"), ("occ", "sm__warps_active.avg.pct_of_peak_sustained_active")):
            if f in m:
                a[k] += m[f][0] * ms                      # time-weighted
        if "launch__registers_per_thread" in m:
            a["regs"] = int(m["launch__registers_per_thread"][0])
    tot = sum(a["ms"] for a in agg.values())
    print("| kernel | launches | ms | share | DRAM read MB | DRAM write MB | L2 hit % | SM thr % | DRAM thr % | regs | warps active % |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        w = a["ms"] or 1.0
        print(f"| {name} | {a['n']} | {a['ms']:.3f} | {100 * a['ms'] / tot:.1f}% | {a['rd'] / 1e6:.0f} | {a['wr'] / 1e6:.0f} | {a['l2'] / w:.0f} | "
              f"{a['sm'] / w:.0f} | {a['dram'] / w:.0f} | {a['regs']} | {a['occ'] / w:.0f} |")
    print(f"| total | {sum(a['n'] for a in agg.values())} | {tot:.3f} | | {sum(a['rd'] for a in agg.values()) / 1e6:.0f} | {sum(a['wr'] for a in agg.values()) / 1e6:.0f} | | | | | |")
    if "--json" in sys.argv:
        out = {n: a["rd"] + a["wr"] for n, a in agg.items()}
        json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
