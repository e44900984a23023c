#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel for the LAST step in the file
(launches from the last k_meta_init on, or everything with --all)."""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    whole = "--all" in sys.argv
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"<.*", "", name).split("::")[-1]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
        rows.append((name, v, row["Grid Size"], row["Block Size"]))
    starts = [i for i, r in enumerate(rows) if r[0] == "k_meta_init"]
    if not whole and starts:
        rows = rows[starts[-1]:]
    agg = collections.OrderedDict()
    for n, v, g, b in rows:
        a = agg.setdefault(n, [0.0, 0, g, b])
        a[0] += v
        a[1] += 1
    tot = sum(a[0] for a in agg.values())
    print("| kernel | launches | ms | share | grid | block |\n|---|---|---|---|---|---|")
    for n, (v, c, g, b) in sorted(agg.items(), key=lambda x: -x[1][0]):
        print(f"| {n} | {c} | {v:.3f} | {100 * v / tot:.1f}% | {g} | {b} |")
    print(f"| total | {sum(a[1] for a in agg.values())} | {tot:.3f} | | | |")


if __name__ == "__main__":
    main()
