#!/usr/bin/env python
"""Per-source-line view of an ncu report for one kernel: joins the SASS page of the report (instruction counts, stall samples)
with the line table of the cubin (`nvdisasm -g`), because the CUDA-source page of the CLI carries no metrics.
usage: ncu_lines.py <report.ncu-rep> <libvcb200.so> <mangled kernel name> [lo_line hi_line]"""
import collections
import csv
import io
import re
import subprocess
import sys
import tempfile
import os


def main():
    rep, so, fun = sys.argv[1:4]
    lo = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    hi = int(sys.argv[5]) if len(sys.argv) > 5 else 10 ** 9
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    base = int(data[0][0], 16)
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    line_of = {}
    cur = None
    inside = False
    for ln in dis.splitlines():
        if ln.startswith("//---------------------"):
            inside = (".text." + fun + " ") in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    for r in data:
        off = int(r[0], 16) - base
        key = line_of.get(off)
        a = agg[key]
        a[0] += int(r[idx["Instructions Executed"]] or 0)
        a[1] += int(r[idx["# Samples"]] or 0)
        for h in stalls:
            v = int(r[idx[h]] or 0)
            if v:
                a[2][h[6:]] += v
    tot_i = sum(a[0] for a in agg.values())
    tot_s = sum(a[1] for a in agg.values())
    print(f"total warp instructions {tot_i}, samples {tot_s}")
    for key in sorted(k for k in agg if k):
        a = agg[key]
        if key[1] < lo or key[1] > hi or (a[0] == 0 and a[1] == 0):
            continue
        print(f"{key[0]}:{key[1]:5d}  instr {a[0]:9d} ({100 * a[0] / tot_i:5.1f}%)  samples {a[1]:6d}  {dict(a[2].most_common(3))}")


if __name__ == "__main__":
    main()
