#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.
usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > x.csv; python tools/ncu_lines.py x.csv [top]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = defaultdict(lambda: [0.0, 0.0, 0.0, ""])   # inst, thread inst, samples, text
stalls = defaultdict(lambda: defaultdict(float))
fname, hdr, cur = None, None, None
for r in csv.reader(open(path)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Line No":
        hdr = r; iI = hdr.index("Instructions Executed"); iT = hdr.index("Thread Instructions Executed"); iS = hdr.index("# Samples")
        st_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) <= iS:
        continue
    if r[0].strip():
        cur = (fname, int(r[0])); agg[cur][3] = r[1].strip()
    if cur is None:
        continue
    try:
        agg[cur][0] += float(r[iI] or 0); agg[cur][1] += float(r[iT] or 0); agg[cur][2] += float(r[iS] or 0)
        for i, h in st_cols:
            stalls[cur][h] += float(r[i] or 0)
    except ValueError:
        pass
ti = sum(v[0] for v in agg.values()); ts = sum(v[2] for v in agg.values())
print("total warp instructions %.4g, samples %.4g" % (ti, ts))
byfile = defaultdict(lambda: [0.0, 0.0])
for (f, l), v in agg.items():
    byfile[f][0] += v[0]; byfile[f][1] += v[2]
for f, v in sorted(byfile.items(), key=lambda x: -x[1][1]):
    print("  file %-28s inst %5.1f%%  samples %5.1f%%" % (f, 100 * v[0] / ti, 100 * v[1] / ts))
print("top lines by stall samples:")
for (f, l), v in sorted(agg.items(), key=lambda x: -x[1][2])[:top]:
    s = stalls[(f, l)]; tops = sorted(s.items(), key=lambda x: -x[1])[:2]
    print("%-16s:%4d inst %4.1f%% samp %4.1f%% lanes %4.1f  %-22s| %s" % (f[:16], l, 100 * v[0] / ti, 100 * v[2] / ts, v[1] / v[0] if v[0] else 0,
                                                                  ",".join("%s%.0f%%" % (k.replace("stall_", ""), 100 * x / max(v[2], 1)) for k, x in tops), v[3][:90]))
