#!/usr/bin/env python3
"""SASS opcode mix of an `ncu --page source --csv --print-source cuda,sass` dump: executed warp instructions per opcode, and the opcodes of
given CUDA lines.  usage: python tools/ncu_sass_mix.py src.csv [file:line ...]"""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
fname = hdr = cur = None
ops = defaultdict(float); line_ops = defaultdict(lambda: defaultdict(float)); tot = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split('/')[-1]; continue
    if r[0] == "Line No":
        hdr = r; iI = hdr.index("Instructions Executed"); continue
    if hdr is None:
        continue
    if r[0].strip():
        try:
            cur = (fname, int(r[0]))
        except ValueError:
            cur = None
        continue
    if len(r) > iI and r[2].startswith('0x'):
        try:
            n = float(r[iI])
        except ValueError:
            continue
        sass = re.sub(r'^@!?U?P\d+\s+', '', r[3].strip())
        full = sass.split()[0]
        ops[full.split('.')[0]] += n; tot += n
        if cur:
            line_ops[cur][full] += n
print("total sass warp-inst %.4g" % tot)
for k, v in sorted(ops.items(), key=lambda x: -x[1])[:45]:
    print("%-10s %6.2f%%" % (k, 100 * v / tot))
for arg in sys.argv[2:]:
    f, l = arg.split(':')
    print(arg, sorted(line_ops[(f, int(l))].items(), key=lambda x: -x[1])[:14])
