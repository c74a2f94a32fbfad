#!/usr/bin/env python
"""Attribute an ncu SASS-level source page (ncu -i X.ncu-rep --page source --csv) to CUDA source
lines using nvdisasm -g output of the same cubin.  Usage:
    ncu_by_line.py source.csv tile.sass <kernel-name-substring> [top]"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l)
cur, inl, addr2line = None, None, {}
for l in lines[start + 1:]:
    if l.startswith(".text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = int(m.group(2))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        addr2line[int(m.group(1), 16)] = cur

rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) > 5]
ix = {h: i for i, h in enumerate(hdr)}
base = int(data[0][0], 16)
agg = defaultdict(lambda: [0.0, 0.0, defaultdict(float)])
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_i = tot_s = 0.0
for r in data:
    off = int(r[0], 16) - base
    ln = addr2line.get(off)
    ni = float(r[ix["Instructions Executed"]] or 0)
    ns = float(r[ix["# Samples"]] or 0)
    a = agg[ln]
    a[0] += ni
    a[1] += ns
    for h in stall_cols:
        v = float(r[ix[h]] or 0)
        if v:
            a[2][h] += v
    tot_i += ni
    tot_s += ns
srcfile = None
for l in lines:
    m = re.search(r'//## File "([^"]+)"', l)
    if m:
        srcfile = m.group(1)
        break
try:
    srclines = open(srcfile).read().split("\n")
except Exception:
    srclines = []
print(f"total warp-instructions {tot_i:.3e}, samples {tot_s:.0f}")
for ln, (ni, ns, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    txt = srclines[ln - 1].strip()[:90] if ln and ln <= len(srclines) else ""
    s3 = ",".join(f"{k[6:]}:{v / max(ns, 1) * 100:.0f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"L{ln}: inst {ni / tot_i * 100:5.2f}%  samp {ns / tot_s * 100:5.2f}%  [{s3}]  {txt}")
