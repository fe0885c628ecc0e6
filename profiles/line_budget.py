#!/usr/bin/env python
"""Executed instructions of one kernel per source file and line range: joins an ncu
`--page source --csv` dump with the nvdisasm -g -c line table of the same cubin.
usage: line_budget.py <ncu_source.csv> <nvdisasm output> <mangled kernel substring> <segments of the launch> [file:lo-hi=label ...]"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, kern, segments = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
ranges = []
for spec in sys.argv[5:]:
    loc, label = spec.split("=")
    f, r = loc.split(":")
    lo, hi = r.split("-")
    ranges.append((f, int(lo), int(hi), label))
lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith("//--------------------- .text.") and kern in l)
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith("//--------------------- ")), len(lines))
cur, seq = ("?", 0), []
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        seq.append(cur)
rows = list(csv.reader(open(src_csv)))
hi_ = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[hi_]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi_ + 1:] if len(r) == len(hdr)]
assert len(seq) == len(data), (len(seq), len(data))
agg = defaultdict(lambda: [0., 0., 0., 0])
for (f, ln), r in zip(seq, data):
    label = "%s (other)" % f
    for rf, lo, hi, lab in ranges:
        if rf == f and lo <= ln <= hi:
            label = lab
            break
    a = agg[label]
    a[0] += float(r[ix["Instructions Executed"]] or 0)
    a[1] += float(r[ix["Thread Instructions Executed"]] or 0)
    a[2] += float(r[ix["# Samples"]] or 0)
    a[3] += 1
W, T, S = (sum(a[k] for a in agg.values()) for k in range(3))
print("%s: %.1f warp instructions, %.0f thread instructions per segment (%d segments), %.1f lanes per instruction" % (kern, W / segments, T / segments, segments, T / W))
for n, (w, t, s, k) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-34s %5.1f %% of warp instr (%6.1f / segment)  %5.1f %% of samples  %4.1f lanes  %4d SASS" % (n, 100 * w / W, w / segments, 100 * s / S, t / max(w, 1), k))
