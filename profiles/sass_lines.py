#!/usr/bin/env python
"""Joins an ncu `--page source --csv` SASS dump of one kernel with nvdisasm -g line info
(same instruction order) and prints the hottest source lines.
usage: sass_lines.py <ncu_source.csv> <nvdisasm -g -c output> <mangled kernel substring>"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

# nvdisasm: instruction lines look like "        /*0000*/   LDC R1, c[0x0][0x37c] ;" preceded by "//## File "...", line N"
lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith("//--------------------- .text.") and kern in l)
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith("//--------------------- ")), len(lines))
cur = ("?", 0)
inl = ""
seq = []
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        inl = m.group(3)
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        seq.append(cur)

rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
print("sass instr: nvdisasm %d, ncu %d" % (len(seq), len(data)))
agg = defaultdict(lambda: [0., 0., 0.])
tot = [0., 0., 0.]
for k, r in enumerate(data):
    key = seq[k] if k < len(seq) else ("?", 0)
    v = [float(r[ix["# Samples"]] or 0), float(r[ix["Instructions Executed"]] or 0), float(r[ix["Thread Instructions Executed"]] or 0)]
    for j in range(3):
        agg[key][j] += v[j]
        tot[j] += v[j]
print("total samples %d, warp instr %.3g, thread instr %.3g, avg threads/instr %.2f" % (tot[0], tot[1], tot[2], tot[2] / tot[1]))
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% samples %5.1f%% instr  thr/instr %5.1f  %s:%d" % (100 * v[0] / tot[0], 100 * v[1] / tot[1], v[2] / max(v[1], 1), key[0], key[1]))
