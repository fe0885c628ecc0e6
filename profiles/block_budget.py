#!/usr/bin/env python
"""Executed instructions of one kernel per basic block (maximal runs of SASS instructions with the
same execution count), from an ncu `--page source --csv` dump joined with the nvdisasm -g -c line
table: where do the warp instructions go, and with how many lanes.
usage: block_budget.py <ncu_source.csv> <nvdisasm output> <mangled kernel substring> <segments of the launch> [top]"""
import csv
import re
import sys
from collections import Counter

src_csv, sass, kern, segments = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
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
        seq.append((cur, l.strip()))
rows = list(csv.reader(open(src_csv)))
hi_ = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[hi_]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi_ + 1:] if len(r) == len(hdr)]
assert len(seq) == len(data), (len(seq), len(data))
blocks, last = [], None
for k, ((loc, text), r) in enumerate(zip(seq, data)):
    ex = float(r[ix["Instructions Executed"]] or 0)
    th = float(r[ix["Thread Instructions Executed"]] or 0)
    sm = float(r[ix["# Samples"]] or 0)
    if last is None or ex != last:
        blocks.append({"first": k, "n": 0, "ex": ex, "th": 0., "sm": 0., "locs": Counter(), "ops": Counter()})
        last = ex
    b = blocks[-1]
    b["n"] += 1
    b["th"] += th
    b["sm"] += sm
    b["locs"]["%s:%d" % loc] += 1
    op = text.split()[1] if not text.split()[1].startswith("@") else text.split()[2]
    b["ops"][op.split(".")[0]] += 1
W = sum(b["n"] * b["ex"] for b in blocks)
T = sum(b["th"] for b in blocks)
S = sum(b["sm"] for b in blocks)
print("%s: %.1f warp instructions, %.0f thread instructions per segment, %.2f lanes per instruction; %d blocks" % (kern, W / segments, T / segments, T / W, len(blocks)))
for b in sorted(blocks, key=lambda b: -b["n"] * b["ex"])[:top]:
    w = b["n"] * b["ex"]
    print("@%5d %4d instr x %6.3f/seg = %5.1f/seg (%4.1f %%)  %4.1f lanes  %4.1f %% samples  %s | %s" % (
        b["first"], b["n"], b["ex"] / segments, w / segments, 100 * w / W, b["th"] / max(w, 1), 100 * b["sm"] / S,
        " ".join("%s*%d" % kv for kv in b["locs"].most_common(3)), " ".join("%s*%d" % kv for kv in b["ops"].most_common(5))))
