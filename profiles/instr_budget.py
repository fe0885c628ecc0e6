#!/usr/bin/env python
"""Executed warp instructions of k_render per step kind and per path segment.
Joins an ncu `--page source --csv` dump (one row per SASS instruction, in address order) with the
nvdisasm -g line table of the same cubin, and attributes every instruction to the step function
whose source lines it came from (helpers without an unambiguous owner inherit the label of the
preceding instruction).  The function line ranges are read from the sources of the commit the
library was built from.
usage: instr_budget.py <ncu_source.csv> <nvdisasm -g -c output> <git rev of the build> <segments of the launch>"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

src_csv, sass, rev, segments = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])


def source(path):
    return subprocess.check_output(["git", "show", "%s:%s" % (rev, path)]).decode().split("\n")


def starts(lines, table):
    """[(first line, label)] for the (regex, label) pairs, in file order; label None = ambiguous."""
    out = [(0, None)]
    for n, l in enumerate(lines, 1):
        for rx, label in table:
            if re.search(rx, l):
                out.append((n, label))
    return sorted(out)


POOL = starts(source("rtxplay_b200/csrc/rtx_pool.cuh"), [
    (r"^struct DevPool|^struct RegPool", "stack push / pop"), (r"RTX_HD f3 ld3\(", None), (r"RTX_HD void begin_ray\(", "new ray"),
    (r"RTX_HD int32_t pop_next\(", "pop_next"), (r"RTX_HD int finish_step\(", None), (r"RTX_HD int step_node\(", "node step"),
    (r"RTX_HD int step_leaf\(", "leaf step"), (r"RTX_HD int step_thing\(", "thing step"), (r"RTX_HD int step_shade\(", "shading"),
    (r"RTX_HD int step_regen\(", "new path")])
KERN = starts(source("rtxplay_b200/csrc/rtx_kernels.cuh"), [
    (r"k_render\( const __grid_constant__", "setup"), (r"// vote: which step kind", "vote"), (r"case K_NODE:", "sticky-loop ballot"),
    (r"case K_LEAF:", "leaf step"), (r"case K_THING:", "thing step"), (r"case K_SHADE:", "shading"), (r"default: \{   // K_REGEN", "new path"),
    (r"^// first hit of the primary ray", "setup")])
CORE = starts(source("rtxplay_b200/csrc/rtx_core.cuh"), [
    (r"^struct Pcg", "shading"), (r"^struct q4", None), (r"RTX_HD bool tri_test\(", "leaf step"), (r"RTX_HD bool bsphere_miss\(", "thing step"),
    (r"RTX_HD bool better\(", None), (r"RTX_HD float slab\(", "node step"), (r"^template <class Stack>", None),
    (r"RTX_HD_CALL void frame_of\(", "shading"), (r"RTX_HD_CALL void primary_ray\(", "new path"), (r"RTX_HD uint64_t tofix\(", "shading"),
    (r"RTX_HD uint32_t sem_depth\(", None)])


def label(table, ln):
    name = None
    for a, n in table:
        if ln >= a:
            name = n
    return name


lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith("//--------------------- .text.") and "k_renderILb0" in l)
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
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
assert len(seq) == len(data), (len(seq), len(data))
agg, last = defaultdict(lambda: [0., 0., 0.]), "?"
for (f, ln), r in zip(seq, data):
    n = label({"rtx_pool.cuh": POOL, "rtx_kernels.cuh": KERN, "rtx_core.cuh": CORE}.get(f, [(0, None)]), ln)
    if n:
        last = n
    a = agg[last]
    a[0] += float(r[ix["Instructions Executed"]] or 0)
    a[1] += float(r[ix["Thread Instructions Executed"]] or 0)
    a[2] += float(r[ix["# Samples"]] or 0)
W, T, S = (sum(a[k] for a in agg.values()) for k in range(3))
print("k_render: %.1f warp instructions, %.0f thread instructions per segment (%d segments)" % (W / segments, T / segments, segments))
for n, (w, t, s) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-20s %5.1f %% of warp instructions (%6.1f / segment)  %5.1f %% of samples  %4.1f lanes" % (n, 100 * w / W, w / segments, 100 * s / S, t / max(w, 1)))
