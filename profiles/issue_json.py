#!/usr/bin/env python
"""profiles/rNN_k_render_issue.json from an ncu `--page raw --csv` dump of one render-kernel launch:
warp instructions and lanes per instruction per path segment -- what bench.py's roofline ("issue") uses.
usage: issue_json.py <ncu_raw.csv> <segments of the launch> <kernel name> <command the capture ran> <out.json>"""
import csv
import json
import sys

raw, segments, kernel, command, out = sys.argv[1], float(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5]
rows = list(csv.reader(open(raw)))
hdr, vals = rows[0], rows[2]
g = lambda k: float(vals[hdr.index(k)].replace(",", ""))
inst = g("smsp__inst_executed.sum")
d = {"kernel": kernel, "command": command, "segments": int(segments), "warp_instructions": int(inst),
     "warp_instructions_per_segment": inst / segments,
     "lanes_per_instruction": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
     "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
     "l1_data_pipe_pct": g("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
     "l1_hit_pct": g("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": g("lts__t_sector_hit_rate.pct"),
     "registers": int(g("launch__registers_per_thread")), "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
     "gpu_time_ms": g("gpu__time_duration.sum") / (1e6 if g("gpu__time_duration.sum") > 1e5 else 1.)}
json.dump(d, open(out, "w"), indent=1)
print(json.dumps(d))
